# coding=utf-8
"""`python main.py --config <json>`: the reference's entry point, served by cfd_taichi_b200/main.py."""
from cfd_taichi_b200.main import main

if __name__ == "__main__":
    main()
