import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Native pieces are compiled once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


def quiet_ps(cfg, **kw):
    """ParticleSystem prints its sizes like the reference does; keep test logs short."""
    import contextlib
    import io
    from cfd_taichi_b200.ParticleSystem import ParticleSystem
    with contextlib.redirect_stdout(io.StringIO()):
        return ParticleSystem(cfg, **kw)


def quiet_solver(cls, ps, cfg):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return cls(ps, cfg)
