"""utils -- config loading with the reference's contract (utils.py:3-11): return the parsed JSON
dict; on any failure print the exception and 'Parsing config file error', then exit with code 3.
Extra helpers serve the headless driver (main.py) of this repository."""
import json
import sys
from pathlib import Path


def read_config(file_name):
    path = Path(file_name)
    try:
        return json.loads(path.read_text())
    except Exception as exc:  # same catch-all as the reference
        print(exc)
        print('Parsing config file error')
        sys.exit(3)


def override_solver(config, name=None, delta_time=None):
    """BASELINE.json's configs run shipped scenes under another solver.name; keep every other block."""
    out = json.loads(json.dumps(config))
    if name is not None:
        out['solver']['name'] = name
    if delta_time is not None:
        out['solver']['delta_time'] = delta_time
    return out
