"""TEST ONLY: run one of the GPU tools on the CPU against the host-compiled library (tools/build_hostemu.py) to catch
Python-level mistakes before a GPU minute is spent:   python tools/host_dry_run.py tools/post_probe.py --n 4 --tet 2
Every number printed is meaningless."""
import ctypes
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import build_hostemu  # noqa: E402
from frontistr_b200 import _lib  # noqa: E402

_lib._lib = _lib.declare(ctypes.CDLL(build_hostemu.build()))
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
