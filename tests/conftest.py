import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# several processes share the cores during the suite (torchrun ranks, the background runs below): the oracle's OpenMP team
# must sleep, not spin, at its barriers (read by libgomp when the oracle library is loaded)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_finish(session):
    """The two long subprocess runs of tests/test_emul_fiber_order.py (race detector with reversed fibers, UBSan build:
    ~1.5 min each, mostly waiting on one core) are started as soon as the collection shows that they are wanted, so that
    they overlap the rest of the CPU suite instead of extending it; the tests themselves only collect the results."""
    if os.environ.get("FB_EMUL_FIBER_ORDER") or os.environ.get("FB_EMUL_UBSAN"):   # (not from inside those runs)
        return
    if session.config.getoption("collectonly", False):
        return
    wanted = {it.name for it in session.items if "test_emul_fiber_order.py::" in it.nodeid}
    if wanted:
        import test_emul_fiber_order as T
        T.start_background_runs(reverse=any("reverse" in n for n in wanted), ubsan=any("undefined_behaviour" in n for n in wanted))
    # likewise the torchrun launches of tests/test_dist_host_api.py (start-up dominated): two at a time, ahead of their tests
    dist = {it.originalname for it in session.items if "test_dist_host_api.py::" in it.nodeid}
    if dist:
        import test_dist_host_api as D
        D.start_background_runs(dist)


def pytest_sessionfinish(session, exitstatus):
    for name in ("test_emul_fiber_order", "test_dist_host_api"):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.stop_background_runs()


@pytest.fixture(scope="session")
def gpu():
    """Initialise the product library on cuda:0; fails loudly when the extension or device is missing."""
    from frontistr_b200 import hecmw
    hecmw.init(0)
    return hecmw
