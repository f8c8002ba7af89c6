"""Intra-warp / intra-block race detector for the product's kernels, on the CPU.

Between two barriers (or shuffles) the threads of a real warp or block are not ordered.  The host emulation
(tests/emul/kernel_emul_impl.hpp) runs the runnable fibers of a block in ascending thread order by default, which would
HIDE a missing __syncwarp / __syncthreads whenever the producer happens to have the lower thread id.  Here the emulated
kernels run again with the fibers scheduled in DESCENDING order and in a fresh RANDOM order per scheduling round: a
kernel whose result depends on that order has a race.  The bodies of the GPU tests (assembly of every formulation, BC,
SpMV, DIAG / SSOR / BILU(0), CG, SCALING, update / nodal stress, total-Lagrangian path, device mat_con) - many of them
bit-exact comparisons - are the judges.  A control kernel with and without its barrier shows the detector has teeth."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_detector_control_kernel():
    import __graft_entry__ as g
    em = C.CDLL(g.build_emul())
    em.fbemul_set_fiber_order.argtypes = [C.c_int, C.c_ulonglong]
    em.fbemul_race_probe.argtypes = [C.c_int, C.POINTER(C.c_double)]
    want = 100.0 + (np.arange(32) + 1) % 32
    res = {}
    for sync in (1, 0):
        for mode, seed in ((0, 0), (1, 0), (2, 5)):
            out = np.zeros(32)
            em.fbemul_set_fiber_order(mode, seed)
            em.fbemul_race_probe(sync, out.ctypes.data_as(C.POINTER(C.c_double)))
            res[sync, mode] = out
    em.fbemul_set_fiber_order(0, 0)
    for mode in (0, 1, 2):
        assert np.array_equal(res[1, mode], want)                 # with the barrier: right under every order
    assert not np.array_equal(res[0, 0], want)                    # without it: ascending order reads unwritten slots ...
    assert np.array_equal(res[0, 1], np.r_[want[:31], -1.0])      # ... descending order only lane 31's
    assert not np.array_equal(res[0, 0], res[0, 1])               # => the result depends on the order: detected


def test_divergent_warp_collectives_are_detected():
    """The lanes of a warp must meet at the SAME __shfl_sync (full masks throughout the product): lanes at different shuffle
    instructions, or some at a shuffle and others at a __syncwarp, is undefined on hardware.  The emulation tags every
    collective with its call site and aborts when the lanes it is about to release disagree - controls: one legal kernel,
    two illegal ones (each in its own process: the detector aborts)."""
    import __graft_entry__ as g
    so = g.build_emul()
    code = ("import ctypes as C, sys; em = C.CDLL(%r); out = (C.c_double * 32)(); "
            "em.fbemul_divergent_probe(int(sys.argv[1]), out); print('returned', out[0])" % so)
    ok = subprocess.run([sys.executable, "-c", code, "0"], capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "returned 1.0" in ok.stdout, ok.stderr[-500:]
    for mode in (1, 2):
        bad = subprocess.run([sys.executable, "-c", code, str(mode)], capture_output=True, text=True, timeout=120)
        assert bad.returncode != 0 and "divergent warp collective" in bad.stderr and "returned" not in bad.stdout, (mode, bad.stderr[-500:])


# the two runs share the machine: the oracle's OpenMP team must not busy-wait on cores the other run is using
_SHARE = dict(OMP_NUM_THREADS=str(max(1, (os.cpu_count() or 2) // 2)), OMP_WAIT_POLICY="PASSIVE")


def _cmd_order(order):
    files = ["tests/test_gpu_suite_on_host_api.py", "tests/test_gpu_tests_on_emulated_kernels.py", "tests/test_p2p_kernels_emulated.py"]
    sel = "not smoke_and_bench and not falls_back and not golden_decks and not nlgeom_newton"      # keep the run under ~1.5 min per order
    return ([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", "-k", sel] + files,
            dict(os.environ, FB_EMUL_FIBER_ORDER=order, **_SHARE))


def _cmd_ubsan():
    return ([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", "tests/test_host_api_fuzz.py",
             "tests/test_gpu_suite_on_host_api.py", "-k",
             "fuzz or assembly_matches or bc_elimination or spmv_bitwise or solve_matches or scaling or bilu0_matches or mat_con_on"],
            dict(os.environ, FB_EMUL_UBSAN="1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1", **_SHARE))


_PROCS = {}


def start_background_runs(reverse=True, ubsan=True):
    """Start the reversed-order run and the UBSan run (two independent pytest subprocesses of ~1.5 min each, on different
    builds of the host-compiled library).  Called by tests/conftest.py right after collection so that they overlap the rest
    of the suite; idempotent."""
    import shutil
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import build_hostemu
    # the builds are made here, one after the other: the subprocesses must not race to write the same generated sources
    if reverse and "reverse" not in _PROCS:
        build_hostemu.build(ubsan=False)
        cmd, env = _cmd_order("reverse")
        _PROCS["reverse"] = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)
    if ubsan and "ubsan" not in _PROCS and shutil.which("g++") is not None:
        build_hostemu.build(ubsan=True)
        build_hostemu.build(ubsan=False)      # (leaves the module's notion of "the" library on the plain build)
        cmd, env = _cmd_ubsan()
        _PROCS["ubsan"] = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)


def stop_background_runs():
    for p in _PROCS.values():
        if p.poll() is None:
            p.kill()
    _PROCS.clear()


@pytest.fixture(scope="module")
def background_runs():
    start_background_runs()
    return _PROCS


def _finish(proc):
    try:
        out, err = proc.communicate(timeout=1500)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, err = proc.communicate()
    tail = out.strip().splitlines()[-1] if out.strip() else err[-500:]
    return proc.returncode, out, err, tail


@pytest.mark.parametrize("order", ["reverse", "random:11"])
def test_kernels_do_not_depend_on_the_order_of_threads_between_barriers(order, background_runs):
    if order != "reverse":
        if os.environ.get("FISTR_B200_EMUL_FULL") != "1":
            pytest.skip("second order of the race detector: FISTR_B200_EMUL_FULL=1 (the reversed order runs by default)")
        cmd, env = _cmd_order(order)
        proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)
    else:
        proc = background_runs["reverse"]
    rc, out, err, tail = _finish(proc)
    assert rc == 0 and " passed" in tail and "failed" not in tail, out[-3000:]


def test_kernels_under_the_undefined_behaviour_sanitizer(background_runs):
    """The whole library compiled for the host with -fsanitize=undefined,alignment,bounds,float-cast-overflow
    -fno-sanitize-recover (tools/build_hostemu.py, FB_EMUL_UBSAN=1): signed overflow in index arithmetic, misaligned
    vector accesses (the fault class of the round-1 incident), out-of-range shifts or casts in ANY kernel abort the run.
    Judges: the randomised host-API suite (mixed meshes, all preconditioners, nonlinear states) and the BC / assembly /
    SpMV / solve bodies of the GPU tests."""
    if "ubsan" not in background_runs:
        pytest.skip("no g++")
    rc, out, err, tail = _finish(background_runs["ubsan"])
    assert rc == 0 and " passed" in tail and "runtime error" not in out + err, (out + err)[-3000:]
