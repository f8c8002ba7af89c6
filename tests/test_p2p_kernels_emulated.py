"""The peer-memory transport of the fused CG iteration (csrc/p2p.cuh + the p2p branches of cg_update_p_kernel,
spmv_sell2_kernel<.., PEER>, cg_fused_update_kernel, cg_check_kernel) executed by R = 2 ... 8 EMULATED GPUs on the CPU:
every rank is a host thread that runs the product's kernels unchanged (csrc/kernel_emul.hpp: blocks of cooperative
fibers, __syncthreads / shuffles / ballots, static shared memory, real atomics for the system-scope loads and stores),
launched in the order and with the sequence numbering of csrc/api_solver.inc, arenas wired as p2p_setup does, ranks
de-synchronised by random delays.  On hardware this transport has run on 2 GPUs only; here the 4- and 8-rank protocol
(all-to-all neighbourhoods of the RCB cube partitions AND slab partitions in which most ranks are not neighbours) is
checked against the oracle's partitioned CG: identical residual history on every rank, equal to the oracle's to
rounding, same iterate X after K iterations.  A negative control (SpMV not waiting for the neighbours' halo flags)
must be caught."""
import ctypes as C

import numpy as np
import pytest

import fb_util as util
from frontistr_b200 import meshgen, partition

dpt, ipt = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _run(parts, iters=24, sigma=0, seed=1, break_halo_wait=0):
    from oracle import oracle as orc
    import __graft_entry__ as g
    em = C.CDLL(g.build_emul())
    cnt = util.cube_cnt(tol=1e-30, maxit=iters)
    R = len(parts)
    mats = [util.oracle_problem(p, cnt) for p in parts]
    keep = []

    def parr(arrs, typ):
        a = (typ * R)()
        for i, x in enumerate(arrs):
            keep.append(x)
            a[i] = x.ctypes.data_as(typ)
        return a

    i32 = lambda x: np.ascontiguousarray(x, np.int32) if len(x) else np.zeros(1, np.int32)
    f64 = lambda x: np.ascontiguousarray(x, np.float64)
    N = np.array([A.N for A in mats], np.int32); NP = np.array([A.NP for A in mats], np.int32)
    nn = np.array([p.n_neighbor_pe for p in parts], np.int32)
    X = [np.zeros(3 * A.NP) for A in mats]; H = [np.zeros(iters) for _ in mats]; err = np.zeros(R, np.int32)
    rc = em.fbemul_p2p_cg(R, N.ctypes.data_as(ipt), NP.ctypes.data_as(ipt),
                          parr([f64(A.D) for A in mats], dpt), parr([f64(A.AL) for A in mats], dpt), parr([f64(A.AU) for A in mats], dpt),
                          parr([i32(A.indexL) for A in mats], ipt), parr([i32(A.itemL) for A in mats], ipt),
                          parr([i32(A.indexU) for A in mats], ipt), parr([i32(A.itemU) for A in mats], ipt),
                          parr([f64(A.B) for A in mats], dpt), nn.ctypes.data_as(ipt), parr([i32(p.neighbor_pe) for p in parts], ipt),
                          parr([i32(p.import_index) for p in parts], ipt), parr([i32(p.import_item) for p in parts], ipt),
                          parr([i32(p.export_index) for p in parts], ipt), parr([i32(p.export_item) for p in parts], ipt),
                          iters, C.c_double(1e-30), sigma, 1, seed, break_halo_wait, parr(X, dpt), parr(H, dpt), err.ctypes.data_as(ipt))
    assert rc == 0
    mats_o = [util.oracle_problem(p, cnt) for p in parts]
    orc.precond_clear()
    ro = orc.solve(parts, mats_o, hist_len=iters)      # stops at MAXIT = iters (code 3001) with the iterate in X
    orc.precond_clear()
    return mats_o, X, H, err, np.array(ro["hist"][:iters])


def _check(parts, **kw):
    mats_o, X, H, err, ho = _run(parts, **kw)
    assert not err.any(), err                                   # no wait timed out, no CG stop event
    for r in range(1, len(parts)):
        assert np.array_equal(H[r], H[0])                       # every rank sums the partials in rank order: bit-identical scalars
    assert np.max(np.abs(H[0] - ho) / ho) <= 1e-11              # (measured: 1e-14)
    for r, A in enumerate(mats_o):
        assert util.rel_l2(X[r][: 3 * A.N], A.X[: 3 * A.N]) <= 1e-11


@pytest.mark.parametrize("axes", ["x", "xy", "xyz"])
def test_peer_memory_cg_on_rcb_partitions(axes):
    """2 / 4 / 8 ranks, the partitions bench.py uses (every rank is a neighbour of every other one)."""
    _check(partition.partition(util.distort(meshgen.hex_cube(5)), axes), seed=len(axes))


@pytest.mark.parametrize("nparts", [3, 5, 8])
def test_peer_memory_cg_on_slab_partitions(nparts):
    """Chains: rank r only touches r-1 and r+1, so most pairs of ranks exchange partial sums but no halo."""
    _check(util.simple_partition(util.distort(meshgen.hex_cube(8, nx=16, ny=2, nz=2)), nparts), seed=10 + nparts)


def test_peer_memory_cg_tet10_and_sell_c_sigma():
    """tet10 rows differ a lot in length: the SELL-C-sigma row permutation (never run on hardware) in the same kernels."""
    mesh = meshgen.tet10_cube(3)
    _check(partition.partition(mesh, "xy"), seed=4)
    _check(partition.partition(mesh, "xy"), seed=5, sigma=64)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_peer_memory_cg_many_schedules(seed):
    _check(partition.partition(util.distort(meshgen.hex_cube(4)), "xyz"), seed=seed, iters=16)


def test_negative_control_spmv_without_halo_wait_is_caught():
    """If the SpMV does not wait for the neighbours' flags it reads halo values of the previous iteration every now and
    then: the residual history must leave the oracle's (otherwise this harness could not see a protocol bug)."""
    bad = 0
    for seed in range(6):
        mats_o, X, H, err, ho = _run(partition.partition(util.distort(meshgen.hex_cube(5)), "xyz"), seed=40 + seed, break_halo_wait=1)
        if np.max(np.abs(H[0] - ho) / ho) > 1e-6 or any(not np.array_equal(H[r], H[0]) for r in range(1, 8)):
            bad += 1
    assert bad >= 1
