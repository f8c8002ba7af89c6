"""N > 1 host path on CPU: two real processes (torch.distributed, gloo, 127.0.0.1), one subdomain each.

Every rank partitions the same global mesh on its own (as bench.py does), builds its local graph with the
product's host-side hecmw_mat_con, and then runs the reference's CG recurrences with
  * hecmw_update_3_R  = send/recv of the export/import tables over gloo,
  * hecmw_allreduce_R1 = dist.all_reduce,
using the oracle for the local arithmetic (no GPU here).  The result must equal the serial oracle solve:
this pins the per-rank partition tables, their cross-rank consistency over a real transport and the
collective structure the GPU path uses (halo exchange before every SpMV, 1- and 2-value all-reduces).
"""
import os
import socket
import sys

import numpy as np
import pytest

import fb_util as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _update_3_R(dist, torch, m, x):
    """hecmw_update_3_R (communication/hecmw_comm_f.F90:669-694; pack/unpack hecmw_solver_SR_33.F90:42-124)."""
    reqs, recv = [], []
    for a, q in enumerate(m.neighbor_pe):
        ex = m.export_item[m.export_index[a]: m.export_index[a + 1]] - 1
        sb = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, 3)[ex]))
        reqs.append(dist.isend(sb, int(q)))
        im = m.import_item[m.import_index[a]: m.import_index[a + 1]] - 1
        rb = torch.zeros((len(im), 3), dtype=torch.float64)
        reqs.append(dist.irecv(rb, int(q)))
        recv.append((im, rb))
    for r in reqs:
        r.wait()
    for im, rb in recv:
        x.reshape(-1, 3)[im] = rb.numpy()


def _worker(rank, world, port, n, tol, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frontistr_b200 import hecmw, meshgen, partition
    from oracle import oracle as orc
    orc.set_num_threads(1)
    mesh = util.distort(meshgen.hex_cube(n))
    dom = partition.rcb_domains(mesh, "x" if world == 2 else "xy")
    m = partition.local_mesh(mesh, dom, rank, world)
    cnt = util.cube_cnt(tol=tol)
    # product host graph build must equal the oracle's on the local mesh
    Ap = hecmw.hecmw_mat_con(m)
    A = util.oracle_problem(m, cnt)
    assert np.array_equal(Ap.indexL, A.indexL) and np.array_equal(Ap.itemU, A.itemU)
    N3 = 3 * A.N

    def allsum(*v):
        t = torch.tensor(v, dtype=torch.float64)
        dist.all_reduce(t)
        return t.numpy()

    pc = orc.Precond(A, precond=3)
    x = np.zeros(3 * A.NP); r = np.zeros(3 * A.NP); p = np.zeros(3 * A.NP)
    bnrm2 = allsum(A.B[:N3] @ A.B[:N3])[0]
    r[:N3] = A.B[:N3]                      # x0 = 0
    it, rho1 = 0, 0.0
    for it in range(1, 2000):
        z = r.copy(); z[N3:] = 0.0
        pc.apply(z)
        rho = allsum(r[:N3] @ z[:N3])[0]
        p[:N3] = z[:N3] if it == 1 else z[:N3] + (rho / rho1) * p[:N3]
        _update_3_R(dist, torch, m, p)
        q = orc.matvec(A, p)
        alpha = rho / allsum(p[:N3] @ q[:N3])[0]
        x[:N3] += alpha * p[:N3]
        r[:N3] -= alpha * q[:N3]
        rho1 = rho
        if np.sqrt(allsum(r[:N3] @ r[:N3])[0] / bnrm2) <= tol:
            break
    _update_3_R(dist, torch, m, x)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), x=x, gid=m.global_node_ID, iters=it,
             n_neighbor=m.n_neighbor_pe)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_cg_over_gloo_equals_serial(world, tmp_path):
    import torch.multiprocessing as mp
    from frontistr_b200 import meshgen
    from oracle import oracle as orc
    n, tol = 6, 1e-10
    mp.spawn(_worker, args=(world, _free_port(), n, tol, str(tmp_path)), nprocs=world, join=True)
    mesh = util.distort(meshgen.hex_cube(n))
    A = util.oracle_problem(mesh, util.cube_cnt(tol=tol))
    orc.precond_clear()
    r1 = orc.solve(mesh, A)
    orc.precond_clear()
    for rank in range(world):
        d = np.load(os.path.join(str(tmp_path), "r%d.npz" % rank))
        assert int(d["n_neighbor"]) == 1
        assert abs(int(d["iters"]) - r1["iter"]) <= 2
        ref = A.X.reshape(-1, 3)[d["gid"] - 1]
        assert util.rel_l2(d["x"].reshape(-1, 3), ref) < 1e-8      # halo included


def test_p2p_protocol_model_eight_ranks():
    """Threaded model (csrc/emul_test.cpp: p2pmodel) of the peer-memory protocol of the fused CG iteration — tagged
    partial sums in parity-double-buffered slots written into every rank's table, one halo flag per source rank,
    monotone sequence numbers, a 1-in-50 iteration on the collective path — with 8 de-synchronised ranks on a fully
    connected and on a chain topology (ranks that are NOT neighbours of each other): no consume may ever see another
    sequence number's payload and no halo read another iteration's value.  The hardware run covered N = 2 only; this
    exercises the N = 8 sequencing logic.  Negative control: without the halo wait the model does report violations."""
    import ctypes as C
    import __graft_entry__ as g
    em = C.CDLL(g.build_emul())

    def run(nbr, iters, seed, rec, broken=0):
        a = np.ascontiguousarray(nbr, np.int32)
        return em.fbemul_p2p_protocol_model(len(a), a.ctypes.data_as(C.POINTER(C.c_int)), iters, seed, rec, broken)

    full = 1 - np.eye(8, dtype=np.int32)
    chain = np.zeros((8, 8), np.int32)
    for i in range(7):
        chain[i, i + 1] = chain[i + 1, i] = 1
    cube = np.zeros((8, 8), np.int32)            # 2x2x2 blocks that only exchange across faces
    for a in range(8):
        for b in (1, 2, 4):
            cube[a, a ^ b] = 1
    for topo in (full, chain, cube):
        for seed in range(3):
            assert run(topo, 300, seed, 50) == 0
    assert run(1 - np.eye(2, dtype=np.int32), 300, 0, 7) == 0
    assert sum(run(chain, 300, seed, 50, broken=1) for seed in range(3)) > 0
