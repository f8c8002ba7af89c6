"""Multi-GPU parity check (run under torchrun on N GPUs of one box, N in 2/4/8):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check.py [--n 24] [--tet 0] [--precond 3]

Every rank partitions the same global mesh (in-process hecmw_part1/RCB), assembles and solves its subdomain
through the C-ABI with halo exchange + all-reduce over NCCL; rank 0 gathers the displacements by global node id
and compares them with the ORACLE's serial solve of the same system (checker only): iteration count within +-2,
relative L2 error <= 1e-9, final true residual below RESID.  Prints one PASS/FAIL line per case.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/: it uses the oracle as checker)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hex", dest="n", type=int, default=24)
    ap.add_argument("--partition", default="rcb", choices=["rcb", "kmetis", "pmetis"])
    ap.add_argument("--empty-last-rank", action="store_true", help="the last rank owns no node at all (N = NP = 0): the mesh is "
                    "split over world-1 domains with KMETIS (or RCB if world-1 is 2 / 4)")
    ap.add_argument("--timelog", type=int, default=0, help="TIMELOG of the !SOLVER card: 1 YES, 2 VERBOSE (statistics over the ranks)")
    ap.add_argument("--scaling", action="store_true", help="!SOLVER SCALING=YES (Iarray(7) = 1): the scale vector is halo-exchanged")
    ap.add_argument("--tet", type=int, default=0)
    ap.add_argument("--precond", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--iterpremax", type=int, default=1, help="sweeps of the preconditioner (precond_33.f90:86-114); > 1 exchanges Z between sweeps")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl", "p2p-or-nccl"],
                    help="p2p: peer memory or fail loudly; p2p-or-nccl: peer memory if every rank can map it, else NCCL (bench.py's auto)")
    ap.add_argument("--host", action="store_true", help="TEST ONLY: no GPUs - every rank loads the host-compiled library "
                    "(tools/build_hostemu.py; NCCL and CUDA IPC are shared-memory stand-ins) and torch.distributed runs on gloo")
    ap.add_argument("--stress", action="store_true", help="also run the strain / stress recovery (update_3d + nodal_stress_3d) on "
                    "every rank and compare the internal nodes with the oracle's serial fstr_NodalStress3D")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=INT", help="fistr_b200_set_option on every rank (e.g. ssor_fused=1)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    dev = "cpu" if args.host else "cuda"
    if args.host:
        import ctypes
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import build_hostemu
        from frontistr_b200 import _lib
        _lib._lib = _lib.declare(ctypes.CDLL(build_hostemu.build()))
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import fb_util as util
    from frontistr_b200 import hecmw, meshgen, partition
    hecmw.init(0 if args.host else lr)
    hecmw.set_option("p2p", {"p2p": 2, "p2p-or-nccl": 1, "nccl": 0}[args.comm])   # 2 = fail loudly if peer memory cannot be mapped
    hecmw.comm_init_from_torch()
    for kv in args.option:
        hecmw.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    gmesh = meshgen.tet10_cube(args.tet) if args.tet else util.distort(meshgen.hex_cube(args.n))
    if args.empty_last_rank:
        nd = world - 1
        dom = (partition.rcb_domains(gmesh, {2: "x", 4: "xy"}[nd]) if nd in (2, 4) else
               (np.zeros(gmesh.n_node, np.int32) if nd == 1 else partition.metis_domains(gmesh, nd, "KMETIS")[0]))
    elif args.partition == "rcb":
        axes = {2: "x", 4: "xy", 8: "xyz"}[world]
        dom = partition.rcb_domains(gmesh, axes)
    else:      # METHOD=KMETIS / PMETIS: every rank computes the same (deterministic) map; any world size
        dom, _ = partition.metis_domains(gmesh, world, args.partition)
    m = partition.local_mesh(gmesh, dom, rank, world)
    cnt = util.cube_cnt(tol=args.tol, precond=args.precond, iterpremax=args.iterpremax)
    A = util.product_problem(hecmw, m, cnt, to_host=False)
    A.Iarray[6] = 1 if args.scaling else 0
    A.Iarray[21] = args.timelog
    cs0 = hecmw.get_option("comm_steps")
    r = hecmw.hecmw_solve(m, A, device_resident=True)
    comm_steps = hecmw.get_option("comm_steps") - cs0
    # halo consistency: X on external nodes must equal the owner's value (hecmw_update_m_R at CG exit)
    xg = torch.zeros(3 * gmesh.n_node, dtype=torch.float64, device=dev)
    own = torch.from_numpy(m.global_node_ID[: m.nn_internal].astype(np.int64) - 1).to(dev)
    xl = torch.from_numpy(A.X.reshape(-1, 3)[: m.nn_internal].copy()).to(dev)
    xg.view(-1, 3)[own] = xl
    dist.all_reduce(xg)
    xgl = xg.cpu().numpy().reshape(-1, 3)
    halo_err = float(np.abs(A.X.reshape(-1, 3) - xgl[m.global_node_ID - 1]).max(initial=0.0))
    t = torch.tensor([halo_err], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = True
    sg = None
    if args.stress:      # every rank works on its own local mesh (overlap elements included): no exchange needed
        hecmw.fstr_UpdateNewton_linear(m, None, A.X)
        ns = hecmw.fstr_NodalStress3D(m)
        sg = torch.zeros((gmesh.n_node, 13), dtype=torch.float64, device=dev)
        loc = np.concatenate([ns["STRAIN"], ns["STRESS"], ns["MISES"][:, None]], axis=1)[: m.nn_internal]
        sg[own] = torch.from_numpy(np.ascontiguousarray(loc)).to(dev)
        dist.all_reduce(sg)
    if rank == 0:
        from oracle import oracle as orc
        Ao = util.oracle_problem(gmesh, cnt)
        Ao.Iarray[6] = 1 if args.scaling else 0
        orc.precond_clear()
        ro = orc.solve(gmesh, Ao)
        err = util.rel_l2(xgl.reshape(-1), Ao.X)
        it_ok = abs(r["iter"] - ro["iter"]) <= 2
        if args.precond != 3:   # SSOR / BILU are subdomain-local: the count to meet is the oracle's on the SAME decomposition
            parts = partition.partition(gmesh, "", domain=dom, nparts=world)
            mats = [util.oracle_problem(p_, cnt) for p_ in parts]
            for a_ in mats:
                a_.Iarray[6] = 1 if args.scaling else 0
            orc.precond_clear()
            rp = orc.solve(parts, mats, ssor_threads=2)
            it_ok = rp["error"] == 0 and abs(r["iter"] - rp["iter"]) <= 2
            print("     oracle on the same %d subdomains: %d iterations" % (world, rp["iter"]), flush=True)
        # SCALING: CG stops on the residual of the SCALED system; the true residual of the back-scaled one is the oracle's, not < tol
        res_ok = (0.5 * ro["resid2"] <= r["resid2"] <= 2.0 * ro["resid2"]) if args.scaling else r["resid2"] < args.tol
        if args.scaling:
            print("     true residual of the back-scaled system: library %.3e, serial oracle %.3e" % (r["resid2"], ro["resid2"]), flush=True)
        ok = (r["error"] == 0 and it_ok and err <= 1e-9 and res_ok and float(t.item()) == 0.0)
        if r["error"] != 0 and r["error"] == ro["error"] and it_ok and err <= 1e-9:
            # the reference algorithm gives up the same way on this system (e.g. 3003 for two DIAG sweeps on tet10): parity
            print("     both the library and the oracle stop with error %d after %d iterations" % (r["error"], r["iter"]), flush=True)
            ok = float(t.item()) == 0.0
        if sg is not None:
            e, s_, q = orc.update_linear(gmesh, xgl.reshape(-1))
            no = orc.nodal_stress_3d(gmesh, e, s_)
            ref = np.concatenate([no["STRAIN"], no["STRESS"], no["MISES"][:, None]], axis=1)
            serr = float(np.abs(sg.cpu().numpy() - ref).max() / np.abs(ref).max())
            ok = ok and serr <= 1e-10
            print("     nodal strain / stress / Mises on the partitioned mesh vs serial oracle: max rel dev %.2e" % serr, flush=True)
        print("%s dist_check world=%d comm=%s%s mesh=%s precond=%d: iter gpu=%d oracle=%d  rel_l2=%.2e  resid2=%.2e  halo_err=%.1e  neighbors=%d  comm_steps/iter=%.2f"
              % ("PASS" if ok else "FAIL", world, args.comm, (" p2p_active=%d" % hecmw.get_option("p2p_active")) + (" HOST-EMULATION" if args.host else ""), ("tet%d" % args.tet) if args.tet else ("hex%d" % args.n), args.precond,
                 r["iter"], ro["iter"], err, r["resid2"], float(t.item()), m.n_neighbor_pe, comm_steps / max(r["iter"], 1)), flush=True)
    hecmw.comm_finalize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
