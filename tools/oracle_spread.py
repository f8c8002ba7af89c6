#!/usr/bin/env python
"""The REFERENCE ALGORITHM's own CG iteration-count spread on a benched mesh (no GPU, oracle only): the oracle's solve of
the undivided system with several OpenMP thread counts and of the same system decomposed into 2 / 4 / 8 RCB subdomains.
DIAG-preconditioned CG is partition- and thread-invariant mathematically; every variant sums the same dot products in a
different order, so the spread of the counts is the reference's own sensitivity to rounding after ~1500 iterations - the
yardstick for north_star's "same iteration count +-2" at this size (tools/parity_big.py --spread-file uses it).

    python tools/oracle_spread.py --n 150 --tol 1e-10 > profiles/r2_oracle_spread_H150.json"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=150)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--threads", default="8,4", help="OpenMP thread counts of the undivided solves")
    ap.add_argument("--decomps", default="x,xyz", help="RCB axis strings of the decomposed solves (x = 2, xy = 4, xyz = 8 subdomains)")
    args = ap.parse_args()
    import fb_util as util
    from frontistr_b200 import meshgen, partition
    from frontistr_b200.workloads import cube_cnt
    from oracle import oracle as orc
    mesh = meshgen.hex_cube(args.n)
    cnt = cube_cnt(tol=args.tol, precond=3, maxit=50000)
    nth0 = orc.num_threads()
    runs = []
    A = util.oracle_problem(mesh, cnt)
    for t in [int(x) for x in args.threads.split(",") if x]:
        orc.set_num_threads(t)
        A.X[:] = 0.0
        orc.precond_clear()
        t0 = time.time()
        r = orc.solve(mesh, A)
        runs.append({"what": "undivided, %d OpenMP threads" % t, "iter": int(r["iter"]), "resid": float(r["resid"]), "wall_s": time.time() - t0})
        print(runs[-1], file=sys.stderr, flush=True)
    del A
    orc.set_num_threads(nth0)
    for axes in [a for a in args.decomps.split(",") if a]:
        parts = partition.partition(mesh, axes)
        mats = [util.oracle_problem(p, cnt) for p in parts]
        orc.precond_clear()
        t0 = time.time()
        r = orc.solve(parts, mats)
        runs.append({"what": "RCB-%s, %d subdomains, %d OpenMP threads" % (axes, len(parts), nth0), "iter": int(r["iter"]),
                     "resid": float(r["resid"]), "wall_s": time.time() - t0})
        print(runs[-1], file=sys.stderr, flush=True)
        del parts, mats
    orc.precond_clear()
    its = [r["iter"] for r in runs]
    print(json.dumps({"workload": "H%d" % args.n, "tol": args.tol, "precond": 3, "runs": runs, "iter_min": min(its), "iter_max": max(its)}), flush=True)


if __name__ == "__main__":
    main()
