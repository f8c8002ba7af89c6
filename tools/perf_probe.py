"""Development probe (not the bench): assemble a hex cube on the GPU and time the pieces of the hot path."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fb_util as util  # noqa: E402
from frontistr_b200 import hecmw, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--tet", type=int, default=0)
ap.add_argument("--variants", default="0,1,2")
ap.add_argument("--solve", type=int, default=1)
ap.add_argument("--precond", type=int, default=3)
ap.add_argument("--maxit", type=int, default=10000)
ap.add_argument("--bps", default="8")
ap.add_argument("--sigma", default="0", help="comma list of sell_sigma values to time (0 = natural row order)")
args = ap.parse_args()

hecmw.init(0)
print(json.dumps(hecmw.device_info()))
print("hbm copy peak GB/s: %.1f   fp64 fma peak TFLOP/s: %.2f" % (hecmw.hbm_copy_peak(), hecmw.fp64_peak()))
t0 = time.time()
mesh = meshgen.tet10_cube(args.tet) if args.tet else meshgen.hex_cube(args.n)
t1 = time.time()
A = hecmw.hecmw_mat_con(mesh)
A.fstr_mat_init()
t2 = time.time()
print("mesh %.2fs  mat_con %.2fs  nodes %d elems %d NPL %d" % (t1 - t0, t2 - t1, mesh.n_node, mesh.n_elem, A.NPL))
hecmw.mat_structure(mesh, A)
for rep in range(3):
    t = hecmw.fstr_StiffMatrix(mesh, A, to_host=False)
    print("assembly rep%d: kernel %.3f ms  total %.1f ms  -> %.1f M elem/s (kernel)" % (rep, t["kernel_ms"], t["total_ms"], mesh.n_elem / t["kernel_ms"] / 1e3))
A.B = np.zeros(3 * A.NP); A.X = np.zeros(3 * A.NP)
cnt = util.cube_cnt(tol=1e-8, precond=args.precond, maxit=args.maxit)
hecmw.fstr_ass_load_cload(mesh, A, cnt["cload"])
hecmw.mat_upload(A, values=False, rhs=True)
n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
hecmw.hecmw_mat_ass_bc(n, d, v)
from frontistr_b200.mshio import apply_solver_card
apply_solver_card(A, cnt["solver"]); A.Iarray[20] = 0; A.Iarray[21] = 0
nb = A.spmv_bytes()
for sigma in [int(x) for x in args.sigma.split(",")]:
    hecmw.set_option("sell_sigma", sigma)
    for bps in [int(x) for x in args.bps.split(",")]:
        hecmw.set_option("spmv_blocks_per_sm", bps)
        for v_ in [int(x) for x in args.variants.split(",")]:
            hecmw.set_option("spmv_variant", v_)
            mean, mn = hecmw.matvec_bench(20, False)
            print("spmv variant %d bps %d sigma %d: mean %.4f ms min %.4f ms  -> %.0f GB/s algorithmic (%.3f GB)" % (v_, bps, sigma, mean, mn, nb / mean / 1e6, nb / 1e9))
hecmw.set_option("spmv_variant", 2); hecmw.set_option("spmv_blocks_per_sm", 8); hecmw.set_option("sell_sigma", 0)
if args.solve:
    hecmw.set_option("time_ax", 1)
    for rep in range(2):
        c0 = hecmw.launch_count()
        t0 = time.time()
        r = hecmw.hecmw_solve(mesh, A, device_resident=True)
        dt = time.time() - t0
        print("solve rep%d: err %d iter %d resid %.3e resid2 %.3e  wall %.3fs  setup %.3fs sol %.3fs  -> %.1f it/s  ax %.3fs launches %d"
              % (rep, r["error"], r["iter"], r["resid"], r["resid2"], dt, r["times"][0], r["times"][1], r["iter"] / r["times"][1], r["times"][3], hecmw.launch_count() - c0))
        A.Iarray[96] = 1
