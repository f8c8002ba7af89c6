"""Runs the BASELINE.json configs other than the bench workload on ONE B200 and prints one JSON line per config
(to be archived under profiles/).  Not the bench: bench.py owns configs[1] (H150).

    python tools/run_configs.py --config 0     # tutorial/01 substitute: tet10 cube m=30 (226,981 nodes), CG + SSOR
    python tools/run_configs.py --config 2     # H200 on one GPU (the 2/4/8-GPU runs are `bench.py --hex 200` under torchrun)
    python tools/run_configs.py --config 3     # T107: tet10 107^3 x 6 (29.8 M DOF), block-Jacobi, assembly-dominated
    python tools/run_configs.py --config 4     # tutorial/03 hyperelastic cylinder, Newton + SSOR (needs FISTR_B200_RUN_NL_GPU=1)

tutorial/01_elastic_hinge/hinge.msh is absent from the reference mount, hence the substitute of similar size.
"""
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
from frontistr_b200.mshio import apply_solver_card  # noqa: E402


def linear_static(mesh, precond, tol, name, sigma=0, maxit=50000, ssor_fused=0):
    t0 = time.time()
    A = hecmw.hecmw_mat_con(mesh)
    t_con = time.time() - t0
    A.fstr_mat_init()
    hecmw.mat_structure(mesh, A)
    asm = [hecmw.fstr_StiffMatrix(mesh, A, to_host=False) for _ in range(3)]
    A.B = np.zeros(3 * A.NP); A.X = np.zeros(3 * A.NP)
    cnt = util.cube_cnt(tol=tol, precond=precond, maxit=maxit)
    hecmw.fstr_ass_load_cload(mesh, A, cnt["cload"])
    hecmw.mat_upload(A, values=False, rhs=True)
    n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
    hecmw.hecmw_mat_ass_bc(n, d, v)
    apply_solver_card(A, cnt["solver"]); A.Iarray[20] = 0; A.Iarray[21] = 0
    hecmw.set_option("time_ax", 1)
    hecmw.set_option("sell_sigma", sigma)
    hecmw.set_option("ssor_fused", ssor_fused)       # PRECOND 1 / 2 x 1 sweep: 2 ncolor + 3 launches per iteration (opt-in)
    c0 = hecmw.launch_count()
    out = []
    for rep in range(2):
        A.X[:] = 0.0; A.Iarray[96] = 1
        r = hecmw.hecmw_solve(mesh, A, device_resident=True)
        out.append(r)
    r = out[-1]
    nb = A.spmv_bytes()
    spmv_ms = 1e3 * float(r["times"][3]) / max(r["iter"], 1)
    return {"config": name, "nodes": mesh.n_node, "elements": mesh.n_elem, "dof": 3 * mesh.n_node, "precond": precond,
            "sell_sigma": sigma, "ssor_fused": ssor_fused, "launches_per_iteration": (hecmw.launch_count() - c0) / max(sum(x["iter"] for x in out), 1),
            "mat_con_s": t_con, "assembly_kernel_ms": asm[-1]["kernel_ms"],
            "assembly_elements_per_s": mesh.n_elem / (asm[-1]["kernel_ms"] * 1e-3), "error": r["error"], "iterations": r["iter"],
            "resid": r["resid"], "resid_true": r["resid2"], "solve_device_s": float(r["times"][4]),
            "cg_iterations_per_s": r["iter"] / float(r["times"][4]), "spmv_ms": spmv_ms,
            "spmv_algorithmic_gbs": nb / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None,
            "iteration_profile_us": [round(1e6 * float(x), 2) for x in r["profile"]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--sigma", type=int, default=0)
    ap.add_argument("--ssor-fused", type=int, default=0, help="config 0: the fused SSOR iteration (option ssor_fused, opt-in)")
    ap.add_argument("--small", action="store_true", help="reduced sizes (smoke run of this tool)")
    args = ap.parse_args()
    hecmw.init(0)
    if args.config == 0:
        res = linear_static(meshgen.tet10_cube(8 if args.small else 30), 1, 1e-8, "tutorial/01 substitute: tet10 cube, CG+SSOR",
                            args.sigma, ssor_fused=args.ssor_fused)
    elif args.config == 2:
        res = linear_static(meshgen.hex_cube(30 if args.small else 200), 3, 1e-8, "H200 on 1 GPU, CG+DIAG", args.sigma)
    elif args.config == 3:
        res = linear_static(meshgen.tet10_cube(10 if args.small else 107), 3, 1e-8, "T107 tet10, CG+DIAG (block-Jacobi)",
                            args.sigma)
    elif args.config == 4:
        if os.environ.get("FISTR_B200_RUN_NL_GPU") != "1":
            raise SystemExit("config 4 uses the nonlinear kernels, which have not run on hardware yet: set FISTR_B200_RUN_NL_GPU=1")
        from frontistr_b200.hecmw_types import kel361BBAR
        g = util.golden_meta()["cylinder"]
        mesh, _, _ = util.load_golden_mesh("cylinder", elemopt361=kel361BBAR)
        mesh.sections = [dict(elemopt361=kel361BBAR, E=1.0, nu=0.3, mtype=1, hyper=tuple(g["mooney_rivlin"]))]
        A = hecmw.hecmw_mat_con(mesh)
        A.fstr_mat_init()
        apply_solver_card(A, (g["solver"][0], g["solver"][1])); A.Iarray[20] = 0; A.Iarray[21] = 0
        bc = hecmw.boundary_dofs(mesh, [tuple(b) for b in g["boundary"]])
        t0 = time.time()
        r = hecmw.fstr_solve_NLGEOM(mesh, A, bc, ([], [], []), g["nlgeom"]["substeps"], g["nlgeom"]["converg"], g["nlgeom"]["max_iter"])
        res = {"config": "tutorial/03 hyperelastic cylinder (Mooney-Rivlin, B-bar TL, CG+SSOR)", "error": r["error"],
               "newton_iters": r["newton_iters"].tolist(), "cg_iterations": r["cg_iters"], "wall_s": time.time() - t0,
               "u3_min": float(r["unode"].reshape(-1, 3)[:, 2].min())}
    else:
        raise SystemExit("config must be 0, 2, 3 or 4 (1 is bench.py)")
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
