"""Development probe for the two paths written last in round 1 (never run on hardware): wall time of hecmw_mat_con on the
host vs on the device (option matcon_device), and CG + BILU(0) (PRECOND=10) vs CG + SSOR / DIAG on a hex cube: iterations,
solver time, number of levels.  Numbers printed here are development numbers, not bench values."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import fb_util as util  # noqa: E402
from frontistr_b200 import hecmw, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--solve-n", type=int, default=60)
args = ap.parse_args()
hecmw.init(0)
mesh = meshgen.hex_cube(args.n)
ref = None
for dev in (0, 1, 1):
    hecmw.set_option("matcon_device", dev)
    t0 = time.perf_counter()
    A = hecmw.hecmw_mat_con(mesh)
    dt = time.perf_counter() - t0
    if ref is None:
        ref = A
    same = all(np.array_equal(getattr(A, k), getattr(ref, k)) for k in ("indexL", "itemL", "indexU", "itemU"))
    print("mat_con %s: n_node=%d NPL=%d  %.3f s  identical=%s" % ("device" if dev else "host  ", mesh.n_node, A.NPL, dt, same), flush=True)
hecmw.set_option("matcon_device", 0)
mesh = meshgen.hex_cube(args.solve_n)
for pc in (3, 1, 10):
    cnt = util.cube_cnt(tol=1e-8, precond=pc)
    A = util.product_problem(hecmw, mesh, cnt, to_host=False)
    r = hecmw.hecmw_solve(mesh, A, device_resident=True)
    print("CG PRECOND=%2d hex%d: error=%d iter=%d  set-up %.3f s  solver %.3f s%s" % (
        pc, args.solve_n, r["error"], r["iter"], r["times"][0], r["times"][1],
        ("  levels=%d launches/apply=%d" % (hecmw.get_option("bilu_levels"), hecmw.get_option("bilu_launches_per_apply"))) if pc == 10 else ""), flush=True)
