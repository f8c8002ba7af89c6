"""Development probe: device time of the strain / stress recovery (update_3d + nodal_stress_3d) per element
formulation on a cube, next to the assembly kernel of the same mesh."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frontistr_b200 import hecmw, meshgen  # noqa: E402
from frontistr_b200.hecmw_types import kel361FI, kel361BBAR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--tet", type=int, default=40)
args = ap.parse_args()
hecmw.init(0)
cases = [("IC", lambda: meshgen.hex_cube(args.n)), ("FI", lambda: meshgen.hex_cube(args.n, elemopt361=kel361FI)),
         ("BBAR", lambda: meshgen.hex_cube(args.n, elemopt361=kel361BBAR)), ("tet10", lambda: meshgen.tet10_cube(args.tet))]
for name, mk in cases:
    mesh = mk()
    A = hecmw.hecmw_mat_con(mesh); A.fstr_mat_init()
    hecmw.mat_structure(mesh, A)
    asm = min(hecmw.fstr_StiffMatrix(mesh, A, to_host=False)["kernel_ms"] for _ in range(3))
    u = 1e-3 * np.sin(np.arange(3 * mesh.n_node, dtype=float))
    upd, nod, wall = [], [], []
    for _ in range(3):
        t0 = time.time()
        hecmw.fstr_UpdateNewton_linear(mesh, None, u)
        hecmw.fstr_NodalStress3D(mesh)
        wall.append(1e3 * (time.time() - t0))
        upd.append(hecmw.get_option("update_kernel_us") / 1e3); nod.append(hecmw.get_option("nodal_stress_kernel_us") / 1e3)
    print("%-5s n_elem=%d: assembly %.3f ms | update %.3f ms (%.1f M elem/s) | nodal stress %.3f ms | both calls incl. copies %.1f ms"
          % (name, mesh.n_elem, asm, min(upd), mesh.n_elem / min(upd) / 1e3, min(nod), min(wall)), flush=True)
