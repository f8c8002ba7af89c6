"""Development probe: where does the assembly kernel spend its time?  (asm_probe switches, see assembly.cuh)"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frontistr_b200 import hecmw, meshgen  # noqa: E402
from frontistr_b200.hecmw_types import kel361FI, kel361BBAR, kel361IC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--tet", type=int, default=0)
ap.add_argument("--probes", default="0,1,2,3")
args = ap.parse_args()
hecmw.init(0)
cases = [("tet10", lambda: meshgen.tet10_cube(args.tet))] if args.tet else \
        [("IC", lambda: meshgen.hex_cube(args.n)), ("FI", lambda: meshgen.hex_cube(args.n, elemopt361=kel361FI)),
         ("BBAR", lambda: meshgen.hex_cube(args.n, elemopt361=kel361BBAR))]
for name, mk in cases:
    mesh = mk()
    A = hecmw.hecmw_mat_con(mesh); A.fstr_mat_init()
    hecmw.mat_structure(mesh, A)
    for icv in ((0, 1, 2) if name == "IC" else (-1,)):      # IC: the three condensation kernels (element_math.cuh)
        hecmw.set_option("ic_inverse", icv)
        for probe in [int(x) for x in args.probes.split(",")]:
            hecmw.set_option("asm_probe", probe)
            ts = [hecmw.fstr_StiffMatrix(mesh, A, to_host=False)["kernel_ms"] for _ in range(3)]
            print("%-5s%s n_elem=%d probe=%d: kernel %.3f ms -> %.1f M elem/s" % (name, "" if icv < 0 else " ic_inverse=%d" % icv, mesh.n_elem, probe, min(ts), mesh.n_elem / min(ts) / 1e3), flush=True)
    hecmw.set_option("asm_probe", 0); hecmw.set_option("ic_inverse", -1)
