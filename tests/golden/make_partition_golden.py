"""Generates tests/golden/part_<case>.npz with the REAL reference partitioner (oracle/_ref/hecmw_part1,
compiled from the reference sources by oracle/build_ref.sh): synthetic cubes written as .msh, split with
METHOD=RCB, and the per-rank HECMW-DMD-ASCII files parsed back.  Run in the build container only."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from frontistr_b200 import meshgen, mshio  # noqa: E402

PART1 = os.path.join(ROOT, "oracle", "_ref", "hecmw_part1")
CASES = {"hex5_xy": (lambda: meshgen.hex_cube(5), "xy"), "hex6_xyz": (lambda: meshgen.hex_cube(6), "xyz"),
         "hex7_zyx": (lambda: meshgen.hex_cube(7), "zyx"), "tet2_xy": (lambda: meshgen.tet10_cube(2), "xy"),
         "tet3_xyz": (lambda: meshgen.tet10_cube(3), "xyz")}
KEYS = ("global_node_ID", "node_ID", "elem_type_index", "elem_type_item", "elem_node_index", "elem_node_item",
        "neighbor_pe", "import_index", "import_item", "export_index", "export_item")


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name, (maker, axes) in CASES.items():
        mesh = maker()
        nd = 1 << len(axes)
        with tempfile.TemporaryDirectory() as tmp:
            meshgen.write_msh(mesh, os.path.join(tmp, "cube.msh"))
            open(os.path.join(tmp, "hecmw_ctrl.dat"), "w").write(
                "!MESH, NAME=part_in, TYPE=HECMW-ENTIRE\n cube.msh\n!MESH, NAME=part_out, TYPE=HECMW-DIST\n cube_p\n")
            open(os.path.join(tmp, "hecmw_part_ctrl.dat"), "w").write(
                "!PARTITION,TYPE=NODE-BASED,METHOD=RCB,DOMAIN=%d\n %s\n" % (nd, ",".join(axes)))
            subprocess.check_call([PART1], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            data = {"axes": axes, "ndomain": nd}
            for r in range(nd):
                with open(os.path.join(tmp, "cube_p.%d" % r), "rb") as f:     # the whole HECMW-DMD-ASCII file, for the writer test
                    data["r%d_sha256" % r] = hashlib.sha256(f.read()).hexdigest()
                d = mshio.read_dist_mesh(os.path.join(tmp, "cube_p.%d" % r))
                data["r%d_n_node" % r] = d.n_node
                data["r%d_nn_internal" % r] = d.nn_internal
                for k in KEYS:
                    data["r%d_%s" % (r, k)] = getattr(d, k)
        np.savez_compressed(os.path.join(out, "part_%s.npz" % name), **data)
        print("wrote part_%s.npz" % name)


if __name__ == "__main__":
    main()
