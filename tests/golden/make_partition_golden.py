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


METIS_CASES = {"hex6_k8": (lambda: meshgen.hex_cube(6), "KMETIS", 8), "hex5_p3": (lambda: meshgen.hex_cube(5), "PMETIS", 3),
               "tet3_k4": (lambda: meshgen.tet10_cube(3), "KMETIS", 4), "tet2_p2": (lambda: meshgen.tet10_cube(2), "PMETIS", 2),
               "hexrag_k5": (lambda: meshgen.hex_cube(3, nx=7, ny=4, nz=5), "KMETIS", 5)}


def main_metis():
    """part_metis_<case>.npz: node -> domain map and the per-rank files' sha256 of the REAL hecmw_part1 built with METIS
    (oracle/_ref/hecmw_part1_metis: the reference's partitioner + the CUDA toolkit's METIS 5 archive through the 32-bit
    front of oracle/metis_shim)."""
    out = os.path.dirname(os.path.abspath(__file__))
    part = os.path.join(ROOT, "oracle", "_ref", "hecmw_part1_metis")
    for name, (maker, method, nd) in METIS_CASES.items():
        mesh = maker()
        with tempfile.TemporaryDirectory() as tmp:
            meshgen.write_msh(mesh, os.path.join(tmp, "cube.msh"))
            open(os.path.join(tmp, "hecmw_ctrl.dat"), "w").write(
                "!MESH, NAME=part_in, TYPE=HECMW-ENTIRE\n cube.msh\n!MESH, NAME=part_out, TYPE=HECMW-DIST\n cube_p\n")
            open(os.path.join(tmp, "hecmw_part_ctrl.dat"), "w").write("!PARTITION,TYPE=NODE-BASED,METHOD=%s,DOMAIN=%d\n" % (method, nd))
            subprocess.check_call([part], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dom = np.full(mesh.n_node, -1, np.int32)
            data = {"method": method, "ndomain": nd}
            for r in range(nd):
                with open(os.path.join(tmp, "cube_p.%d" % r), "rb") as f:
                    data["r%d_sha256" % r] = hashlib.sha256(f.read()).hexdigest()
                d = mshio.read_dist_mesh(os.path.join(tmp, "cube_p.%d" % r))
                dom[d.global_node_ID[: d.nn_internal] - 1] = r
                for k in KEYS:
                    data["r%d_%s" % (r, k)] = getattr(d, k)
            assert dom.min() >= 0
            data["domain"] = dom
        np.savez_compressed(os.path.join(out, "part_metis_%s.npz" % name), **data)
        print("wrote part_metis_%s.npz" % name)


if __name__ == "__main__":
    if "--metis" in sys.argv:
        main_metis()
    else:
        main()
