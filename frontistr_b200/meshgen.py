"""Synthetic structured cube meshes in hecmwST_local_mesh form, plus a .msh writer.

* ``hex_cube(n)``    — n^3 C3D8 (etype 361) cells on [0,L]^3, node id = 1 + i + (n+1) j + (n+1)^2 k,
                       element node order 1-4 bottom CCW, 5-8 top (lib/element/hex8n.f90).
* ``tet10_cube(m)``  — m^3 cells, each split into 6 tets (Kuhn), mid-edge nodes on the (2m+1)^3 lattice;
                       connectivity is stored in FrontISTR-internal order (vertices, m12, m23, m31, m14,
                       m24, m34 — i.e. after fistr1/src/common/hecmw2fstr_connect_conv.c:19 was applied).

Node groups: ``FIX`` = face x=0, ``CL1`` = face x=L (mirrors examples/static/exA).
"""
from __future__ import annotations

import numpy as np

from .hecmw_types import hecmwST_local_mesh, kel361IC

# internal(FrontISTR) position -> file(HEC-MW) position for etype 342,
# fistr1/src/common/hecmw2fstr_connect_conv.c:19  (internal[j] = file[T[j]-1])
TABLE342 = np.array([1, 2, 3, 4, 7, 5, 6, 8, 9, 10]) - 1


def hex_cube(n: int, L: float = 1.0, E: float = 210000.0, nu: float = 0.3,
             elemopt361: int = kel361IC, nx=None, ny=None, nz=None) -> hecmwST_local_mesh:
    nx = nx or n; ny = ny or n; nz = nz or n
    px, py, pz = nx + 1, ny + 1, nz + 1
    k, j, i = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
    node = np.empty((px * py * pz, 3))
    node[:, 0] = i.reshape(-1) * (L / n)
    node[:, 1] = j.reshape(-1) * (L / n)
    node[:, 2] = k.reshape(-1) * (L / n)
    ke, je, ie = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    n0 = (1 + ie + px * je + px * py * ke).reshape(-1).astype(np.int64)
    conn = np.stack([n0, n0 + 1, n0 + 1 + px, n0 + px,
                     n0 + px * py, n0 + 1 + px * py, n0 + 1 + px + px * py, n0 + px + px * py], axis=1)
    ne = conn.shape[0]
    m = hecmwST_local_mesh()
    m.n_node = m.nn_internal = node.shape[0]
    m.n_elem = ne
    m.node = np.ascontiguousarray(node.reshape(-1))
    m.n_elem_type = 1
    m.elem_type_index = np.array([0, ne], np.int32)
    m.elem_type_item = np.array([361], np.int32)
    m.elem_node_index = (8 * np.arange(ne + 1)).astype(np.int32)
    m.elem_node_item = np.ascontiguousarray(conn.reshape(-1).astype(np.int32))
    m.section_ID = np.ones(ne, np.int32)
    m.sections = [dict(elemopt361=elemopt361, E=E, nu=nu)]
    m.global_node_ID = np.arange(1, m.n_node + 1, dtype=np.int32)
    ids = np.arange(1, m.n_node + 1, dtype=np.int32)
    m.node_group = {"FIX": ids[i.reshape(-1) == 0], "CL1": ids[i.reshape(-1) == nx]}
    return m


# Kuhn split of the unit cube into 6 tets, vertices as (dx,dy,dz) corner offsets; positive orientation
_KUHN = [((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)),
         ((0, 0, 0), (1, 1, 0), (0, 1, 0), (1, 1, 1)),
         ((0, 0, 0), (0, 1, 0), (0, 1, 1), (1, 1, 1)),
         ((0, 0, 0), (0, 1, 1), (0, 0, 1), (1, 1, 1)),
         ((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1)),
         ((0, 0, 0), (1, 0, 1), (1, 0, 0), (1, 1, 1))]
_EDGES = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]   # m12, m23, m31, m14, m24, m34


def tet10_cube(m_cells: int, L: float = 1.0, E: float = 210000.0, nu: float = 0.3) -> hecmwST_local_mesh:
    mc = m_cells
    p = 2 * mc + 1
    k, j, i = np.meshgrid(np.arange(p), np.arange(p), np.arange(p), indexing="ij")
    h = L / (2 * mc)
    node = np.stack([i.reshape(-1) * h, j.reshape(-1) * h, k.reshape(-1) * h], axis=1)
    ke, je, ie = np.meshgrid(np.arange(mc), np.arange(mc), np.arange(mc), indexing="ij")
    ie = ie.reshape(-1); je = je.reshape(-1); ke = ke.reshape(-1)

    def nid(ii, jj, kk):
        return 1 + ii + p * jj + p * p * kk

    conns = []
    for tet in _KUHN:
        v = [(2 * (ie + d[0]), 2 * (je + d[1]), 2 * (ke + d[2])) for d in tet]
        cols = [nid(*vv) for vv in v]
        for a, b in _EDGES:
            cols.append(nid((v[a][0] + v[b][0]) // 2, (v[a][1] + v[b][1]) // 2, (v[a][2] + v[b][2]) // 2))
        conns.append(np.stack(cols, axis=1))
    conn = np.stack(conns, axis=1).reshape(-1, 10)       # cell-major, 6 tets per cell
    ne = conn.shape[0]
    msh = hecmwST_local_mesh()
    msh.n_node = msh.nn_internal = node.shape[0]
    msh.n_elem = ne
    msh.node = np.ascontiguousarray(node.reshape(-1))
    msh.n_elem_type = 1
    msh.elem_type_index = np.array([0, ne], np.int32)
    msh.elem_type_item = np.array([342], np.int32)
    msh.elem_node_index = (10 * np.arange(ne + 1)).astype(np.int32)
    msh.elem_node_item = np.ascontiguousarray(conn.reshape(-1).astype(np.int32))
    msh.section_ID = np.ones(ne, np.int32)
    msh.sections = [dict(elemopt361=kel361IC, E=E, nu=nu)]
    msh.global_node_ID = np.arange(1, msh.n_node + 1, dtype=np.int32)
    ids = np.arange(1, msh.n_node + 1, dtype=np.int32)
    msh.node_group = {"FIX": ids[i.reshape(-1) == 0], "CL1": ids[i.reshape(-1) == p - 1]}
    return msh


def write_msh(mesh: hecmwST_local_mesh, path: str, title: str = "synthetic cube") -> None:
    """HEC-MW entire-mesh text (.msh), SURVEY Appendix A.2 — for feeding hecmw_part1."""
    xyz = mesh.node.reshape(-1, 3)
    with open(path, "w") as f:
        f.write("!HEADER\n %s\n!NODE\n" % title)
        for nidx in range(mesh.n_node):
            f.write("%d, %.15g, %.15g, %.15g\n" % (mesh.global_node_ID[nidx], *xyz[nidx]))
        for it in range(mesh.n_elem_type):
            et = int(mesh.elem_type_item[it])
            f.write("!ELEMENT, TYPE=%d\n" % et)
            for e in range(mesh.elem_type_index[it], mesh.elem_type_index[it + 1]):
                c = mesh.elem_node_item[mesh.elem_node_index[e]: mesh.elem_node_index[e + 1]]
                c = mesh.global_node_ID[c - 1]
                if et == 342:       # internal -> file order
                    cf = np.empty(10, c.dtype); cf[TABLE342] = c; c = cf
                f.write("%d, %s\n" % (e + 1, ", ".join(str(int(v)) for v in c)))
        s = mesh.sections[0]
        f.write("!SECTION, TYPE=SOLID, EGRP=ALL, MATERIAL=M1\n 1.0\n")
        f.write("!MATERIAL, NAME=M1, ITEM=1\n!ITEM=1, SUBITEM=2\n %.15g, %.15g\n" % (s["E"], s["nu"]))
        for name, ids in mesh.node_group.items():
            f.write("!NGROUP, NGRP=%s\n" % name)
            g = mesh.global_node_ID[np.asarray(ids) - 1]
            for a in range(0, len(g), 8):
                f.write(" " + ", ".join(str(int(v)) for v in g[a:a + 8]) + "\n")
        f.write("!END\n")
