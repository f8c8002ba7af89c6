"""In-process node-based domain decomposition producing what hecmw_part1 writes per rank
(hecmw1/tools/partitioner/hecmw_partition.c; on-disk form hecmw1/src/common/hecmw_io_dist.c:1888-2284):

* internal nodes first in ascending global id, then the external (halo) nodes in ascending global id;
* the elements that touch at least one internal node, in global order inside each element-type block;
* neighbour ranks ascending; per neighbour the import and export node lists in ascending global id;
* node_ID = (local id in the owning rank, owning rank).

The node->domain map comes from the library's own RCB (fistr_b200_part_rcb: same bisection contract as the
reference, ties on a coordinate broken by node index) or from any other source (`domain=`).  Given the same map,
equality of every table with the real hecmw_part1 is pinned by tests/test_partition.py against fixtures generated
with oracle/_ref/hecmw_part1.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, dp, ip
from .hecmw_types import hecmwST_local_mesh

_AX = {"x": 0, "y": 1, "z": 2}
_NN = {341: 4, 342: 10, 351: 6, 352: 15, 361: 8, 362: 20}


def rcb_domains(mesh: hecmwST_local_mesh, axes: str) -> np.ndarray:
    """axes: one letter per bisection level, e.g. 'xyz' for 8 domains (the RCB line of hecmw_part_ctrl.dat)."""
    ax = np.array([_AX[a] for a in axes.lower().replace(",", "")], np.int32)
    dom = np.zeros(mesh.n_node, np.int32)
    node = np.ascontiguousarray(mesh.node, np.float64)
    check(_lib.lib().fistr_b200_part_rcb(mesh.n_node, dp(node), len(ax), ip(ax), ip(dom)))
    return dom


def metis_available() -> bool:
    return bool(_lib.lib().fistr_b200_part_metis_available())


def metis_domains(mesh: hecmwST_local_mesh, n_domain: int, method: str = "KMETIS"):
    """`!PARTITION,TYPE=NODE-BASED,METHOD=KMETIS|PMETIS,DOMAIN=n` of hecmw_part_ctrl.dat: the reference's node graph of
    element edges partitioned by METIS 5 (csrc/partition_metis.hpp).  Returns (domain per node, edge cut)."""
    import ctypes as C
    meth = {"PMETIS": 1, "KMETIS": 2}[method.upper()]
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    eti, ett = i32(mesh.elem_type_index), i32(mesh.elem_type_item)
    eni, enn = i32(mesh.elem_node_index), i32(mesh.elem_node_item)
    dom = np.zeros(mesh.n_node, np.int32)
    cut = C.c_longlong(0)
    check(_lib.lib().fistr_b200_part_metis(meth, int(n_domain), mesh.n_node, mesh.n_elem_type, ip(eti), ip(ett), ip(eni), ip(enn),
                                           ip(dom), C.byref(cut)))
    return dom, int(cut.value)


def domains_from_control(mesh: hecmwST_local_mesh, ctrl_text: str):
    """The node->domain map hecmw_part1 would compute for this hecmw_part_ctrl.dat content
    (`!PARTITION,TYPE=NODE-BASED,METHOD={RCB|KMETIS|PMETIS},DOMAIN=n[,DEPTH=1][,UCD=...]`, RCB followed by a line of axes
    `x,y,z`; parser hecmw1/tools/partitioner/hecmw_part_get_control.c:76-460).  Returns (domain, n_domain).
    Node-based, depth 1 only - what the solver path uses."""
    lines = [ln.strip() for ln in ctrl_text.splitlines() if ln.strip() and not ln.strip().startswith(("#", "!!"))]
    head = next((ln for ln in lines if ln.upper().startswith("!PARTITION")), None)
    if head is None:
        raise ValueError("hecmw_part_ctrl.dat: no !PARTITION line")
    opts = {}
    for tok in head.split(",")[1:]:
        k, _, v = tok.partition("=")
        opts[k.strip().upper()] = v.strip()
    if opts.get("TYPE", "NODE-BASED").upper() != "NODE-BASED":
        raise ValueError("only TYPE=NODE-BASED is on this path (got %s)" % opts["TYPE"])
    if int(opts.get("DEPTH", "1")) != 1:
        raise ValueError("only DEPTH=1 is on this path")
    nd = int(opts["DOMAIN"])
    method = opts.get("METHOD", "").upper()
    if method == "RCB":
        rest = lines[lines.index(head) + 1:]
        axes = "".join(rest[0].replace(",", " ").split()).lower() if rest else ""
        if nd != 1 << len(axes):
            raise ValueError("RCB: DOMAIN=%d needs %d axes (hecmw_part_get_control.c:271-340)" % (nd, max(nd.bit_length() - 1, 0)))
        return rcb_domains(mesh, axes), nd
    if method in ("KMETIS", "PMETIS"):
        return metis_domains(mesh, nd, method)[0], nd
    raise ValueError("METHOD must be RCB, KMETIS or PMETIS")


def local_mesh(mesh: hecmwST_local_mesh, domain: np.ndarray, rank: int, nparts: int) -> hecmwST_local_mesh:
    """The distributed mesh of `rank` (what hecmw_part1 writes to <name>.<rank>), built by the library's native
    host code (csrc/partition.hpp: LocalMesh) from the node->domain map."""
    L = _lib.lib()
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    eti, ett = i32(mesh.elem_type_index), i32(mesh.elem_type_item)
    eni, enn = i32(mesh.elem_node_index), i32(mesh.elem_node_item)
    dom = i32(domain)
    sizes = np.zeros(8, np.int32)
    check(L.fistr_b200_part_local_count(mesh.n_node, mesh.n_elem_type, ip(eti), ip(ett), ip(eni), ip(enn), ip(dom),
                                        int(rank), int(nparts), ip(sizes)))
    n_node, nn_int, n_elem, n_type, n_conn, n_nb, n_imp, n_exp = [int(v) for v in sizes]
    z = lambda n: np.zeros(max(n, 1), np.int32)
    gnode, node_ID = z(n_node), z(2 * n_node)
    l_eti, l_ett, l_eni, l_enn, gelem = z(n_type + 1), z(n_type), z(n_elem + 1), z(n_conn), z(n_elem)
    nb, imp_i, imp, exp_i, exp = z(n_nb), z(n_nb + 1), z(n_imp), z(n_nb + 1), z(n_exp)
    check(L.fistr_b200_part_local_fill(ip(gnode), ip(node_ID), ip(l_eti), ip(l_ett), ip(l_eni), ip(l_enn), ip(gelem),
                                       ip(nb), ip(imp_i), ip(imp), ip(exp_i), ip(exp)))
    gnode = gnode[:n_node].astype(np.int64)
    m = hecmwST_local_mesh()
    m.n_node, m.nn_internal, m.n_elem, m.n_elem_type = n_node, nn_int, n_elem, n_type
    m.node = np.ascontiguousarray(mesh.node.reshape(-1, 3)[gnode].reshape(-1))
    m.global_node_ID = mesh.global_node_ID[gnode].astype(np.int32)
    m.node_ID = node_ID[: 2 * n_node].reshape(-1, 2).copy()
    m.elem_type_index, m.elem_type_item = l_eti[: n_type + 1].copy(), l_ett[:n_type].copy()
    m.elem_node_index, m.elem_node_item = l_eni[: n_elem + 1].copy(), l_enn[:n_conn].copy()
    kg = gelem[:n_elem].astype(np.int64)
    m.section_ID = (mesh.section_ID[kg] if mesh.section_ID is not None else np.ones(len(kg))).astype(np.int32)
    m.global_elem_ID = (kg + 1).astype(np.int32)
    m.sections = mesh.sections
    m.my_rank = rank; m.PETOT = nparts; m.zero = 1 if rank == 0 else 0
    g2l = np.zeros(mesh.n_node, np.int32)
    g2l[gnode] = np.arange(1, n_node + 1, dtype=np.int32)
    m.node_group = {}
    for k, v in mesh.node_group.items():
        l = g2l[np.asarray(v, np.int64) - 1]
        m.node_group[k] = l[l > 0].astype(np.int32)
    m.n_neighbor_pe = n_nb
    m.neighbor_pe = nb[:n_nb].copy()
    m.import_index, m.import_item = imp_i[: n_nb + 1].copy(), imp[:n_imp].copy()
    m.export_index, m.export_item = exp_i[: n_nb + 1].copy(), exp[:n_exp].copy()
    return m


def _local_mesh_numpy(mesh: hecmwST_local_mesh, domain: np.ndarray, rank: int, nparts: int) -> hecmwST_local_mesh:
    """Independent numpy formulation of local_mesh (kept as a cross-check of the native builder)."""
    owner = domain
    n_glob = mesh.n_node
    # owner-local numbering of every node: position among its owner's nodes in ascending global id
    order = np.lexsort((np.arange(n_glob), owner))
    local_in_owner = np.empty(n_glob, np.int64)
    counts = np.bincount(owner, minlength=nparts)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    local_in_owner[order] = np.arange(n_glob) - starts[owner[order]] + 1
    internal = np.nonzero(owner == rank)[0]
    m = hecmwST_local_mesh()
    tindex = [0]; nindex_parts = []; items = []; types = []; kept_global = []
    for it in range(mesh.n_elem_type):
        et = int(mesh.elem_type_item[it]); nn = _NN[et]
        e0, e1 = int(mesh.elem_type_index[it]), int(mesh.elem_type_index[it + 1])
        s0 = int(mesh.elem_node_index[e0])
        conn = mesh.elem_node_item[s0: s0 + (e1 - e0) * nn].reshape(-1, nn)
        keep = np.nonzero((owner[conn - 1] == rank).any(axis=1))[0]
        if len(keep) == 0:
            continue
        types.append(et); items.append(conn[keep]); kept_global.append(keep + e0)
        tindex.append(tindex[-1] + len(keep))
    used = np.unique(np.concatenate([c.reshape(-1) for c in items])) - 1 if items else np.zeros(0, np.int64)
    ext = used[owner[used] != rank]                     # ascending global id
    loc_nodes = np.concatenate([internal, ext])
    g2l = np.zeros(n_glob, np.int32)
    g2l[loc_nodes] = np.arange(1, len(loc_nodes) + 1, dtype=np.int32)
    m.n_node = len(loc_nodes); m.nn_internal = len(internal)
    m.node = np.ascontiguousarray(mesh.node.reshape(-1, 3)[loc_nodes].reshape(-1))
    m.global_node_ID = mesh.global_node_ID[loc_nodes].astype(np.int32)
    m.node_ID = np.stack([local_in_owner[loc_nodes], owner[loc_nodes]], axis=1).astype(np.int32)
    m.n_elem = tindex[-1]; m.n_elem_type = len(types)
    m.elem_type_index = np.array(tindex, np.int32); m.elem_type_item = np.array(types, np.int32)
    sizes = np.concatenate([np.full(c.shape[0], c.shape[1], np.int64) for c in items]) if items else np.zeros(0, np.int64)
    m.elem_node_index = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    m.elem_node_item = np.concatenate([g2l[c - 1].reshape(-1) for c in items]).astype(np.int32) if items else np.zeros(0, np.int32)
    kg = np.concatenate(kept_global) if kept_global else np.zeros(0, np.int64)
    m.section_ID = (mesh.section_ID[kg] if mesh.section_ID is not None else np.ones(len(kg))).astype(np.int32)
    m.global_elem_ID = (kg + 1).astype(np.int32)
    m.sections = mesh.sections
    m.my_rank = rank; m.PETOT = nparts; m.zero = 1 if rank == 0 else 0
    m.node_group = {}
    for k, v in mesh.node_group.items():
        l = g2l[np.asarray(v, np.int64) - 1]
        m.node_group[k] = l[l > 0].astype(np.int32)
    # communication tables
    ext_owner = owner[ext]
    nbrs = np.unique(ext_owner)
    # exports: my internal nodes that appear as external nodes of rank q  <=> nodes of elements that
    # touch an internal node of q.  Computed symmetrically from q's point of view.
    imp_index = [0]; imp = []; exp_index = [0]; exp = []
    exp_lists = _export_lists(mesh, owner, rank, nparts)
    nbr_all = sorted(set(nbrs.tolist()) | set(exp_lists.keys()))
    for q in nbr_all:
        mine = ext[ext_owner == q]
        imp.append(g2l[mine]); imp_index.append(imp_index[-1] + len(mine))
        ex = exp_lists.get(q, np.zeros(0, np.int64))
        exp.append(g2l[ex]); exp_index.append(exp_index[-1] + len(ex))
    m.n_neighbor_pe = len(nbr_all)
    m.neighbor_pe = np.array(nbr_all, np.int32)
    m.import_index = np.array(imp_index, np.int32); m.export_index = np.array(exp_index, np.int32)
    m.import_item = np.concatenate(imp).astype(np.int32) if imp else np.zeros(0, np.int32)
    m.export_item = np.concatenate(exp).astype(np.int32) if exp else np.zeros(0, np.int32)
    return m


def _export_lists(mesh, owner, rank, nparts):
    """For every other rank q: my internal nodes (global 0-based, ascending) that q holds as external,
    i.e. nodes owned by `rank` in elements that have a node owned by q."""
    out = {}
    cand = []
    for it in range(mesh.n_elem_type):
        nn = _NN[int(mesh.elem_type_item[it])]
        e0, e1 = int(mesh.elem_type_index[it]), int(mesh.elem_type_index[it + 1])
        s0 = int(mesh.elem_node_index[e0])
        conn = mesh.elem_node_item[s0: s0 + (e1 - e0) * nn].reshape(-1, nn) - 1
        ow = owner[conn]
        mixed = (ow == rank).any(axis=1) & (ow != rank).any(axis=1)
        c = conn[mixed]; o = ow[mixed]
        if len(c) == 0:
            continue
        # all (my node, other owner) pairs inside each mixed element
        for a in range(nn):
            mine = o[:, a] == rank
            if not mine.any():
                continue
            for b in range(nn):
                sel = mine & (o[:, b] != rank)
                if sel.any():
                    cand.append(np.stack([o[sel, b].astype(np.int64), c[sel, a].astype(np.int64)], axis=1))
    if not cand:
        return out
    pairs = np.unique(np.concatenate(cand), axis=0)     # sorted by (q, node)
    for q in np.unique(pairs[:, 0]):
        out[int(q)] = pairs[pairs[:, 0] == q, 1]
    return out


def partition(mesh: hecmwST_local_mesh, axes: str, domain=None, nparts=None):
    """All local meshes of an RCB split along `axes`, or of any node->domain map given as `domain` (+ nparts)."""
    dom = rcb_domains(mesh, axes) if domain is None else np.ascontiguousarray(domain, np.int32)
    if nparts is None:
        nparts = 1 << len(axes.replace(",", ""))
    return [local_mesh(mesh, dom, r, nparts) for r in range(nparts)]
