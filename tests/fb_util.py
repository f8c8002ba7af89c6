"""Shared helpers for the parity tests: fixture loading, problem set-up with the oracle or the product."""
import json
import os

import numpy as np

from frontistr_b200 import meshgen
from frontistr_b200.hecmw_types import hecmwST_local_mesh, hecmwST_matrix, kel361IC
from frontistr_b200.mshio import apply_solver_card

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_meta():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def load_golden_mesh(name, elemopt361=kel361IC):
    allmeta = golden_meta()
    meta = allmeta.get(name)
    if meta is None:      # mesh-only fixture: take the cards of the (first) deck that names it as its mesh
        meta = next(v for v in allmeta.values() if v.get("mesh") == name)
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = hecmwST_local_mesh()
    m.node = d["node"]
    m.n_node = m.nn_internal = len(m.node) // 3
    m.elem_type_index = d["elem_type_index"]; m.elem_type_item = d["elem_type_item"]
    m.elem_node_index = d["elem_node_index"]; m.elem_node_item = d["elem_node_item"]
    m.n_elem = int(m.elem_type_index[-1]); m.n_elem_type = len(m.elem_type_item)
    m.section_ID = np.ones(m.n_elem, np.int32)
    m.sections = [dict(elemopt361=elemopt361, E=meta["E"], nu=meta["nu"])]
    m.global_node_ID = d["global_node_ID"]
    m.node_group = {k[5:]: d[k] for k in d.files if k.startswith("ngrp_")}
    cnt = dict(boundary=[tuple(b) for b in meta["boundary"]], cload=[tuple(c) for c in meta["cload"]],
               solver=(meta["solver"][0], meta["solver"][1]))
    return m, cnt, meta


def oracle_solve_deck(name):
    """Any linear golden deck (exA-exG: CLOAD, face pressure, volume loads, thermal load) through the ORACLE:
    K, load vector, Dirichlet elimination, CG.  Returns mesh, the solved matrix (X = displacement), meta and the
    temperature fields (TT, T0, alpha, reftemp) or None."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    mesh, cnt, meta = load_golden_mesh(name)
    A = orc.mat_con(mesh)
    orc.stiffmatrix(mesh, A)
    thermal = None
    if "thermal" in meta:
        th = meta["thermal"]
        thermal = (np.full(mesh.n_node, th["temperature"]), np.full(mesh.n_node, th["initial"]), th["alpha"], th["reftemp"])
        orc.ass_thermal_load(mesh, A, *thermal)
    elif "pressure" in meta:
        pr = meta["pressure"]
        orc.ass_dload_pressure(mesh, A, pr["elems"], pr["face"], pr["value"])
    elif "dload" in meta:
        orc.ass_dload_volume(mesh, A, meta["dload"]["ltype"], meta["dload"]["params"], meta["rho"])
    else:
        hecmw.fstr_ass_load_cload(mesh, A, cnt["cload"])
    n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
    for i in range(len(n)):
        orc.mat_ass_bc(A, int(n[i]), int(d[i]), float(v[i]))
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    orc.precond_clear()
    r = orc.solve(mesh, A)
    orc.precond_clear()
    assert r["error"] == 0 and A.Iarray[80] == 1
    return mesh, A, meta, thermal


def product_solve_deck(hecmw, name):
    """The same decks through the product's C-ABI (assembly, BC elimination and CG on the device; the right-hand side
    of the distributed-load decks comes from the restated load routines, which are the caller side of the path)."""
    from oracle import oracle as orc
    mesh, cnt, meta = load_golden_mesh(name)
    A = hecmw.hecmw_mat_con(mesh)
    A.fstr_mat_init().hecMAT_init()
    hecmw.mat_structure(mesh, A)
    hecmw.fstr_StiffMatrix(mesh, A, to_host=False)
    thermal = None
    if "thermal" in meta:
        th = meta["thermal"]
        thermal = (np.full(mesh.n_node, th["temperature"]), np.full(mesh.n_node, th["initial"]), th["alpha"], th["reftemp"])
        orc.ass_thermal_load(mesh, A, *thermal)
    elif "pressure" in meta:
        orc.ass_dload_pressure(mesh, A, meta["pressure"]["elems"], meta["pressure"]["face"], meta["pressure"]["value"])
    elif "dload" in meta:
        orc.ass_dload_volume(mesh, A, meta["dload"]["ltype"], meta["dload"]["params"], meta["rho"])
    else:
        hecmw.fstr_ass_load_cload(mesh, A, cnt["cload"])
    hecmw.mat_upload(A, values=False, rhs=True)
    n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
    hecmw.hecmw_mat_ass_bc(n, d, v)
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    r = hecmw.hecmw_solve(mesh, A, device_resident=True)
    assert r["error"] == 0 and A.Iarray[80] == 1
    return mesh, A, meta, thermal


STRESS_KEYS = ("11", "22", "33", "12", "23", "13")


def check_golden_stress(out, meta, digits_rel=1.0e-4):
    """The 26 strain / stress / Mises max/min lines of a golden log (nodal + elemental) against the dict returned by
    (oracle or product) fstr_NodalStress3D: the harness tolerance 1e-4 absolute AND the 5 printed digits."""
    g = meta["golden_stress"]
    worst = 0.0
    for where, E, S, M in (("node", "STRAIN", "STRESS", "MISES"), ("elem", "ESTRAIN", "ESTRESS", "EMISES")):
        for k, key in enumerate(STRESS_KEYS):
            for arr, tag in ((out[E][:, k], "E" + key), (out[S][:, k], "S" + key)):
                for got, want in ((arr.max(), g[where][tag][0]), (arr.min(), g[where][tag][1])):
                    assert abs(got - want) <= meta["tolerance_abs"] + digits_rel * abs(want), (where, tag, got, want)
                    if abs(want) > 1.0e-6:
                        worst = max(worst, abs(got - want) / abs(want))
        for got, want in ((out[M].max(), g[where]["SMS"][0]), (out[M].min(), g[where]["SMS"][1])):
            # (exF has stress-free nodes: min SMS ~ 1e-8 is the CG tolerance, not a digit to match)
            assert abs(got - want) <= digits_rel * abs(want) + 1e-6, (where, "SMS", got, want)
    return worst


def load_dist_fixture(name, rank):
    d = np.load(os.path.join(GOLDEN, "%s_%d.npz" % (name, rank)))
    m = hecmwST_local_mesh()
    for k in ("node", "node_ID", "global_node_ID", "elem_type_index", "elem_type_item", "elem_node_index",
              "elem_node_item", "neighbor_pe", "import_index", "import_item", "export_index", "export_item"):
        setattr(m, k, d[k])
    m.n_node = int(d["n_node"]); m.nn_internal = int(d["nn_internal"])
    m.n_elem = int(m.elem_type_index[-1]); m.n_elem_type = len(m.elem_type_item)
    m.n_neighbor_pe = len(m.neighbor_pe); m.my_rank = int(d["my_rank"]); m.PETOT = int(d["PETOT"])
    m.section_ID = np.ones(m.n_elem, np.int32)
    return m


from frontistr_b200.workloads import cube_cnt  # noqa: E402,F401  (the headline load case lives in the package)


def oracle_problem(mesh, cnt, apply_bc=True):
    """mat_con + fstr_StiffMatrix + CLOAD + fstr_AddBC with the ORACLE (checker only)."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    A = orc.mat_con(mesh)
    orc.stiffmatrix(mesh, A)
    hecmw.fstr_ass_load_cload(mesh, A, cnt["cload"])
    if apply_bc:
        n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
        for i in range(len(n)):
            orc.mat_ass_bc(A, int(n[i]), int(d[i]), float(v[i]))
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    return A


def product_problem(hecmw, mesh, cnt, to_host=True, apply_bc=True):
    """The same through the product's C-ABI (needs a GPU)."""
    A = hecmw.hecmw_mat_con(mesh)
    A.fstr_mat_init().hecMAT_init()
    hecmw.mat_structure(mesh, A)
    hecmw.fstr_StiffMatrix(mesh, A, to_host=False)
    hecmw.fstr_ass_load_cload(mesh, A, cnt["cload"])
    hecmw.mat_upload(A, values=False, rhs=True)
    if apply_bc:
        n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
        hecmw.hecmw_mat_ass_bc(n, d, v)
    if to_host:
        hecmw.mat_download(A, values=True, rhs=True)
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    return A


def oracle_iteration_spread(mesh, cnt, axes_list=("x", "y", "z", "xy", "yx", "zx", "zy", "xyz", "zyx", "xx", "xyx"), metis=True):
    """CG iteration counts of the reference algorithm (oracle, ONE thread: deterministic) on the same system decomposed
    into 2..8 subdomains by RCB along several axis strings and by KMETIS / PMETIS: DIAG-preconditioned CG is
    partition-invariant mathematically, so the spread of these counts is the reference's own sensitivity to the
    association order of its dot products (every decomposition sums the same terms in a different order)."""
    from oracle import oracle as orc
    from frontistr_b200 import partition
    nth = orc.num_threads()
    orc.set_num_threads(1)
    out = []
    try:
        decomps = [(partition.rcb_domains(mesh, axes), 1 << len(axes)) for axes in axes_list]
        if metis and partition.metis_available():
            decomps += [(partition.metis_domains(mesh, nd, meth)[0], nd) for nd in range(2, 9) for meth in ("KMETIS", "PMETIS")]
        for dom, nd in decomps:
            parts = partition.partition(mesh, "", domain=dom, nparts=nd)
            mats = [oracle_problem(p, cnt) for p in parts]
            orc.precond_clear()
            out.append(int(orc.solve(parts, mats)["iter"]))
    finally:
        orc.set_num_threads(nth)
        orc.precond_clear()
    return out


def oracle_history_spread(mesh, cnt, h_serial, n, axes_list=("x", "z", "xy", "xyz")):
    """Largest relative deviation, over the first n iterations, between the residual history of the reference algorithm
    (oracle, one thread) on the undivided system and on RCB decompositions of it: how far two executions of the SAME
    algorithm drift apart through the association order of the dot products alone."""
    from oracle import oracle as orc
    from frontistr_b200 import partition
    nth = orc.num_threads()
    orc.set_num_threads(1)
    worst = 0.0
    try:
        for axes in axes_list:
            parts = partition.partition(mesh, axes)
            orc.precond_clear()
            h = orc.solve(parts, [oracle_problem(p, cnt) for p in parts], hist_len=len(h_serial) + 8, ssor_threads=2)["hist"]
            m = min(n, len(h))
            worst = max(worst, float(np.max(np.abs(h[:m] / h_serial[:m] - 1.0))))
    finally:
        orc.set_num_threads(nth)
        orc.precond_clear()
    return worst


def distort(mesh, amp=0.15, seed=0):
    """Move interior nodes of a structured cube so elements are no longer congruent."""
    rng = np.random.default_rng(seed)
    xyz = mesh.node.reshape(-1, 3)
    lo, hi = xyz.min(0), xyz.max(0)
    interior = np.all((xyz > lo + 1e-12) & (xyz < hi - 1e-12), axis=1)
    # spacing from the first element edge
    c = mesh.elem_node_item[:2] - 1
    h = np.linalg.norm(xyz[c[1]] - xyz[c[0]])
    xyz[interior] += amp * h * (rng.random((int(interior.sum()), 3)) - 0.5)
    mesh.node = np.ascontiguousarray(xyz.reshape(-1))
    return mesh


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def simple_partition(mesh, nparts, axis=0):
    """Node-based partition with depth-1 overlap of a mesh into slabs along `axis`, producing the same
    kind of tables hecmw_part1 writes (internal nodes first, then external grouped by owner; elements
    touching an internal node; import/export tables ordered consistently).  Test helper only."""
    xyz = mesh.node.reshape(-1, 3)
    order = np.argsort(xyz[:, axis], kind="stable")
    owner = np.empty(mesh.n_node, np.int32)
    for p in range(nparts):
        owner[order[p * mesh.n_node // nparts:(p + 1) * mesh.n_node // nparts]] = p
    conn_all = mesh.elem_node_item
    eidx = mesh.elem_node_index
    parts = []
    et_of_elem = np.empty(mesh.n_elem, np.int32)
    for it in range(mesh.n_elem_type):
        et_of_elem[mesh.elem_type_index[it]:mesh.elem_type_index[it + 1]] = mesh.elem_type_item[it]
    for p in range(nparts):
        internal = np.nonzero(owner == p)[0]            # global 0-based, ascending
        keep = [e for e in range(mesh.n_elem) if np.any(owner[conn_all[eidx[e]:eidx[e + 1]] - 1] == p)]
        used = np.unique(np.concatenate([conn_all[eidx[e]:eidx[e + 1]] - 1 for e in keep]))
        ext = used[owner[used] != p]
        ext = ext[np.lexsort((ext, owner[ext]))]         # grouped by owner rank, ascending id inside
        loc_nodes = np.concatenate([internal, ext])
        g2l = -np.ones(mesh.n_node, np.int64); g2l[loc_nodes] = np.arange(len(loc_nodes)) + 1
        m = hecmwST_local_mesh()
        m.n_node = len(loc_nodes); m.nn_internal = len(internal)
        m.node = np.ascontiguousarray(xyz[loc_nodes].reshape(-1))
        m.global_node_ID = (loc_nodes + 1).astype(np.int32)
        types = sorted(set(int(et_of_elem[e]) for e in keep))
        tindex = [0]; nindex = [0]; items = []
        for et in types:
            for e in keep:
                if et_of_elem[e] != et:
                    continue
                c = g2l[conn_all[eidx[e]:eidx[e + 1]] - 1]
                items.append(c); nindex.append(nindex[-1] + len(c))
            tindex.append(len(nindex) - 1)
        m.n_elem = len(keep); m.n_elem_type = len(types)
        m.elem_type_index = np.array(tindex, np.int32); m.elem_type_item = np.array(types, np.int32)
        m.elem_node_index = np.array(nindex, np.int32); m.elem_node_item = np.concatenate(items).astype(np.int32)
        m.section_ID = np.ones(m.n_elem, np.int32)
        m.sections = mesh.sections
        m.my_rank = p; m.PETOT = nparts
        m.node_group = {k: g2l[np.asarray(v) - 1][g2l[np.asarray(v) - 1] > 0].astype(np.int32)
                        for k, v in mesh.node_group.items()}
        m._ext = ext; m._owner = owner; m._g2l = g2l
        parts.append(m)
    for p, m in enumerate(parts):
        nbrs = sorted(set(int(o) for o in m._owner[m._ext]))
        m.n_neighbor_pe = len(nbrs); m.neighbor_pe = np.array(nbrs, np.int32)
        imp_i = [0]; imp = []; exp_i = [0]; exp = []
        for q in nbrs:
            mine_from_q = m._ext[m._owner[m._ext] == q]            # ascending global id
            imp.extend(m._g2l[mine_from_q]); imp_i.append(len(imp))
            q_from_me = parts[q]._ext[parts[q]._owner[parts[q]._ext] == p]
            exp.extend(m._g2l[q_from_me]); exp_i.append(len(exp))
        m.import_index = np.array(imp_i, np.int32); m.import_item = np.array(imp, np.int32)
        m.export_index = np.array(exp_i, np.int32); m.export_item = np.array(exp, np.int32)
    return parts
