"""Minimal readers for the data formats either side of the hot path.

* ``read_msh``  — HEC-MW entire mesh (.msh): !NODE, !ELEMENT (361/342), !SECTION, !MATERIAL,
                  !NGROUP (list and GENERATE).  Mirrors what hecmw_get_mesh + hecmw2fstr_mesh_conv
                  leave in hecmwST_local_mesh: unused nodes dropped and local ids assigned in
                  ascending global-id order (hecmw1/src/common/hecmw_io_mesh.c:3039,3895), 342
                  connectivity permuted to FrontISTR order (fistr1/src/common/hecmw2fstr_connect_conv.c:19).
* ``read_cnt``  — the analysis control file: !BOUNDARY, !CLOAD, !SOLVER (fistr1/src/common/
                  fstr_ctrl_common.f90:69-169, fstr_setup.f90:1473-1509).
* ``read_dist_mesh`` — HECMW-DMD-ASCII (hecmw1/src/common/hecmw_io_dist.c:1888-2284) written by hecmw_part1.

This is host-side plumbing for tests/bench; the reference's own readers stay in place in a real build.
"""
from __future__ import annotations

import os
import re
import warnings

import numpy as np

from .hecmw_types import hecmwST_local_mesh, hecmwST_matrix, kel361IC
from .meshgen import TABLE342

_NN = {341: 4, 342: 10, 351: 6, 352: 15, 361: 8, 362: 20}


def _opts(header: str) -> dict:
    out = {}
    for tok in header.split(",")[1:]:
        tok = tok.strip()
        if "=" in tok:
            k, v = tok.split("=", 1)
            out[k.strip().upper()] = v.strip()
        elif tok:
            out[tok.upper()] = True
    return out


def read_msh(path: str, elemopt361: int = kel361IC) -> hecmwST_local_mesh:
    """HEC-MW entire mesh (.msh) reader for the cards the hot path needs (!NODE, !ELEMENT 361 / 342, !NGROUP, !MATERIAL /
    !ITEM, !SECTION; hecmw1/src/common/hecmw_io_hec.c).  The !NODE and !ELEMENT blocks - all but a few lines of a large
    mesh - are parsed block-wise with numpy; the result is the one of the line-by-line reader (unused nodes dropped and
    nodes renumbered in ascending global id, elements sorted by id inside their type, hecmw_io_mesh.c:3039-3060)."""
    try:
        return _read_msh_bulk(path, elemopt361)
    except _Irregular:
        return _read_msh_line_by_line(path, elemopt361)


class _Irregular(Exception):
    pass


def _numbers(lines, dtype):
    with warnings.catch_warnings():
        warnings.simplefilter("error")               # a token numpy cannot read ends the parse with a warning
        try:
            return np.fromstring(" ".join(lines).replace(",", " "), dtype=dtype, sep=" ")
        except (ValueError, DeprecationWarning) as e:
            raise _Irregular(str(e))


def _read_msh_bulk(path: str, elemopt361: int) -> hecmwST_local_mesh:
    with open(path) as f:
        lines = f.read().split("\n")
    # cut the file into (header line, data lines) blocks
    blocks, cur = [], None
    for raw in lines:
        line = raw.strip()
        if not line or line.startswith("!!") or line.startswith("#"):
            continue
        if line.startswith("!"):
            cur = (line, [])
            blocks.append(cur)
            if line.split(",")[0].strip().upper() == "!END":
                break
        elif cur is not None:
            cur[1].append(line)
    node_id, node_xyz = [], []
    elems = {}                      # etype -> [ids], [connectivity rows]
    ngroups, materials, sections = {}, {}, []
    mat_name = None
    for head, data in blocks:
        key = head.split(",")[0].strip().upper()
        opts = _opts(head)
        if key == "!NODE":
            v = _numbers(data, np.float64)
            if v.size % 4:
                raise _Irregular("!NODE block")
            v = v.reshape(-1, 4)
            ids = v[:, 0].astype(np.int64)
            if not np.array_equal(ids, v[:, 0]):
                raise _Irregular("node id")
            node_id.append(ids); node_xyz.append(v[:, 1:])
        elif key == "!ELEMENT":
            et = int(opts["TYPE"])
            if et not in (361, 342):
                raise NotImplementedError("element type %d is outside the hot-path scope" % et)
            nn = 8 if et == 361 else 10
            v = _numbers(data, np.int64)
            if v.size % (nn + 1):
                raise _Irregular("!ELEMENT block")
            v = v.reshape(-1, nn + 1)
            e = elems.setdefault(et, ([], []))
            e[0].append(v[:, 0]); e[1].append(v[:, 1:])
        elif key == "!NGROUP":
            g = ngroups.setdefault(opts["NGRP"], [])
            for line in data:
                vals = [t.strip() for t in line.split(",") if t.strip() != ""]
                if opts.get("GENERATE"):
                    a_, b_, s_ = int(vals[0]), int(vals[1]), int(vals[2]) if len(vals) > 2 else 1
                    g.extend(range(a_, b_ + 1, s_))
                else:
                    g.extend(int(t) for t in vals)
        elif key == "!MATERIAL":
            mat_name = opts.get("NAME")
            materials.setdefault(mat_name, {})
        elif key == "!ITEM" or key.startswith("!ITEM="):
            if data and "E" not in materials[mat_name]:
                vals = [t.strip() for t in data[0].split(",") if t.strip() != ""]
                materials[mat_name]["E"] = float(vals[0]); materials[mat_name]["nu"] = float(vals[1])
        elif key == "!SECTION":
            sections.append(dict(opts))
    if not elems or not node_id:
        raise _Irregular("no !NODE / !ELEMENT block")
    nid = np.concatenate(node_id); xyz = np.concatenate(node_xyz)
    order = np.argsort(nid, kind="stable")
    nid_s = nid[order]
    last = np.ones(nid_s.size, bool)
    last[:-1] = nid_s[1:] != nid_s[:-1]              # a node defined twice: the later definition stands
    nid_s, xyz_s = nid_s[last], xyz[order][last]
    types = sorted(elems)
    conn = {}
    for et in types:
        ids = np.concatenate(elems[et][0]); c = np.concatenate(elems[et][1])
        if np.unique(ids).size != ids.size:
            raise _Irregular("element defined twice")
        conn[et] = c[np.argsort(ids, kind="stable")]
    used = np.unique(np.concatenate([conn[et].reshape(-1) for et in types]))
    pos = np.searchsorted(nid_s, used)
    if pos.size and (pos.max() >= nid_s.size or not np.array_equal(nid_s[pos], used)):
        raise KeyError("element node without a !NODE definition")
    m = hecmwST_local_mesh()
    m.n_node = m.nn_internal = int(used.size)
    m.node = np.ascontiguousarray(xyz_s[pos], np.float64).reshape(-1)
    m.global_node_ID = used.astype(np.int32)
    tindex, nindex, items = [0], [np.zeros(1, np.int64)], []
    for et in types:
        c = (np.searchsorted(used, conn[et]) + 1).astype(np.int32)
        if et == 342:
            c = c[:, TABLE342]                         # hecmw2fstr_mesh_conv: file order -> FrontISTR order
        items.append(c.reshape(-1))
        nindex.append(int(nindex[-1][-1]) + c.shape[1] * np.arange(1, c.shape[0] + 1, dtype=np.int64))
        tindex.append(tindex[-1] + c.shape[0])
    m.n_elem = tindex[-1]
    m.n_elem_type = len(types)
    m.elem_type_index = np.array(tindex, np.int32)
    m.elem_type_item = np.array(types, np.int32)
    m.elem_node_index = np.concatenate(nindex).astype(np.int32)
    m.elem_node_item = np.concatenate(items).astype(np.int32)
    m.section_ID = np.ones(m.n_elem, np.int32)
    sec = sections[0] if sections else {}
    mat = materials.get(sec.get("MATERIAL"), next(iter(materials.values())) if materials else {})
    m.sections = [dict(elemopt361=elemopt361, E=mat.get("E", 0.0), nu=mat.get("nu", 0.0))]
    m.node_group = {}
    for k, v in ngroups.items():                     # nodes removed as unused also leave their groups (hecmw_io_mesh.c:3039-3060)
        v = np.asarray(v, np.int64)
        p = np.searchsorted(used, v)
        ok = (p < used.size) & (used[np.minimum(p, max(used.size - 1, 0))] == v) if v.size else np.zeros(0, bool)
        m.node_group[k] = (p[ok] + 1).astype(np.int32)
    return m


def _read_msh_line_by_line(path: str, elemopt361: int = kel361IC) -> hecmwST_local_mesh:
    """The plain reader: one Python pass over the lines.  read_msh falls back to it for anything its bulk parse does not
    recognise as regular."""
    nodes = {}
    elems = {}       # etype -> list[(eid, [nodes])]
    ngroups = {}
    materials = {}
    sections = []
    cur = None
    opts = {}
    mat_name = None
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("!!") or line.startswith("#"):
                continue
            if line.startswith("!"):
                key = line.split(",")[0].strip().upper()
                opts = _opts(line)
                cur = key
                if key == "!MATERIAL":
                    mat_name = opts.get("NAME")
                    materials.setdefault(mat_name, {})
                elif key == "!ITEM":
                    # '!ITEM=1, SUBITEM=2'
                    cur = "!ITEMDATA"
                elif key.startswith("!ITEM="):
                    cur = "!ITEMDATA"
                elif key == "!SECTION":
                    sections.append(dict(opts))
                elif key == "!END":
                    break
                continue
            vals = [v.strip() for v in line.split(",") if v.strip() != ""]
            if cur == "!NODE":
                nodes[int(vals[0])] = (float(vals[1]), float(vals[2]), float(vals[3]))
            elif cur == "!ELEMENT":
                et = int(opts["TYPE"])
                elems.setdefault(et, []).append((int(vals[0]), [int(v) for v in vals[1:]]))
            elif cur == "!NGROUP":
                g = ngroups.setdefault(opts["NGRP"], [])
                if opts.get("GENERATE"):
                    a, b, s = int(vals[0]), int(vals[1]), int(vals[2]) if len(vals) > 2 else 1
                    g.extend(range(a, b + 1, s))
                else:
                    g.extend(int(v) for v in vals)
            elif cur == "!ITEMDATA":
                if "E" not in materials[mat_name]:
                    materials[mat_name]["E"] = float(vals[0])
                    materials[mat_name]["nu"] = float(vals[1])
            # !SECTION data line (thickness) and !HEADER text are ignored
    for et in elems:
        if et not in (361, 342):
            raise NotImplementedError("element type %d is outside the hot-path scope" % et)
    used = sorted({n for et in elems for _, c in elems[et] for n in c})
    gid = np.array(used, np.int64)
    loc = {g: i + 1 for i, g in enumerate(used)}
    m = hecmwST_local_mesh()
    m.n_node = m.nn_internal = len(used)
    m.node = np.array([nodes[g] for g in used], np.float64).reshape(-1)
    m.global_node_ID = gid.astype(np.int32)
    types = sorted(elems)
    tindex = [0]
    nindex = [0]
    items = []
    for et in types:
        lst = sorted(elems[et], key=lambda t: t[0])
        for _, c in lst:
            c = np.array([loc[v] for v in c], np.int32)
            if et == 342:
                c = c[TABLE342]
            items.append(c)
            nindex.append(nindex[-1] + len(c))
        tindex.append(tindex[-1] + len(lst))
    m.n_elem = tindex[-1]
    m.n_elem_type = len(types)
    m.elem_type_index = np.array(tindex, np.int32)
    m.elem_type_item = np.array(types, np.int32)
    m.elem_node_index = np.array(nindex, np.int32)
    m.elem_node_item = np.concatenate(items).astype(np.int32)
    m.section_ID = np.ones(m.n_elem, np.int32)
    sec = sections[0] if sections else {}
    mat = materials.get(sec.get("MATERIAL"), next(iter(materials.values())) if materials else {})
    m.sections = [dict(elemopt361=elemopt361, E=mat.get("E", 0.0), nu=mat.get("nu", 0.0))]
    # nodes removed as unused also leave their groups (hecmw_io_mesh.c:3039-3060)
    m.node_group = {k: np.array([loc[g] for g in v if g in loc], np.int32) for k, v in ngroups.items()}
    return m


def read_cnt(path: str) -> dict:
    out = dict(boundary=[], cload=[], solver=None)
    cur = None
    sol_lines = []
    sol_opts = {}
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("!!") or line.startswith("#"):
                continue
            if line.startswith("!"):
                key = line.split(",")[0].strip().upper()
                cur = key
                if key == "!SOLVER":
                    sol_opts = _opts(line)
                if key == "!END":
                    break
                continue
            vals = [v.strip() for v in line.split(",") if v.strip() != ""]
            if cur == "!BOUNDARY":
                out["boundary"].append((vals[0], int(vals[1]), int(vals[2]), float(vals[3]) if len(vals) > 3 else 0.0))
            elif cur == "!CLOAD":
                out["cload"].append((vals[0], int(vals[1]), float(vals[2])))
            elif cur == "!SOLVER":
                sol_lines.append(vals)
    out["solver"] = (sol_opts, sol_lines)
    return out


_METHODS = {"CG": 1, "BICGSTAB": 2, "GMRES": 3, "GPBICG": 4}


def apply_solver_card(hecMAT: hecmwST_matrix, solver) -> None:
    """!SOLVER -> Iarray/Rarray, fistr1/src/common/fstr_setup.f90:1473-1509."""
    opts, lines = solver
    I, R = hecMAT.Iarray, hecMAT.Rarray
    if "METHOD" in opts:
        mth = opts["METHOD"].upper()      # name or the reference's numeric code (examples/static/exD/D300.cnt: METHOD=1)
        I[1] = int(mth) if mth.isdigit() else _METHODS[mth]
    if "PRECOND" in opts:
        I[2] = int(opts["PRECOND"])
    if "NSET" in opts:
        I[3] = int(opts["NSET"])
    yn = {"NO": 0, "YES": 1, "VERBOSE": 2}
    if "ITERLOG" in opts:
        I[20] = yn[opts["ITERLOG"].upper()]
    if "TIMELOG" in opts:
        I[21] = yn[opts["TIMELOG"].upper()]
    if "SCALING" in opts:                 # svIarray(7), fstr_ctrl_common.f90:123
        I[6] = yn[opts["SCALING"].upper()]
    if "USEJAD" in opts:                  # svIarray(33) :126 (refused by the library: it keeps its own SELL copy)
        I[32] = yn[opts["USEJAD"].upper()]
    if "DUMPTYPE" in opts:                # svIarray(31) :124: 0..3 or NONE / MM / CSR / BSR
        d = opts["DUMPTYPE"].upper()
        I[30] = int(d) if d.isdigit() else {"NONE": 0, "MM": 1, "CSR": 2, "BSR": 3}[d]
    if "DUMPEXIT" in opts:                # svIarray(32) :125
        I[31] = yn[opts["DUMPEXIT"].upper()]
    if "METHOD2" in opts:                 # svIarray(8) :129: the Krylov method to switch to after CG diverged
        m2 = opts["METHOD2"].upper()
        I[7] = int(m2) if m2.isdigit() else _METHODS[m2]
    if len(lines) > 0:
        v = lines[0]
        if len(v) > 0: I[0] = int(v[0])
        if len(v) > 1: I[4] = min(4, max(0, int(v[1])))
        if len(v) > 2: I[5] = int(v[2])
        if len(v) > 3: I[33] = int(v[3])
    if len(lines) > 1:
        v = lines[1]
        if len(v) > 0: R[0] = float(v[0])
        if len(v) > 1: R[1] = float(v[1])
        if len(v) > 2: R[2] = float(v[2])
    I[98] = 1


class _Tok:
    def __init__(self, path):
        with open(path) as f:
            self.first = f.readline()
            self.lines = f.read().split("\n")
        self.li = 0
        self.buf = []

    def line(self):
        """next raw line (for string fields)"""
        self.buf = []
        s = self.lines[self.li]
        self.li += 1
        return s

    def _fill(self):
        while not self.buf:
            self.buf = self.lines[self.li].split()
            self.li += 1

    def i(self):
        self._fill()
        return int(self.buf.pop(0))

    def f(self):
        self._fill()
        return float(self.buf.pop(0))

    def _bulk(self, n, dtype):
        """Arrays are written a fixed number of values per line and start on a fresh line (print_int_ary /
        print_double_ary, hecmw_io_dist.c:1819-1886): take the lines the array must occupy and let numpy parse them in
        one go.  Anything that does not add up (ragged lines, a pending token buffer) goes through the token loop."""
        if self.buf or n < 64 or self.li >= len(self.lines):
            return None
        per = len(self.lines[self.li].split())
        if per == 0:
            return None
        nl = -(-n // per)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")            # a malformed token ends the parse early: caught by the size check
            try:
                arr = np.fromstring(" ".join(self.lines[self.li:self.li + nl]), dtype=dtype, sep=" ")
            except ValueError:
                return None
        if arr.size != n:
            return None
        self.li += nl
        return arr

    def ia(self, n):
        fast = self._bulk(n, np.int64)
        if fast is not None:
            return fast
        out = np.empty(n, np.int64)
        k = 0
        while k < n:
            self._fill()
            t = min(n - k, len(self.buf))
            out[k:k + t] = [int(v) for v in self.buf[:t]]
            del self.buf[:t]
            k += t
        return out

    def fa(self, n):
        fast = self._bulk(n, np.float64)
        if fast is not None:
            return fast
        out = np.empty(n, np.float64)
        k = 0
        while k < n:
            self._fill()
            t = min(n - k, len(self.buf))
            out[k:k + t] = [float(v) for v in self.buf[:t]]
            del self.buf[:t]
            k += t
        return out


def read_dist_mesh(path: str, elemopt361: int = kel361IC, E=None, nu=None) -> hecmwST_local_mesh:
    """HECMW-DMD-ASCII reader following the field order of hecmw1/src/common/hecmw_io_dist.c
    (get_global_info :245-, get_node_info, get_elem_info, get_comm_info; writer :1888-2284).
    Only the leading blocks needed by the hot path are parsed (global, node, elem, comm);
    material E/nu are supplied by the caller or parsed from the material block when simple."""
    t = _Tok(path)
    mver = re.search(r"version=(\d+)", t.first)
    version = int(mver.group(1)) if mver else 0
    # ---- global info
    t.i()  # hecmw_flag_adapt
    t.i()  # hecmw_flag_initcon
    parttype = t.i()  # hecmw_flag_parttype (1 node-based, 2 element-based)
    t.i()  # hecmw_flag_partdepth
    t.i()  # hecmw_flag_version
    if version >= 4:
        t.i()  # hecmw_flag_partcontact
    # gridfile (one line, may be empty), n_file, files, header flag (+ header line), zero_temp
    t.buf = []
    t.line()               # gridfile
    nfile = t.i()
    t.buf = []
    for _ in range(nfile):
        t.line()
    flag_header = t.i()
    t.buf = []
    if flag_header == 1:
        t.line()           # header
    zero_temp = t.f()
    # ---- node info
    n_node = t.i()
    n_node_gross = t.i() if version >= 2 else n_node
    nn_middle = t.i() if version >= 4 else n_node
    nn_internal = t.i()
    if parttype != 1 and nn_internal > 0:     # element-based / unknown: node_internal_list
        t.ia(nn_internal)
    node_ID = t.ia(2 * n_node).reshape(-1, 2)
    global_node_ID = t.ia(n_node)
    node = t.fa(3 * n_node)
    n_dof = t.i()
    n_dof_grp = t.i()
    t.ia(n_dof_grp + 1)    # node_dof_index
    t.ia(n_dof_grp)        # node_dof_item
    # initial conditions absent when flag_initcon == 0
    # ---- elem info
    n_elem = t.i()
    n_elem_gross = t.i() if version >= 2 else n_elem
    ne_internal = t.i()
    if parttype != 2 and ne_internal > 0:     # node-based / unknown: elem_internal_list
        t.ia(ne_internal)
    t.ia(2 * n_elem)       # elem_ID
    t.ia(n_elem)           # global_elem_ID
    elem_type = t.ia(n_elem)
    n_elem_type = t.i()
    elem_type_index = t.ia(n_elem_type + 1)
    elem_type_item = t.ia(n_elem_type)
    elem_node_index = t.ia(n_elem + 1)
    elem_node_item = t.ia(int(elem_node_index[-1]))
    section_ID = t.ia(n_elem)
    emi = t.ia(n_elem + 1)     # elem_mat_ID_index
    t.ia(int(emi[-1]))         # elem_mat_ID_item
    t.i()                      # n_elem_mat_ID
    # ---- comm info
    zero = t.i()
    t.i()                  # HECMW_COMM
    PETOT = t.i()
    t.i()                  # PEsmpTOT
    my_rank = t.i()
    t.i()                  # errnof
    t.i()                  # n_subdomain
    n_neighbor_pe = t.i()
    neighbor_pe = t.ia(n_neighbor_pe)
    import_index = t.ia(n_neighbor_pe + 1)
    import_item = t.ia(int(import_index[-1]))
    export_index = t.ia(n_neighbor_pe + 1)
    export_item = t.ia(int(export_index[-1]))
    m = hecmwST_local_mesh()
    m.n_node, m.nn_internal, m.n_elem, m.n_dof = n_node, nn_internal, n_elem, n_dof
    m.node = node
    m.node_ID = node_ID.astype(np.int32)
    m.global_node_ID = global_node_ID.astype(np.int32)
    m.n_elem_type = n_elem_type
    m.elem_type_index = elem_type_index.astype(np.int32)
    m.elem_type_item = elem_type_item.astype(np.int32)
    m.elem_node_index = elem_node_index.astype(np.int32)
    eni = elem_node_item.astype(np.int32)
    # hecmw2fstr_mesh_conv: 342 file order -> FrontISTR order
    s342 = elem_node_index[:-1][elem_type == 342]
    if s342.size:
        eni[s342[:, None] + np.arange(10)[None, :]] = eni[s342[:, None] + np.asarray(TABLE342)[None, :]]
    m.elem_node_item = eni
    m.section_ID = section_ID.astype(np.int32)
    m.sections = [dict(elemopt361=elemopt361, E=E, nu=nu)]
    m.n_neighbor_pe = n_neighbor_pe
    m.neighbor_pe = neighbor_pe.astype(np.int32)
    m.import_index = import_index.astype(np.int32)
    m.import_item = import_item.astype(np.int32)
    m.export_index = export_index.astype(np.int32)
    m.export_item = export_item.astype(np.int32)
    m.my_rank, m.PETOT, m.zero = my_rank, PETOT, zero
    return m


# ---------------------------------------------------------------------------------------------------
# HECMW-DMD-ASCII writer: what hecmw_part1 writes per rank (hecmw1/src/common/hecmw_io_dist.c:1888-2839,
# print_header .. print_contact_info), so that a mesh partitioned in-process (frontistr_b200/partition.py) can
# be handed to an unmodified fistr1.  Byte-identical to the real tool for the meshes meshgen.write_msh emits
# (tests/test_partition.py::test_dist_writer_equals_hecmw_part1).
# ---------------------------------------------------------------------------------------------------
def _fmt_rows(tok, cols):
    """`cols` tokens per line, a blank between tokens; a ragged last line ends with a blank before its newline - what the
    '%d' / '%.16E' + (' ' | newline every `cols`) loops of print_int_ary / print_double_ary produce."""
    n = len(tok)
    if not n:
        return ""
    full = n - n % cols
    out = [" ".join(tok[i:i + cols]) + "\n" for i in range(0, full, cols)]
    if n % cols:
        out.append(" ".join(tok[full:]) + " \n")
    return "".join(out)


def _fmt_int_ary(vals, cols=10):
    """print_int_ary (hecmw_io_dist.c:1819-1843): '%d' + (' ' | newline every `cols`), trailing newline if ragged."""
    return _fmt_rows(list(map(str, np.asarray(vals, dtype=np.int64).tolist())), cols)


def _fmt_dbl_ary(vals, cols):
    return _fmt_rows(["%.16E" % v for v in np.asarray(vals, dtype=np.float64).tolist()], cols)


def write_dist_mesh(path: str, local: hecmwST_local_mesh, global_mesh: hecmwST_local_mesh, domain: np.ndarray,
                    gridfile: str = "cube.msh", header: str = "synthetic cube", material_name: str = "M1") -> None:
    """`local` = partition.local_mesh(global_mesh, domain, rank, nparts).  Element connectivity of type 342 is
    converted back to the HEC-MW file order (fistr1/src/common/hecmw2fstr_connect_conv.c:19)."""
    from .partition import _NN
    rank, nparts = local.my_rank, local.PETOT
    owner = np.asarray(domain)
    # ---- element ownership (set_elem_belong_domain_nb, hecmw_partition.c:3163-3181): lowest node domain;
    # elem_ID = (running number among the owner's own elements, owner); shared elements per neighbour
    n_elem_g = global_mesh.n_elem
    eown = np.empty(n_elem_g, np.int64)
    touch = np.zeros((n_elem_g, nparts), bool)
    for it in range(global_mesh.n_elem_type):
        nn = _NN[int(global_mesh.elem_type_item[it])]
        e0, e1 = int(global_mesh.elem_type_index[it]), int(global_mesh.elem_type_index[it + 1])
        s0 = int(global_mesh.elem_node_index[e0])
        conn = global_mesh.elem_node_item[s0: s0 + (e1 - e0) * nn].reshape(-1, nn)
        ow = owner[conn - 1]
        eown[e0:e1] = ow.min(axis=1)
        for q in range(nparts):
            touch[e0:e1, q] = (ow == q).any(axis=1)
    own_id = np.zeros(n_elem_g, np.int64)                 # running number of every element among its owner's own elements
    for q in range(nparts):
        mine = np.nonzero(eown == q)[0]
        own_id[mine] = np.arange(1, len(mine) + 1)
    ge = local.global_elem_ID.astype(np.int64) - 1       # 0-based global element index of my elements
    etype_of = np.empty(n_elem_g, np.int64)
    for it in range(global_mesh.n_elem_type):
        etype_of[global_mesh.elem_type_index[it]: global_mesh.elem_type_index[it + 1]] = global_mesh.elem_type_item[it]
    internal = np.nonzero(eown[ge] == rank)[0] + 1
    out = []
    w = out.append
    w("!HECMW-DMD-ASCII version=4\n")
    # ---- global info (print_global_info :1909-1990)
    for v in (0, 0, 1, 1, 4, 0):                          # adapt, initcon, parttype NODE-BASED, depth 1, version 4, contact
        w("%d\n" % v)
    w("%s\n" % gridfile)
    w("0\n")                                             # n_file
    w("1\n%s\n" % header)                               # header flag + header
    w("%.16E\n" % 0.0)                                   # zero_temp
    # ---- node info (:1991-2083)
    w("%d\n%d\n%d\n%d\n" % (local.n_node, local.n_node, 0, local.nn_internal))
    w(_fmt_int_ary(local.node_ID.reshape(-1), 2))
    w(_fmt_int_ary(local.global_node_ID))
    w(_fmt_dbl_ary(local.node, 3))
    w("3\n1\n")
    w(_fmt_int_ary([0, local.nn_internal]))
    w(_fmt_int_ary([3]))
    # ---- elem info (:2084-2191)
    w("%d\n%d\n%d\n" % (local.n_elem, local.n_elem, len(internal)))
    w(_fmt_int_ary(internal))
    eid = np.stack([own_id[ge], eown[ge]], axis=1).reshape(-1)
    w(_fmt_int_ary(eid, 2))
    w(_fmt_int_ary(ge + 1))
    w(_fmt_int_ary(etype_of[ge]))
    w("%d\n" % local.n_elem_type)
    w(_fmt_int_ary(local.elem_type_index))
    w(_fmt_int_ary(local.elem_type_item))
    w(_fmt_int_ary(local.elem_node_index))
    eni = np.array(local.elem_node_item, np.int64)
    inv342 = np.argsort(TABLE342)                          # FrontISTR order -> file order
    for e in np.nonzero(etype_of[ge] == 342)[0]:
        s0 = int(local.elem_node_index[e])
        tmp = np.empty(10, np.int64); tmp[TABLE342] = eni[s0:s0 + 10]
        eni[s0:s0 + 10] = tmp
    w(_fmt_int_ary(eni))
    w(_fmt_int_ary(local.section_ID))
    w(_fmt_int_ary(np.arange(local.n_elem + 1)))          # elem_mat_ID_index
    w(_fmt_int_ary(np.ones(local.n_elem, np.int64)))      # elem_mat_ID_item
    w("%d\n" % local.n_elem)                              # n_elem_mat_ID
    # ---- comm info (:2192-2283)
    for v in (1 if rank == 0 else 0, 0, nparts, 1, rank, 0, nparts, local.n_neighbor_pe):
        w("%d\n" % v)
    w(_fmt_int_ary(local.neighbor_pe))
    w(_fmt_int_ary(local.import_index)); w(_fmt_int_ary(local.import_item))
    w(_fmt_int_ary(local.export_index)); w(_fmt_int_ary(local.export_item))
    sh_index, sh_item = [0], []
    for q in local.neighbor_pe:
        sh = np.nonzero(touch[ge, int(q)])[0] + 1
        sh_item += sh.tolist(); sh_index.append(len(sh_item))
    w(_fmt_int_ary(sh_index)); w(_fmt_int_ary(sh_item))
    # ---- section info (:2351-2410): one SOLID section, thickness 1.0, material 1
    w("1\n"); w(_fmt_int_ary([1])); w(_fmt_int_ary([0])); w(_fmt_int_ary([0, 1])); w(_fmt_int_ary([1]))
    w(_fmt_int_ary([0, 0])); w(_fmt_int_ary([0, 1])); w(_fmt_dbl_ary([1.0], 5))
    # ---- material info (:2411-2473): ITEM=1, SUBITEM=2 (E, nu)
    sct = global_mesh.sections[0]
    w("1\n1\n2\n2\n%s\n" % material_name)
    w(_fmt_int_ary([0, 1])); w(_fmt_int_ary([0, 2])); w(_fmt_int_ary([0, 1, 2]))
    w(_fmt_dbl_ary([sct["E"], sct["nu"]], 5)); w(_fmt_dbl_ary([0.0, 0.0], 5))
    # ---- mpc, amplitude (:2474-2570)
    w("0\n0\n")
    # ---- node groups (:2571-2602): ALL + the mesh's groups, members in ascending global id mapped to local ids
    g2l = np.zeros(global_mesh.n_node + 1, np.int64)
    g2l[local.global_node_ID] = np.arange(1, local.n_node + 1)
    names = ["ALL"] + list(global_mesh.node_group.keys())
    members = [np.arange(1, global_mesh.n_node + 1)] + [np.asarray(global_mesh.global_node_ID)[np.asarray(v) - 1]
                                                        for v in global_mesh.node_group.values()]
    w("%d\n" % len(names))
    for nm in names:
        w("%s\n" % nm)
    idx, items = [0], []
    for mem in members:
        l = g2l[np.asarray(mem, np.int64)]
        items += l[l > 0].tolist(); idx.append(len(items))
    w(_fmt_int_ary(idx)); w(_fmt_int_ary(items))
    # ---- element groups (:2603-2634): ALL
    w("1\nALL\n"); w(_fmt_int_ary([0, local.n_elem])); w(_fmt_int_ary(np.arange(1, local.n_elem + 1)))
    # ---- surface groups, refine info, contact pairs
    w("0\n0\n0\n")
    with open(path, "w") as f:
        f.write("".join(out))



def read_mat_dump(path: str):
    """Reads a file written by hecmw_mat_dump / _rhs / _solution (hecmw1/src/solver/matrix/hecmw_matrix_dump.f90; here
    csrc/matdump.hpp) - also one written by a genuine fistr1 run with `!SOLVER, DUMPTYPE=...`.
    .rhs / .sol -> 1-D array.  .bsr -> dict(nrow, ncol, ndof, index, item (1-based block columns), value [nnz, ndof, ndof]).
    .csr -> dict(nrow, ncol, index, item (1-based), value).  .mm -> dict(nrow, ncol, row, col (1-based), value)."""
    ext = os.path.splitext(path)[1].lower()
    with open(path) as f:
        lines = f.read().split("\n")
    if ext in (".rhs", ".sol"):
        return np.array([float(x) for x in lines if x.strip()])
    body = [ln for ln in lines if ln.strip() and not ln.startswith("%")]
    head = [int(x) for x in body[0].split()]
    if ext == ".mm":
        nrow, ncol, nnz = head
        t = np.array([ln.split() for ln in body[1: 1 + nnz]])
        return dict(nrow=nrow, ncol=ncol, row=t[:, 0].astype(np.int64), col=t[:, 1].astype(np.int64), value=t[:, 2].astype(float))
    if ext == ".csr":
        nrow, ncol, nnz = head
        index = np.array(body[1: 2 + nrow], np.int64)
        item = np.array(body[2 + nrow: 2 + nrow + nnz], np.int64)
        value = np.array(body[2 + nrow + nnz: 2 + nrow + 2 * nnz], float)
        return dict(nrow=nrow, ncol=ncol, index=index, item=item, value=value)
    if ext == ".bsr":
        nrow, ncol, nnz, ndof = head
        index = np.array(body[1: 2 + nrow], np.int64)
        item = np.array(body[2 + nrow: 2 + nrow + nnz], np.int64)
        value = np.array(body[2 + nrow + nnz: 2 + nrow + nnz + nnz * ndof * ndof], float).reshape(nnz, ndof, ndof)
        return dict(nrow=nrow, ncol=ncol, ndof=ndof, index=index, item=item, value=value)
    raise ValueError("unknown dump file type: " + path)
