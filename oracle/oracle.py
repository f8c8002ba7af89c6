"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/liboracle_hecmw.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity pinning status: see oracle/hecmw_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


class orc_mesh(C.Structure):
    _fields_ = [("n_node", C.c_int), ("nn_internal", C.c_int), ("n_elem", C.c_int), ("n_elem_type", C.c_int),
                ("elem_type_index", c_int_p), ("elem_type_item", c_int_p),
                ("elem_node_index", c_int_p), ("elem_node_item", c_int_p),
                ("node", c_dbl_p), ("section_ID", c_int_p),
                ("n_neighbor_pe", C.c_int), ("neighbor_pe", c_int_p),
                ("import_index", c_int_p), ("import_item", c_int_p),
                ("export_index", c_int_p), ("export_item", c_int_p),
                ("my_rank", C.c_int), ("PETOT", C.c_int)]


class orc_section(C.Structure):
    _fields_ = [("elemopt361", C.c_int), ("E", C.c_double), ("nu", C.c_double),
                ("mtype", C.c_int), ("hyper", C.c_double * 3)]


class orc_matrix(C.Structure):
    _fields_ = [("N", C.c_int), ("NP", C.c_int), ("NPL", C.c_int), ("NPU", C.c_int), ("NDOF", C.c_int),
                ("indexL", c_int_p), ("itemL", c_int_p), ("indexU", c_int_p), ("itemU", c_int_p),
                ("D", c_dbl_p), ("AL", c_dbl_p), ("AU", c_dbl_p), ("B", c_dbl_p), ("X", c_dbl_p),
                ("Iarray", c_int_p), ("Rarray", c_dbl_p)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_hecmw.so")
    src = os.path.join(_HERE, "hecmw_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-B", "-C", _HERE, "liboracle_hecmw.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_hecmw.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_precond_setup.restype = C.c_void_p
        _LIB.orc_precond_setup.argtypes = [C.POINTER(orc_matrix), C.c_int, C.c_double, C.c_int, C.c_int]
        _LIB.orc_precond_apply.argtypes = [C.c_void_p, c_dbl_p]
        _LIB.orc_precond_free.argtypes = [C.c_void_p]
        _LIB.orc_precond_ncolor.argtypes = [C.c_void_p]
        _LIB.orc_precond_perm.argtypes = [C.c_void_p, c_int_p]
        _LIB.orc_mat_ass_bc.argtypes = [C.POINTER(orc_matrix), C.c_int, C.c_int, C.c_double]
    return _LIB


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None and a.size else C.cast(None, c_int_p)


def _dp(a):
    return a.ctypes.data_as(c_dbl_p) if a is not None and a.size else C.cast(None, c_dbl_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _Keep:
    """ctypes struct + the numpy arrays it points into"""

    def __init__(self, struct, refs):
        self.s = struct
        self.refs = refs


def c_mesh(m) -> _Keep:
    arrs = dict(eti=_i32(m.elem_type_index), ett=_i32(m.elem_type_item), eni=_i32(m.elem_node_index),
                enn=_i32(m.elem_node_item), node=np.ascontiguousarray(m.node, np.float64),
                sec=_i32(m.section_ID), nb=_i32(m.neighbor_pe), ii=_i32(m.import_index),
                it=_i32(m.import_item), ei=_i32(m.export_index), et=_i32(m.export_item))
    s = orc_mesh(m.n_node, m.nn_internal, m.n_elem, m.n_elem_type, _ip(arrs["eti"]), _ip(arrs["ett"]),
                 _ip(arrs["eni"]), _ip(arrs["enn"]), _dp(arrs["node"]), _ip(arrs["sec"]),
                 m.n_neighbor_pe, _ip(arrs["nb"]), _ip(arrs["ii"]), _ip(arrs["it"]), _ip(arrs["ei"]),
                 _ip(arrs["et"]), m.my_rank, m.PETOT)
    return _Keep(s, arrs)


def c_matrix(A) -> _Keep:
    s = orc_matrix(A.N, A.NP, A.NPL, A.NPU, 3, _ip(A.indexL), _ip(A.itemL), _ip(A.indexU), _ip(A.itemU),
                   _dp(A.D), _dp(A.AL), _dp(A.AU), _dp(A.B), _dp(A.X), _ip(A.Iarray), _dp(A.Rarray))
    return _Keep(s, A)


def c_sections(m):
    arr = (orc_section * len(m.sections))()
    for i, s in enumerate(m.sections):
        arr[i].elemopt361 = int(s["elemopt361"]); arr[i].E = float(s["E"]); arr[i].nu = float(s["nu"])
        arr[i].mtype = int(s.get("mtype", 0))
        for k, v in enumerate(s.get("hyper", (0.0, 0.0, 0.0))):
            arr[i].hyper[k] = float(v)
    return arr


def mat_con(mesh):
    """hecmw_mat_con -> new hecmwST_matrix with index/item arrays."""
    from frontistr_b200.hecmw_types import hecmwST_matrix
    L = lib()
    cm = c_mesh(mesh)
    N, NP, NPL, NPU = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = L.orc_mat_con_build(C.byref(cm.s), C.byref(N), C.byref(NP), C.byref(NPL), C.byref(NPU))
    assert rc == 0
    A = hecmwST_matrix(N.value, NP.value, NPL.value, NPU.value)
    A.indexL = np.zeros(NP.value + 1, np.int32); A.indexU = np.zeros(NP.value + 1, np.int32)
    A.itemL = np.zeros(max(NPL.value, 1), np.int32)[:NPL.value]
    A.itemU = np.zeros(max(NPU.value, 1), np.int32)[:NPU.value]
    L.orc_mat_con_fetch(_ip(A.indexL), _ip(A.itemL), _ip(A.indexU), _ip(A.itemU))
    A.fstr_mat_init().hecMAT_init()
    return A


def stiffmatrix(mesh, A):
    cm = c_mesh(mesh); ca = c_matrix(A); sec = c_sections(mesh)
    rc = lib().orc_stiffmatrix(C.byref(cm.s), sec, C.byref(ca.s))
    assert rc == 0, rc


def element_stiffness(kind: str, ecoord: np.ndarray, E: float, nu: float, etype: int = 361) -> np.ndarray:
    """ecoord: (nn,3).  Returns (3nn,3nn) with K[i,j] (row-major view of the symmetric result)."""
    L = lib()
    nn = ecoord.shape[0]
    ec = np.ascontiguousarray(ecoord, np.float64)  # (nn,3) C-order == (3,nn) column-major
    K = np.zeros((3 * nn, 3 * nn), np.float64)
    if kind == "IC":
        L.orc_STF_C3D8IC(C.c_int(nn), _dp(ec), C.c_double(E), C.c_double(nu), _dp(K))
    elif kind == "BBAR":
        L.orc_STF_C3D8Bbar(C.c_int(nn), _dp(ec), C.c_double(E), C.c_double(nu), _dp(K))
    else:
        L.orc_STF_C3(C.c_int(etype), C.c_int(nn), _dp(ec), C.c_double(E), C.c_double(nu), _dp(K))
    return K.T.copy()   # column-major (i,j) -> K[i,j]


def ass_dload_volume(mesh, A, ltype: int, params, rho: float):
    """!DLOAD on EGRP=ALL, volume types (fstr_ass_load.f90:131-260 -> DL_C3): adds the element load vectors into A.B."""
    L = lib()
    nodes = mesh.node.reshape(-1, 3)
    pr = np.zeros(7)
    pr[: len(params)] = params
    for it in range(mesh.n_elem_type):
        et = int(mesh.elem_type_item[it])
        for e in range(int(mesh.elem_type_index[it]), int(mesh.elem_type_index[it + 1])):
            nod = mesh.elem_node_item[mesh.elem_node_index[e]: mesh.elem_node_index[e + 1]]
            ec = np.ascontiguousarray(nodes[nod - 1])
            v = np.zeros(3 * len(nod))
            rc = L.orc_DL_C3_volume(et, len(nod), _dp(ec), C.c_double(rho), int(ltype), _dp(pr), _dp(v))
            assert rc == 0
            np.add.at(A.B.reshape(-1, 3), nod - 1, v.reshape(-1, 3))


def ass_thermal_load(mesh, A, temp, temp0, alp: float, ref_temp: float):
    """!TEMPERATURE load (fstr_ass_load.f90:277-430): TLOAD_C3D8IC for the 361 IC element, TLOAD_C3 otherwise."""
    L = lib()
    nodes = mesh.node.reshape(-1, 3)
    for it in range(mesh.n_elem_type):
        et = int(mesh.elem_type_item[it])
        for e in range(int(mesh.elem_type_index[it]), int(mesh.elem_type_index[it + 1])):
            nod = mesh.elem_node_item[mesh.elem_node_index[e]: mesh.elem_node_index[e + 1]]
            sec = mesh.sections[int(mesh.section_ID[e]) - 1]
            ec = np.ascontiguousarray(nodes[nod - 1])
            tt = np.ascontiguousarray(np.asarray(temp, float)[nod - 1]); t0 = np.ascontiguousarray(np.asarray(temp0, float)[nod - 1])
            v = np.zeros(3 * len(nod))
            if et == 361 and sec["elemopt361"] == 3:
                L.orc_TLOAD_C3D8IC(len(nod), _dp(ec), _dp(tt), _dp(t0), C.c_double(sec["E"]), C.c_double(sec["nu"]),
                                   C.c_double(alp), C.c_double(ref_temp), _dp(v))
            else:
                L.orc_TLOAD_C3(et, len(nod), _dp(ec), _dp(tt), _dp(t0), C.c_double(sec["E"]), C.c_double(sec["nu"]),
                               C.c_double(alp), C.c_double(ref_temp), _dp(v))
            np.add.at(A.B.reshape(-1, 3), nod - 1, v.reshape(-1, 3))


def ass_dload_pressure(mesh, A, elems, face: int, val: float):
    """!DLOAD P<face> on an element group (0-based element indices): adds the face pressure load vectors into A.B."""
    L = lib()
    nodes = mesh.node.reshape(-1, 3)
    et_of = np.empty(mesh.n_elem, np.int32)
    for it in range(mesh.n_elem_type):
        et_of[mesh.elem_type_index[it]: mesh.elem_type_index[it + 1]] = mesh.elem_type_item[it]
    for e in elems:
        nod = mesh.elem_node_item[mesh.elem_node_index[e]: mesh.elem_node_index[e + 1]]
        ec = np.ascontiguousarray(nodes[nod - 1])
        v = np.zeros(3 * len(nod))
        assert L.orc_DL_C3_surface(int(et_of[e]), len(nod), _dp(ec), int(face), C.c_double(val), _dp(v)) == 0
        np.add.at(A.B.reshape(-1, 3), nod - 1, v.reshape(-1, 3))


def mat_ass_bc(A, inode: int, idof: int, rhs: float):
    ca = c_matrix(A)
    lib().orc_mat_ass_bc(C.byref(ca.s), int(inode), int(idof), float(rhs))


def matvec(A, x: np.ndarray) -> np.ndarray:
    ca = c_matrix(A)
    y = np.zeros(3 * A.NP)
    x = np.ascontiguousarray(x, np.float64)
    lib().orc_matvec_33_local(C.byref(ca.s), _dp(x), _dp(y))
    return y


def solve(meshes, mats, hist_len: int = 0, ssor_threads: int = 0):
    """hecmw_solve on ndom rank-emulated subdomains.  Returns dict(error, iter, resid, resid2, hist, times)."""
    L = lib()
    if not isinstance(meshes, (list, tuple)):
        meshes, mats = [meshes], [mats]
    nd = len(meshes)
    cms = [c_mesh(m) for m in meshes]
    cas = [c_matrix(a) for a in mats]
    marr = (orc_mesh * nd)(*[k.s for k in cms])
    aarr = (orc_matrix * nd)(*[k.s for k in cas])
    it, rs, rs2 = C.c_int(), C.c_double(), C.c_double()
    hist = np.zeros(max(hist_len, 1))
    times = np.zeros(5)
    L.orc_set_ssor_threads(int(ssor_threads))
    err = L.orc_solve(nd, marr, aarr, C.byref(it), C.byref(rs), C.byref(rs2), _dp(hist), int(hist_len), _dp(times))
    return dict(error=err, iter=it.value, resid=rs.value, resid2=rs2.value,
                hist=hist[: min(hist_len, max(it.value, 0))], times=times)


def nlgeom_static(mesh, A, boundary, cload, substeps=10, converg=1.0e-3, max_iter=50, spring=((), (), ())):
    """`!STATIC, TYPE=NLGEOM` with one `!STEP, SUBSTEPS=n` (fstr_solve_NLGEOM + fstr_Newton), total Lagrangian,
    elastic.  boundary = (nodes, dofs, values) triplets as fstr_AddBC visits them; cload, spring = (nodes, dofs, values).
    Returns dict(error, unode, unode_steps[substeps, 3*n_node], newton_iters, cg_iters)."""
    L = lib()
    cm = c_mesh(mesh); ca = c_matrix(A); sec = c_sections(mesh)
    bn, bd, bv = _i32(boundary[0]), _i32(boundary[1]), np.ascontiguousarray(boundary[2], np.float64)
    cn, cd, cv = _i32(cload[0]), _i32(cload[1]), np.ascontiguousarray(cload[2], np.float64)
    sn, sd, sv = _i32(spring[0]), _i32(spring[1]), np.ascontiguousarray(spring[2], np.float64)
    un = np.zeros(3 * mesh.n_node)
    steps = np.zeros((substeps, 3 * mesh.n_node))
    nit = np.zeros(substeps, np.int32)
    cg = C.c_int()
    precond_clear()
    rc = L.orc_nlgeom_static(C.byref(cm.s), sec, C.byref(ca.s), len(bn), _ip(bn), _ip(bd), _dp(bv), len(cn), _ip(cn),
                             _ip(cd), _dp(cv), len(sn), _ip(sn), _ip(sd), _dp(sv), C.c_int(substeps), C.c_double(converg),
                             C.c_int(max_iter), _dp(un),
                             _dp(steps), _ip(nit), C.byref(cg))
    precond_clear()
    return dict(error=rc, unode=un, unode_steps=steps, newton_iters=nit, cg_iters=cg.value)


def nlgeom_state(mesh, substep: int):
    """Gauss-point strain / stress [n_elem][8][6] after sub-step `substep` (1-based) of the last nlgeom_static run."""
    e = np.zeros((mesh.n_elem, 8, 6)); s = np.zeros((mesh.n_elem, 8, 6))
    assert lib().orc_nlgeom_fetch_state(int(substep), _dp(e), _dp(s)) == 0
    return e, s


def precond_clear():
    lib().orc_precond_clear()


class Precond:
    def __init__(self, A, precond=3, sigma_diag=1.0, ncolor_in=10, nthreads=1):
        self._ca = c_matrix(A)
        self.h = lib().orc_precond_setup(C.byref(self._ca.s), precond, float(sigma_diag), ncolor_in, nthreads)
        self.N = A.N

    def apply(self, zp: np.ndarray):
        lib().orc_precond_apply(self.h, _dp(zp))
        return zp

    @property
    def ncolor(self):
        return lib().orc_precond_ncolor(self.h)

    def perm(self):
        p = np.zeros(self.N, np.int32)
        lib().orc_precond_perm(self.h, _ip(p))
        return p

    def __del__(self):
        try:
            lib().orc_precond_free(self.h)
        except Exception:
            pass


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def num_threads() -> int:
    return lib().orc_num_threads()


# ---- strain / stress recovery of the linear static analysis (fstr_UpdateNewton INFINITE + fstr_NodalStress3D) -------
def update_linear(mesh, utot, TT=None, T0=None, alpha=0.0, reftemp=0.0):
    """Gauss-point strain / stress [n_elem][8][6] and QFORCE(3*n_node) at utot = unode + dunode."""
    L = lib()
    cm, sec = c_mesh(mesh), c_sections(mesh)
    u = np.ascontiguousarray(utot, np.float64)
    e = np.zeros((mesh.n_elem, 8, 6)); s = np.zeros((mesh.n_elem, 8, 6)); q = np.zeros(3 * mesh.n_node)
    tt = np.ascontiguousarray(TT, np.float64) if TT is not None else None
    t0 = np.ascontiguousarray(T0, np.float64) if T0 is not None else None
    L.orc_update_linear.argtypes = [C.POINTER(orc_mesh), C.POINTER(orc_section), c_dbl_p, c_dbl_p, c_dbl_p, C.c_double,
                                    C.c_double, c_dbl_p, c_dbl_p, c_dbl_p]
    rc = L.orc_update_linear(C.byref(cm.s), sec, _dp(u), _dp(tt), _dp(t0), float(alpha), float(reftemp), _dp(e), _dp(s), _dp(q))
    assert rc == 0, rc
    return e, s, q


def nodal_stress_3d(mesh, gp_strain, gp_stress):
    """fstr_NodalStress3D -> dict(STRAIN (n_node,6), STRESS, MISES, ESTRAIN (n_elem,6), ESTRESS, EMISES)."""
    L = lib()
    cm = c_mesh(mesh)
    ge = np.ascontiguousarray(gp_strain, np.float64); gs = np.ascontiguousarray(gp_stress, np.float64)
    out = dict(STRAIN=np.zeros((mesh.n_node, 6)), STRESS=np.zeros((mesh.n_node, 6)), MISES=np.zeros(mesh.n_node),
               ESTRAIN=np.zeros((mesh.n_elem, 6)), ESTRESS=np.zeros((mesh.n_elem, 6)), EMISES=np.zeros(mesh.n_elem))
    L.orc_nodal_stress_3d.argtypes = [C.POINTER(orc_mesh)] + [c_dbl_p] * 8
    rc = L.orc_nodal_stress_3d(C.byref(cm.s), _dp(ge), _dp(gs), *[_dp(out[k]) for k in
                                                                  ("STRAIN", "STRESS", "MISES", "ESTRAIN", "ESTRESS", "EMISES")])
    assert rc == 0, rc
    return out


def element_update_linear(kind, ecoord, u, E, nu, etype=361, TT=None, T0=None, alpha=0.0, reftemp=0.0):
    """One element: kind 'IC' / 'FI' / 'BBAR' / 'C3' -> (strain [ngp][6], stress [ngp][6], qf [3 nn])."""
    L = lib()
    ec = np.ascontiguousarray(ecoord, np.float64); uu = np.ascontiguousarray(u, np.float64)
    nn = ec.shape[0]
    ngp = 8 if etype == 361 else 4
    e = np.zeros((ngp, 6)); s = np.zeros((ngp, 6)); q = np.zeros(3 * nn)
    tt = np.ascontiguousarray(TT, np.float64) if TT is not None else None
    t0 = np.ascontiguousarray(T0, np.float64) if T0 is not None else None
    if kind == "IC":
        L.orc_UpdateST_C3D8IC.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_double, C.c_double, c_dbl_p, c_dbl_p, C.c_double,
                                          C.c_double, c_dbl_p, c_dbl_p, c_dbl_p]
        L.orc_UpdateST_C3D8IC(nn, _dp(ec), _dp(uu), E, nu, _dp(tt), _dp(t0), alpha, reftemp, _dp(q), _dp(e), _dp(s))
    else:
        L.orc_UPDATE_C3_lin.argtypes = [C.c_int, C.c_int, c_dbl_p, c_dbl_p, C.c_double, C.c_double, C.c_int, c_dbl_p, c_dbl_p,
                                        C.c_double, C.c_double, c_dbl_p, c_dbl_p, c_dbl_p]
        L.orc_UPDATE_C3_lin(etype, nn, _dp(ec), _dp(uu), E, nu, int(kind == "BBAR"), _dp(tt), _dp(t0), alpha, reftemp,
                            _dp(q), _dp(e), _dp(s))
    return e, s, q


def part_rcb(mesh, axes: str):
    """Node -> domain map of the reference's RCB (hecmw_partition.c:1570-1660) incl. the tie behaviour of its quick sort."""
    ax = _i32([{"x": 0, "y": 1, "z": 2}[a] for a in axes.lower().replace(",", "")])
    dom = np.zeros(mesh.n_node, np.int32)
    node = np.ascontiguousarray(mesh.node, np.float64)
    L = lib()
    L.orc_part_rcb.argtypes = [C.c_int, c_dbl_p, C.c_int, c_int_p, c_int_p]
    rc = L.orc_part_rcb(mesh.n_node, _dp(node), len(ax), _ip(ax), _ip(dom))
    assert rc == 0, rc
    return dom
