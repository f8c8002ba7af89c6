"""Host-side mirror of the reference's call sites for the hot path, over the C-ABI.

Same names and argument meaning as the Fortran interface the path is a drop-in for:

    hecmw_mat_con(hecMESH, hecMAT)            hecmw1/src/solver/matrix/hecmw_mat_con.f90:23
    fstr_StiffMatrix(hecMESH, hecMAT)         fistr1/src/analysis/static/fstr_StiffMatrix.f90:18
    hecmw_mat_ass_bc(hecMAT, inode, idof, RHS)  hecmw1/src/solver/matrix/hecmw_mat_ass.f90:292 (batched)
    hecmw_solve(hecMESH, hecMAT)              hecmw1/src/solver/hecmw_solver.f90:9
    hecmw_matvec(hecMESH, hecMAT, X, Y)       hecmw1/src/solver/las/hecmw_solver_las.f90:57
    hecmw_precond_setup/apply/clear           hecmw1/src/solver/precond/hecmw_precond.f90:28-123
    hecmw_update_3_R(hecMESH, X)              hecmw1/src/solver/communication/hecmw_comm_f.F90:669

All compute runs in frontistr_b200/libfistr_b200.so (hand-written sm_100a kernels); this module only
marshals numpy buffers.  It never imports oracle/.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dp, ip
from .hecmw_types import hecmwST_local_mesh, hecmwST_matrix

_i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)


def init(device: int = -1):
    check(_lib.lib().fistr_b200_init(int(device)))


def finalize():
    check(_lib.lib().fistr_b200_finalize())


def set_option(key: str, value: int):
    check(_lib.lib().fistr_b200_set_option(key.encode(), int(value)))


def get_option(key: str) -> int:
    return int(_lib.lib().fistr_b200_get_option(key.encode()))


def launch_count() -> int:
    return int(_lib.lib().fistr_b200_launch_count())


def device_info():
    name = C.create_string_buffer(256)
    sm = C.c_int()
    mem = C.c_int64()
    check(_lib.lib().fistr_b200_device_info(name, 256, C.byref(sm), C.byref(mem)))
    return dict(name=name.value.decode(), sm_count=sm.value, mem_bytes=mem.value)


# ---- communicator ---------------------------------------------------------------------------
def comm_init_from_torch():
    """One process per GPU: broadcast the NCCL id over torch.distributed, then hand it to the library."""
    import torch.distributed as dist
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    idbuf = (C.c_char * 128)()
    if rank == 0:
        check(_lib.lib().fistr_b200_comm_get_unique_id(idbuf))
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    check(_lib.lib().fistr_b200_comm_init(raw, rank, world))


def comm_finalize():
    check(_lib.lib().fistr_b200_comm_finalize())


# ---- hecmw_mat_con ---------------------------------------------------------------------------
def hecmw_mat_con(hecMESH: hecmwST_local_mesh, hecMAT: hecmwST_matrix | None = None) -> hecmwST_matrix:
    L = _lib.lib()
    A = hecMAT if hecMAT is not None else hecmwST_matrix()
    N, NP, NPL, NPU = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    eti, ett = _i32(hecMESH.elem_type_index), _i32(hecMESH.elem_type_item)
    eni, enn = _i32(hecMESH.elem_node_index), _i32(hecMESH.elem_node_item)
    check(L.fistr_b200_mat_con_count(hecMESH.n_node, hecMESH.nn_internal, hecMESH.n_elem_type, ip(eti), ip(ett),
                                     ip(eni), ip(enn), C.byref(N), C.byref(NP), C.byref(NPL), C.byref(NPU)))
    A.N, A.NP, A.NPL, A.NPU, A.NDOF = N.value, NP.value, NPL.value, NPU.value, 3
    A.indexL = np.zeros(A.NP + 1, np.int32)
    A.indexU = np.zeros(A.NP + 1, np.int32)
    A.itemL = np.zeros(A.NPL, np.int32)
    A.itemU = np.zeros(A.NPU, np.int32)
    check(L.fistr_b200_mat_con_fill(ip(A.indexL), ip(A.itemL), ip(A.indexU), ip(A.itemU)))
    return A


def _structure(hecMESH, hecMAT):
    m, A = hecMESH, hecMAT
    m._c = dict(nb=_i32(m.neighbor_pe), ii=_i32(m.import_index), it=_i32(m.import_item),
                ei=_i32(m.export_index), et=_i32(m.export_item))
    c = m._c
    check(_lib.lib().fistr_b200_mat_structure(A.N, A.NP, A.NPL, A.NPU, ip(A.indexL), ip(A.itemL), ip(A.indexU),
                                             ip(A.itemU), m.n_neighbor_pe, ip(c["nb"]), ip(c["ii"]), ip(c["it"]),
                                             ip(c["ei"]), ip(c["et"])))


def mat_structure(hecMESH, hecMAT):
    _structure(hecMESH, hecMAT)


def mat_upload(hecMAT, values=True, rhs=True):
    A = hecMAT
    check(_lib.lib().fistr_b200_mat_upload(dp(A.D) if values else None, dp(A.AL) if values else None,
                                          dp(A.AU) if values else None, dp(A.B) if rhs else None))


def mat_download(hecMAT, values=True, rhs=True, x=False):
    A = hecMAT
    if values and A.D is None:
        A.hecMAT_init()
    check(_lib.lib().fistr_b200_mat_download(dp(A.D) if values else None, dp(A.AL) if values else None,
                                            dp(A.AU) if values else None, dp(A.B) if rhs else None,
                                            dp(A.X) if x else None))


def host_register(*arrays):
    """Page-lock numpy arrays that will be passed repeatedly through the borrowed-pointer C-ABI."""
    for a in arrays:
        if a is not None and a.size:
            check(_lib.lib().fistr_b200_host_register(a.ctypes.data, a.nbytes))


def host_unregister(*arrays):
    for a in arrays:
        if a is not None and a.size:
            check(_lib.lib().fistr_b200_host_unregister(a.ctypes.data))


# ---- fstr_StiffMatrix ------------------------------------------------------------------------
def fstr_StiffMatrix(hecMESH, hecMAT, to_host: bool = True):
    """Element integration + assembly on the device.  With to_host the D/AL/AU arrays of hecMAT are
    filled (drop-in behaviour); otherwise the matrix stays device-resident.  Returns kernel/total ms."""
    m, A = hecMESH, hecMAT
    if A.D is None and to_host:
        A.hecMAT_init()
    eti, ett = _i32(m.elem_type_index), _i32(m.elem_type_item)
    if not hasattr(m, "_conn") or m._conn[2] is not m.elem_node_item:
        m._conn = (_i32(m.elem_node_index), _i32(m.elem_node_item), m.elem_node_item)
    eni, enn = m._conn[0], m._conn[1]
    sec = _i32(m.section_ID)
    ns = len(m.sections)
    form = np.array([s["elemopt361"] for s in m.sections], np.int32)
    E = np.array([s["E"] for s in m.sections], np.float64)
    nu = np.array([s["nu"] for s in m.sections], np.float64)
    node = np.ascontiguousarray(m.node, np.float64)
    t = np.zeros(2)
    check(_lib.lib().fistr_b200_stiffmatrix_3d(m.n_node, dp(node), m.n_elem, m.n_elem_type, ip(eti), ip(ett), ip(eni),
                                              ip(enn), ip(sec), ns, ip(form), dp(E), dp(nu),
                                              dp(A.D) if to_host else None, dp(A.AL) if to_host else None,
                                              dp(A.AU) if to_host else None, dp(t)))
    return dict(kernel_ms=float(t[0]), total_ms=float(t[1]))


# ---- strain / stress recovery of the linear static analysis ---------------------------------------------------
def fstr_UpdateNewton_linear(hecMESH, unode, dunode, temperature=None, temperature0=None, ref_temp=0.0):
    """fstr_UpdateNewton with nlgeom_flag == INFINITE (fstr_Update.f90:25-293): strain / stress of every Gauss point
    at unode + dunode (kept on the device) and QFORCE (returned).  Sections may carry "alpha" (expansion coefficient)."""
    m = hecMESH
    eti, ett = _i32(m.elem_type_index), _i32(m.elem_type_item)
    if not hasattr(m, "_conn") or m._conn[2] is not m.elem_node_item:
        m._conn = (_i32(m.elem_node_index), _i32(m.elem_node_item), m.elem_node_item)
    eni, enn = m._conn[0], m._conn[1]
    sec = _i32(m.section_ID)
    ns = len(m.sections)
    form = np.array([s["elemopt361"] for s in m.sections], np.int32)
    E = np.array([s["E"] for s in m.sections], np.float64)
    nu = np.array([s["nu"] for s in m.sections], np.float64)
    alpha = np.array([s.get("alpha", 0.0) for s in m.sections], np.float64)
    node = np.ascontiguousarray(m.node, np.float64)
    un = np.ascontiguousarray(unode, np.float64) if unode is not None else None
    dun = np.ascontiguousarray(dunode, np.float64) if dunode is not None else None
    tt = np.ascontiguousarray(temperature, np.float64) if temperature is not None else None
    t0 = np.ascontiguousarray(temperature0, np.float64) if temperature0 is not None else None
    qf = np.zeros(3 * m.n_node)
    check(_lib.lib().fistr_b200_update_3d(m.n_node, dp(node), m.n_elem, m.n_elem_type, ip(eti), ip(ett), ip(eni), ip(enn),
                                   ip(sec), ns, ip(form), dp(E), dp(nu), dp(alpha) if tt is not None else None,
                                   dp(un) if un is not None else None, dp(dun) if dun is not None else None,
                                   dp(tt) if tt is not None else None, dp(t0) if t0 is not None else None,
                                   C.c_double(ref_temp), dp(qf)))
    return qf


def fstr_NodalStress3D(hecMESH):
    """fstr_NodalStress3D (fstr_NodalStress.f90:13-272) on the Gauss-point state of the last update: dict with
    STRAIN / STRESS (n_node, 6), MISES (n_node), ESTRAIN / ESTRESS (n_elem, 6), EMISES (n_elem)."""
    m = hecMESH
    out = dict(STRAIN=np.zeros((m.n_node, 6)), STRESS=np.zeros((m.n_node, 6)), MISES=np.zeros(m.n_node),
               ESTRAIN=np.zeros((m.n_elem, 6)), ESTRESS=np.zeros((m.n_elem, 6)), EMISES=np.zeros(m.n_elem))
    check(_lib.lib().fistr_b200_nodal_stress_3d(*[dp(out[k]) for k in ("STRAIN", "STRESS", "MISES", "ESTRAIN", "ESTRESS", "EMISES")]))
    return out


# ---- nonlinear static (total Lagrangian): fstr_StiffMatrix / fstr_UpdateNewton under NLGEOM, fstr_Newton ----------
def _nl_args(m):
    eti, ett = _i32(m.elem_type_index), _i32(m.elem_type_item)
    if not hasattr(m, "_conn") or m._conn[2] is not m.elem_node_item:
        m._conn = (_i32(m.elem_node_index), _i32(m.elem_node_item), m.elem_node_item)
    eni, enn = m._conn[0], m._conn[1]
    sec = _i32(m.section_ID)
    ns = len(m.sections)
    form = np.array([s["elemopt361"] for s in m.sections], np.int32)
    E = np.array([s["E"] for s in m.sections], np.float64)
    nu = np.array([s["nu"] for s in m.sections], np.float64)
    mtype = np.array([s.get("mtype", 0) for s in m.sections], np.int32)
    hyper = np.array([list(s.get("hyper", (0.0, 0.0, 0.0))) for s in m.sections], np.float64).reshape(-1)
    node = np.ascontiguousarray(m.node, np.float64)
    keep = (eti, ett, eni, enn, sec, form, E, nu, mtype, hyper, node)
    return keep, (m.n_node, dp(node), m.n_elem, m.n_elem_type, ip(eti), ip(ett), ip(eni), ip(enn), ip(sec), ns, ip(form),
                  dp(E), dp(nu), ip(mtype), dp(hyper))


def nl_state_reset(hecMESH):
    check(_lib.lib().fistr_b200_nl_state_reset(hecMESH.n_elem))


def nl_state_download(hecMESH):
    e = np.zeros((hecMESH.n_elem, 8, 6)); s = np.zeros((hecMESH.n_elem, 8, 6))
    check(_lib.lib().fistr_b200_nl_state_download(dp(e), dp(s)))
    return e, s


def fstr_StiffMatrix_NL(hecMESH, hecMAT, u):
    """fstr_StiffMatrix with nlgeom_flag == TOTALLAG: tangent stiffness at u = unode + dunode, left on the device."""
    keep, args = _nl_args(hecMESH)
    u = np.ascontiguousarray(u, np.float64)
    t = np.zeros(2)
    check(_lib.lib().fistr_b200_nl_stiffmatrix_3d(*args, dp(u), None, None, None, dp(t)))
    return dict(kernel_ms=float(t[0]), total_ms=float(t[1]))


def fstr_UpdateNewton(hecMESH, unode, dunode):
    """fstr_UpdateNewton: strain / stress of every Gauss point (device) and QFORCE (returned)."""
    keep, args = _nl_args(hecMESH)
    un = np.ascontiguousarray(unode, np.float64); dun = np.ascontiguousarray(dunode, np.float64)
    qf = np.zeros(3 * hecMESH.n_node)
    check(_lib.lib().fistr_b200_nl_update_3d(*args, dp(un), dp(dun), dp(qf)))
    return qf


def fstr_solve_NLGEOM(hecMESH, hecMAT, boundary, cload, substeps=10, converg=1.0e-3, max_iter=50):
    """One loading step of `!STATIC, TYPE=NLGEOM` (fstr_solve_NLGEOM.f90:84-160 + fstr_Newton,
    fstr_solve_NonLinear.f90:29-167), single subdomain: tangent assembly, Dirichlet elimination, CG solve and the
    strain / stress / internal-force update all run on the device through the C-ABI; this loop only moves the
    3*n_node vectors.  boundary = (nodes, dofs, values) as fstr_AddBC visits them, cload likewise.
    Returns dict(error, unode, unode_steps, newton_iters, cg_iters)."""
    m, A = hecMESH, hecMAT
    n3 = 3 * m.n_node
    bn, bd, bv = _i32(boundary[0]), _i32(boundary[1]), np.ascontiguousarray(boundary[2], np.float64)
    cn, cd, cv = np.asarray(cload[0], np.int64), np.asarray(cload[1], np.int64), np.asarray(cload[2], np.float64)
    if A.X is None or A.B is None:
        A.B = np.zeros(n3); A.X = np.zeros(n3)
    mat_structure(m, A)
    nl_state_reset(m)
    unode = np.zeros(n3); QF = np.zeros(n3)
    steps = np.zeros((substeps, n3)); nit = np.zeros(substeps, np.int32)
    cg_total, time, dt, rc = 0, 0.0, 1.0 / substeps, 0
    bidx = 3 * (bn.astype(np.int64) - 1) + bd - 1
    for sub in range(1, substeps + 1):
        f1, f2 = min(time, 1.0), min(time + dt, 1.0)
        dunode = np.zeros(n3)
        GL = np.zeros(n3)
        if len(cn):
            np.add.at(GL, 3 * (cn - 1) + cd - 1, f2 * cv)
        B = GL - QF
        converged = False
        for it in range(1, max_iter + 1):
            fstr_StiffMatrix_NL(m, A, unode + dunode)
            A.B[:] = B
            mat_upload(A, values=False, rhs=True)
            hecmw_mat_ass_bc(bn, bd, bv * (f2 - f1) if it == 1 else np.zeros(len(bv)))
            if sub == 1 and it == 1:
                A.Iarray[97] = 1
            A.Iarray[96] = 2 if it == 1 else 1
            A.X[:] = 0.0
            r = hecmw_solve(m, A, device_resident=True)
            cg_total += r["iter"]
            if r["error"] not in (0, 2002):
                rc = r["error"]; break
            dunode += A.X
            QF = fstr_UpdateNewton(m, unode, dunode)
            B = GL - QF
            B[bidx] = 0.0
            N3 = 3 * m.nn_internal
            res = np.sqrt(B[:N3] @ B[:N3]); xnrm = np.sqrt(A.X[:N3] @ A.X[:N3]); qnrm = np.sqrt(QF[:N3] @ QF[:N3])
            if qnrm < 1.0e-8:
                qnrm = 1.0
            dunrm = xnrm if it == 1 else np.sqrt(dunode[:N3] @ dunode[:N3])
            if A.Iarray[80] == 1 and (res / qnrm < converg or xnrm / dunrm < converg):
                converged = True; break
            if it == max_iter or res / qnrm > 1.0e10:
                rc = -1; break
        if rc or not converged:
            rc = rc or -1
            break
        nit[sub - 1] = it
        unode += dunode
        steps[sub - 1] = unode
        time += dt
    return dict(error=rc, unode=unode, unode_steps=steps, newton_iters=nit, cg_iters=cg_total)


# ---- hecmw_mat_ass_bc ------------------------------------------------------------------------
def hecmw_mat_ass_bc(nodes, dofs, values):
    """Batched Dirichlet elimination on the device-resident matrix/RHS (one entry per constrained DOF,
    in the order fstr_AddBC visits them)."""
    n = _i32(nodes); d = _i32(dofs); v = np.ascontiguousarray(values, np.float64)
    check(_lib.lib().fistr_b200_mat_ass_bc(len(n), ip(n), ip(d), dp(v)))


# ---- hecmw_solve -----------------------------------------------------------------------------
def hecmw_solve(hecMESH, hecMAT, device_resident: bool = False):
    """hecmw_solve(hecMESH, hecMAT).  Result in hecMAT.X, flags in hecMAT.Iarray(81/82), as in the
    reference.  Returns dict(error, iter, resid, resid2, times, hist)."""
    m, A = hecMESH, hecMAT
    if not hasattr(m, "_c"):
        m._c = dict(nb=_i32(m.neighbor_pe), ii=_i32(m.import_index), it=_i32(m.import_item),
                    ei=_i32(m.export_index), et=_i32(m.export_item))
    c = m._c
    it, rs = C.c_int(), C.c_double()
    times = np.zeros(5)
    L = _lib.lib()
    rc = L.fistr_b200_solve_33(A.N, A.NP, A.NPL, A.NPU,
                               None if device_resident else dp(A.D), None if device_resident else dp(A.AL),
                               None if device_resident else dp(A.AU),
                               ip(A.indexL), ip(A.itemL), ip(A.indexU), ip(A.itemU),
                               None if device_resident else dp(A.B), dp(A.X), ip(A.Iarray), dp(A.Rarray),
                               m.n_neighbor_pe, ip(c["nb"]), ip(c["ii"]), ip(c["it"]), ip(c["ei"]), ip(c["et"]),
                               m.my_rank, m.PETOT, C.byref(it), C.byref(rs), dp(times))
    check(rc, allow_solver_codes=True)
    hist = np.zeros(max(min(int(A.Iarray[0]), 1 << 22), 1))      # the library keeps at most 4 Mi entries
    nh = L.fistr_b200_solve_history(dp(hist), len(hist))
    r2 = C.c_double()
    L.fistr_b200_solve_resid2(C.byref(r2))
    prof = np.zeros(4)
    L.fistr_b200_solve_profile(dp(prof))
    return dict(error=rc, iter=it.value, resid=rs.value, resid2=r2.value, times=times, hist=hist[:nh], profile=prof)


def hecmw_mat_dump(dump_type: int, what: int = 1, numcall: int = 0, directory: str = ""):
    """hecmw_mat_dump + hecmw_mat_dump_rhs (what=1) / hecmw_mat_dump_solution (what=2) of the device-resident system
    (matrix/hecmw_matrix_dump.f90); dump_type 1 MM, 2 CSR, 3 BSR."""
    check(_lib.lib().fistr_b200_mat_dump(int(dump_type), int(what), int(numcall), directory.encode()))


def hecmw_matvec(X: np.ndarray, n_internal3: int) -> np.ndarray:
    x = np.ascontiguousarray(X, np.float64)
    y = np.zeros(n_internal3)
    check(_lib.lib().fistr_b200_matvec_33(dp(x), dp(y)))
    return y


def hecmw_matresid(X: np.ndarray, B, n_internal3: int):
    """hecmw_matresid + hecmw_rel_resid_L2: (R = B - A X on the owned rows, sqrt(r.r / b.b)); B = None: the resident RHS."""
    x = np.ascontiguousarray(X, np.float64)
    b = None if B is None else np.ascontiguousarray(B, np.float64)
    r = np.zeros(n_internal3)
    rel = C.c_double()
    check(_lib.lib().fistr_b200_matresid_33(dp(x), dp(b) if b is not None else None, dp(r), C.byref(rel)))
    return r, rel.value


def matvec_bench(nrep=20, flush_l2=False):
    a, b = C.c_double(), C.c_double()
    check(_lib.lib().fistr_b200_matvec_bench(int(nrep), int(flush_l2), C.byref(a), C.byref(b)))
    return a.value, b.value


def hecmw_precond_setup(precond: int, sigma_diag: float = 1.0, ncolor_in: int = 10):
    check(_lib.lib().fistr_b200_precond_33_setup(int(precond), float(sigma_diag), int(ncolor_in)))


def hecmw_precond_apply(ZP: np.ndarray) -> np.ndarray:
    check(_lib.lib().fistr_b200_precond_33_apply(dp(ZP)))
    return ZP


def hecmw_precond_clear():
    check(_lib.lib().fistr_b200_precond_33_clear())


def precond_info(N: int):
    nc = C.c_int()
    perm = np.zeros(max(N, 1), np.int32)
    check(_lib.lib().fistr_b200_precond_33_info(C.byref(nc), ip(perm)))
    return nc.value, perm[:N]


def hecmw_update_3_R(X: np.ndarray) -> np.ndarray:
    check(_lib.lib().fistr_b200_update_3_R(dp(X)))
    return X


def fp64_peak() -> float:
    v = C.c_double()
    check(_lib.lib().fistr_b200_fp64_peak(C.byref(v)))
    return v.value


def hbm_copy_peak() -> float:
    v = C.c_double()
    check(_lib.lib().fistr_b200_hbm_copy_peak(C.byref(v)))
    return v.value


def debug_ssor_order(hecMAT, ncolor_in: int = 10):
    """Host-only: the multicolour ordering the GPU SSOR will use (no GPU needed)."""
    A = hecMAT
    perm = np.zeros(max(A.N, 1), np.int32)
    ci = np.zeros(A.N + 2, np.int32)
    nc = C.c_int()
    check(_lib.lib().fistr_b200_debug_ssor_order(A.N, ip(A.indexL), ip(A.itemL), ip(A.indexU), ip(A.itemU),
                                                int(ncolor_in), ip(perm), C.byref(nc), ip(ci)))
    return perm[:A.N], nc.value, ci[: nc.value + 1]


# ---- load / BC helpers mirroring fstr_ass_load (CLOAD) and fstr_AddBC -----------------------------
def fstr_ass_load_cload(hecMESH, hecMAT, cloads):
    """CLOAD part of fstr_ass_load (fistr1/src/analysis/static/fstr_ass_load.f90:59-93): B += value on
    dof of every node of the group."""
    for grp, dof, val in cloads:
        ids = np.asarray(hecMESH.node_group[grp], np.int64)
        np.add.at(hecMAT.B, 3 * (ids - 1) + dof - 1, val)


def boundary_dofs(hecMESH, boundaries):
    """(node, dof, value) triplets in the order fstr_AddBC (fstr_AddBC.f90:54-109) visits them."""
    nodes, dofs, vals = [], [], []
    for grp, d0, d1, val in boundaries:
        ids = np.asarray(hecMESH.node_group[grp], np.int32)
        nd = d1 - d0 + 1
        nodes.append(np.repeat(ids, nd))
        dofs.append(np.tile(np.arange(d0, d1 + 1, dtype=np.int32), len(ids)))
        vals.append(np.full(len(ids) * nd, val))
    if not nodes:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)
    return np.concatenate(nodes), np.concatenate(dofs), np.concatenate(vals)
