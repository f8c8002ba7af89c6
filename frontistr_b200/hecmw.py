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
    hist = np.zeros(max(int(A.Iarray[0]), 1))
    nh = L.fistr_b200_solve_history(dp(hist), len(hist))
    r2 = C.c_double()
    L.fistr_b200_solve_resid2(C.byref(r2))
    prof = np.zeros(4)
    L.fistr_b200_solve_profile(dp(prof))
    return dict(error=rc, iter=it.value, resid=rs.value, resid2=r2.value, times=times, hist=hist[:nh], profile=prof)


def hecmw_matvec(X: np.ndarray, n_internal3: int) -> np.ndarray:
    x = np.ascontiguousarray(X, np.float64)
    y = np.zeros(n_internal3)
    check(_lib.lib().fistr_b200_matvec_33(dp(x), dp(y)))
    return y


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
