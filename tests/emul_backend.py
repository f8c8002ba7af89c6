"""TEST INFRASTRUCTURE: a CPU stand-in for the `gpu` fixture whose assembly / Dirichlet elimination / strain-stress
update / nodal-stress entry points execute the PRODUCT'S OWN CUDA KERNELS on the host (libfistr_b200_emul.so:
assembly.cuh compiled unchanged through csrc/kernel_emul.hpp — warps as 32 lock-step fibers, grid-stride
kernels thread after thread in scrambled order — and launched in the sequence of csrc/api_assembly.inc).  The SpMV runs the
product's kernels too (blocks of fibers); only the Krylov solve is delegated to the oracle (the CG kernels are exercised
separately, on several emulated ranks, by tests/test_p2p_kernels_emulated.py).

Used to run the bodies of GPU tests whose kernels have not run on hardware yet, so that what they assert is checked
against real kernel code before a GPU minute is spent.  It is NOT the product path and proves nothing about launch
configuration, memory residency or performance."""
import ctypes as C

import numpy as np

from frontistr_b200 import hecmw as _real

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


class EmulBackend:
    hecmw_mat_con = staticmethod(_real.hecmw_mat_con)
    boundary_dofs = staticmethod(_real.boundary_dofs)
    fstr_ass_load_cload = staticmethod(_real.fstr_ass_load_cload)

    def __init__(self):
        import __graft_entry__ as g
        self.em = C.CDLL(g.build_emul())
        self.em.fbemul_misaligned_count.restype = C.c_long
        self.opts = {}
        self._A = None
        self._gp = None
        self._keep = None

    # ---- options
    def set_option(self, key, value):
        self.opts[key] = value

    def get_option(self, key):
        if key == "ic_inverse_active":
            return self._icinv()
        if key in ("ic_selftest_ppb", "ic_time_us_smem"):   # the library's kernel selection has no counterpart here
            return 0
        return self.opts.get(key, -1)

    def _icinv(self):
        v = self.opts.get("ic_inverse", -1)
        return v if v in (0, 1, 2) else (1 if v == -2 else 0)      # -2 = "pick": stands for kernel 1 here; default = kernel 0

    # ---- marshalling
    def _mesh_args(self, m):
        k = dict(eti=_i32(m.elem_type_index), ett=_i32(m.elem_type_item), eni=_i32(m.elem_node_index), enn=_i32(m.elem_node_item),
                 sec=_i32(m.section_ID), node=np.ascontiguousarray(m.node, np.float64),
                 form=np.array([s["elemopt361"] for s in m.sections], np.int32),
                 E=np.array([s["E"] for s in m.sections], np.float64), nu=np.array([s["nu"] for s in m.sections], np.float64),
                 mtype=np.array([s.get("mtype", 0) for s in m.sections], np.int32),
                 hyper=np.array([list(s.get("hyper", (0.0, 0.0, 0.0))) for s in m.sections], np.float64).reshape(-1),
                 alpha=np.array([s.get("alpha", 0.0) for s in m.sections], np.float64))
        self._keep = k
        return k, (m.n_node, _d(k["node"]), m.n_elem, m.n_elem_type, _i(k["eti"]), _i(k["ett"]), _i(k["eni"]), _i(k["enn"]),
                   _i(k["sec"]), len(m.sections), _i(k["form"]), _d(k["E"]), _d(k["nu"]))

    def _struct_args(self, A):
        return (A.NP, A.NPL, A.NPU, _i(A.indexL), _i(A.itemL), _i(A.indexU), _i(A.itemU))

    def _vals_to(self, A, vals):
        A.D[:] = vals[: 9 * A.NP]; A.AL[:] = vals[9 * A.NP: 9 * (A.NP + A.NPL)]; A.AU[:] = vals[9 * (A.NP + A.NPL):]

    # ---- residency (host arrays ARE the "device" copy here)
    def mat_structure(self, mesh, A):
        self._A = A
        if A.D is None:
            A.hecMAT_init()

    def mat_upload(self, A, values=True, rhs=True):
        self._A = A

    def mat_download(self, A, values=True, rhs=True, x=False):
        pass

    # ---- assembly
    def fstr_StiffMatrix(self, mesh, A, to_host=True):
        self.mat_structure(mesh, A)
        k, args = self._mesh_args(mesh)
        vals = np.zeros(9 * (A.NP + A.NPL + A.NPU))
        rc = self.em.fbemul_stiffmatrix(*args, *self._struct_args(A), int(self.opts.get("asm_gather", 1) != 0), self._icinv(), _d(vals))
        assert rc == 0 and self.em.fbemul_misaligned_count() == 0
        self._vals_to(A, vals)
        return dict(kernel_ms=0.0, total_ms=0.0)

    def hecmw_mat_ass_bc(self, nodes, dofs, values):
        A = self._A
        n, d, v = _i32(nodes), _i32(dofs), np.ascontiguousarray(values, np.float64)
        vals = np.concatenate([A.D, A.AL, A.AU])
        rc = self.em.fbemul_mat_ass_bc(len(n), _i(n), _i(d), _d(v), A.NP, A.NPL, _i(A.indexL), _i(A.itemL), _i(A.indexU),
                                       _i(A.itemU), _d(vals), _d(A.B))
        assert rc == 0
        self._vals_to(A, vals)

    # ---- solve / SpMV: oracle (these kernels ran on hardware)
    def hecmw_solve(self, mesh, A, device_resident=False):
        from oracle import oracle as orc
        orc.precond_clear()
        r = orc.solve(mesh, A, hist_len=int(A.Iarray[0]))
        orc.precond_clear()
        r["profile"] = np.zeros(4)
        return r

    def hecmw_matvec(self, X, n_internal3):
        """fistr_b200_matvec_33 (single rank): SELL repack + the SpMV kernel of the selected variant, on the host."""
        A = self._A
        x = np.ascontiguousarray(X, np.float64)
        y = np.zeros(3 * A.NP)
        red = C.c_double(0.0)
        v = self.opts.get("spmv_variant", 2)
        rc = self.em.fbemul_spmv(A.N, A.NP, _d(A.D), _d(A.AL), _d(A.AU), _i(A.indexL), _i(A.itemL), _i(A.indexU), _i(A.itemU),
                                 int(v if v in (0, 1, 2) else 2), int(max(self.opts.get("sell_sigma", 0), 0)), 0, _d(x), None, _d(y),
                                 C.byref(red))
        assert rc == 0 and self.em.fbemul_misaligned_count() == 0
        return y[:n_internal3]

    # ---- linear strain / stress recovery
    def fstr_UpdateNewton_linear(self, mesh, unode, dunode, temperature=None, temperature0=None, ref_temp=0.0):
        k, args = self._mesh_args(mesh)
        u = np.ascontiguousarray((0.0 if unode is None else np.asarray(unode)) + (0.0 if dunode is None else np.asarray(dunode)))
        tt = np.ascontiguousarray(temperature, np.float64) if temperature is not None else None
        t0 = np.ascontiguousarray(temperature0, np.float64) if temperature0 is not None else None
        e = np.zeros((mesh.n_elem, 8, 6)); s = np.zeros((mesh.n_elem, 8, 6)); q = np.zeros(3 * mesh.n_node)
        out = dict(STRAIN=np.zeros((mesh.n_node, 6)), STRESS=np.zeros((mesh.n_node, 6)), MISES=np.zeros(mesh.n_node),
                   ESTRAIN=np.zeros((mesh.n_elem, 6)), ESTRESS=np.zeros((mesh.n_elem, 6)), EMISES=np.zeros(mesh.n_elem))
        rc = self.em.fbemul_update_pipeline(*args, _d(k["alpha"]) if tt is not None else None, _d(u), _d(tt), _d(t0),
                                            C.c_double(ref_temp), min(self._icinv(), 1), _d(e), _d(s), _d(q),
                                            *[_d(out[n]) for n in ("STRAIN", "STRESS", "MISES", "ESTRAIN", "ESTRESS", "EMISES")])
        assert rc == 0 and self.em.fbemul_misaligned_count() == 0
        self._gp = (e, s)
        self._nodal = out
        return q

    def fstr_NodalStress3D(self, mesh):
        if getattr(self, "_nodal", None) is not None:
            return self._nodal
        e, s = self._gp                                  # nonlinear state: the nodal kernels alone
        eni, enn = _i32(mesh.elem_node_index), _i32(mesh.elem_node_item)
        out = dict(STRAIN=np.zeros((mesh.n_node, 6)), STRESS=np.zeros((mesh.n_node, 6)), MISES=np.zeros(mesh.n_node),
                   ESTRAIN=np.zeros((mesh.n_elem, 6)), ESTRESS=np.zeros((mesh.n_elem, 6)), EMISES=np.zeros(mesh.n_elem))
        eqf = np.zeros(3 * len(enn)); q = np.zeros(3 * mesh.n_node)
        rc = self.em.fbemul_nodal_pipeline(mesh.n_node, mesh.n_elem, _i(eni), _i(enn), _d(e), _d(s), _d(eqf), 5, _d(q),
                                           *[_d(out[n]) for n in ("STRAIN", "STRESS", "MISES", "ESTRAIN", "ESTRESS", "EMISES")])
        assert rc == 0
        return out

    # ---- nonlinear (total Lagrangian)
    def nl_state_reset(self, mesh):
        self._gp = (np.zeros((mesh.n_elem, 8, 6)), np.zeros((mesh.n_elem, 8, 6)))
        self._nodal = None

    def nl_state_download(self, mesh):
        return self._gp[0].copy(), self._gp[1].copy()

    def fstr_StiffMatrix_NL(self, mesh, A, u):
        self.mat_structure(mesh, A)
        k, args = self._mesh_args(mesh)
        uu = np.ascontiguousarray(u, np.float64)
        vals = np.zeros(9 * (A.NP + A.NPL + A.NPU))
        rc = self.em.fbemul_nl_stiffmatrix(*args, _i(k["mtype"]), _d(k["hyper"]), _d(uu), _d(self._gp[0]), _d(self._gp[1]),
                                           *self._struct_args(A), int(self.opts.get("asm_gather", 1) != 0), _d(vals))
        assert rc == 0 and self.em.fbemul_misaligned_count() == 0
        self._vals_to(A, vals)
        return {}

    def fstr_UpdateNewton(self, mesh, unode, dunode):
        k, args = self._mesh_args(mesh)
        u = np.ascontiguousarray(np.asarray(unode) + np.asarray(dunode))
        q = np.zeros(3 * mesh.n_node)
        rc = self.em.fbemul_nl_update(*args, _i(k["mtype"]), _d(k["hyper"]), _d(u), _d(self._gp[0]), _d(self._gp[1]), _d(q))
        assert rc == 0
        self._nodal = None
        return q

    def fstr_solve_NLGEOM(self, mesh, A, boundary, cload, substeps=10, converg=1.0e-3, max_iter=50):
        """The product's own Newton driver (frontistr_b200.hecmw.fstr_solve_NLGEOM) with its device calls bound to this
        backend for the duration of the call."""
        names = ("mat_structure", "nl_state_reset", "mat_upload", "fstr_StiffMatrix_NL", "fstr_UpdateNewton",
                 "hecmw_mat_ass_bc", "hecmw_solve")
        saved = {n: getattr(_real, n) for n in names}
        try:
            for n in names:
                setattr(_real, n, getattr(self, n))
            return _real.fstr_solve_NLGEOM(mesh, A, boundary, cload, substeps, converg, max_iter)
        finally:
            for n, f in saved.items():
                setattr(_real, n, f)
