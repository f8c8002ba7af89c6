"""TEST INFRASTRUCTURE: a CPU stand-in for the `gpu` fixture (frontistr_b200.hecmw bound to a B200), backed by the
oracle, used ONLY to dry-run the bodies of GPU tests on the CPU so that a Python-level mistake in a test that has not
run on hardware yet (wrong attribute, wrong argument order, bad fixture key) is caught here instead of turning the
round-end GPU suite red.  It proves nothing about the product: every number it returns is the oracle's."""
import numpy as np

from frontistr_b200 import hecmw as _real


class ShadowBackend:
    def __init__(self):
        self._A = None
        self._mesh = None
        self.opts = {}

    # host-only entry points of the real module work without a GPU
    hecmw_mat_con = staticmethod(_real.hecmw_mat_con)
    boundary_dofs = staticmethod(_real.boundary_dofs)
    fstr_ass_load_cload = staticmethod(_real.fstr_ass_load_cload)

    def set_option(self, key, value):
        self.opts[key] = value

    def get_option(self, key):
        if key == "ic_inverse_active":      # the library reports which calInverse kernel it used
            v = self.opts.get("ic_inverse", -1)
            return v if v in (0, 1, 2) else (1 if v == -2 else 0)      # -2 = "pick": stands for kernel 1 here; default = kernel 0
        if key in ("ic_selftest_ppb", "ic_time_us_smem"):   # the library's kernel selection has no counterpart here
            return 0
        return self.opts.get(key, -1)

    def mat_structure(self, mesh, A):
        self._mesh, self._A = mesh, A

    def fstr_StiffMatrix(self, mesh, A, to_host=True):
        from oracle import oracle as orc
        if A.D is None:
            A.hecMAT_init()
        orc.stiffmatrix(mesh, A)
        self._mesh, self._A = mesh, A
        return dict(kernel_ms=0.0, total_ms=0.0)

    def mat_upload(self, A, values=True, rhs=True):
        self._A = A

    def mat_download(self, A, values=True, rhs=True, x=False):
        pass

    def hecmw_mat_ass_bc(self, nodes, dofs, values):
        from oracle import oracle as orc
        for n, d, v in zip(nodes, dofs, values):
            orc.mat_ass_bc(self._A, int(n), int(d), float(v))

    def hecmw_solve(self, mesh, A, device_resident=False):
        from oracle import oracle as orc
        orc.precond_clear()
        r = orc.solve(mesh, A, hist_len=int(A.Iarray[0]))
        orc.precond_clear()
        r["profile"] = np.zeros(4)
        return r

    def fstr_UpdateNewton_linear(self, mesh, unode, dunode, temperature=None, temperature0=None, ref_temp=0.0):
        from oracle import oracle as orc
        u = (0.0 if unode is None else np.asarray(unode)) + (0.0 if dunode is None else np.asarray(dunode))
        eo = so = qo = None
        for isec, sct in enumerate(mesh.sections):      # the oracle takes ONE expansion coefficient per run
            e, s, q = orc.update_linear(mesh, u, temperature, temperature0, sct.get("alpha", 0.0), ref_temp)
            if eo is None:
                eo, so, qo = e.copy(), s.copy(), q
            sel = np.asarray(mesh.section_ID) == isec + 1
            eo[sel], so[sel] = e[sel], s[sel]
        self._gp = (mesh, eo, so)
        return qo

    def fstr_NodalStress3D(self, mesh):
        from oracle import oracle as orc
        return orc.nodal_stress_3d(self._gp[0], self._gp[1], self._gp[2])

    def hecmw_matvec(self, X, n_internal3):
        from oracle import oracle as orc
        return orc.matvec(self._A, np.ascontiguousarray(X, float))[:n_internal3]


def patch_nonlinear(monkeypatch):
    """Replaces the five device calls used by frontistr_b200.hecmw.fstr_solve_NLGEOM with oracle-backed stand-ins
    (module-level patch: fstr_solve_NLGEOM calls them as module functions).  Returns the hecmw module, usable as `gpu`."""
    import ctypes as C

    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    L = orc.lib()
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    st = {}

    def _elems(m):
        for it in range(m.n_elem_type):
            et = int(m.elem_type_item[it])
            for e in range(int(m.elem_type_index[it]), int(m.elem_type_index[it + 1])):
                yield e, et, np.ascontiguousarray(m.elem_node_item[m.elem_node_index[e]: m.elem_node_index[e + 1]], np.int32)

    def _sec(m, e, et):
        sec = orc.c_sections(m)
        s = sec[int(m.section_ID[e]) - 1]
        return sec, s, int(et == 361 and s.elemopt361 == 2)

    def nl_state_reset(m):
        st["e"] = np.zeros((m.n_elem, 8, 6)); st["s"] = np.zeros((m.n_elem, 8, 6))

    def mat_structure(m, A):
        st["A"] = A
        if A.D is None:
            A.hecMAT_init()

    def stiff_nl(m, A, u):
        A.D[:] = 0; A.AL[:] = 0; A.AU[:] = 0
        cm = orc.c_matrix(A)
        xyz = m.node.reshape(-1, 3)
        for e, et, nod in _elems(m):
            keep, s, bbar = _sec(m, e, et)
            ec = np.ascontiguousarray(xyz[nod - 1]); ue = np.ascontiguousarray(np.asarray(u).reshape(-1, 3)[nod - 1])
            K = np.zeros((3 * len(nod), 3 * len(nod)))
            L.orc_STF_C3_TL(et, len(nod), d(ec), d(ue), d(st["e"][e]), d(st["s"][e]), C.byref(s), bbar, d(K))
            L.orc_mat_ass_elem(C.byref(cm.s), len(nod), nod.ctypes.data_as(C.POINTER(C.c_int)), d(K), 3 * len(nod))
        return {}

    def update(m, unode, dunode):
        qf = np.zeros(3 * m.n_node)
        xyz = m.node.reshape(-1, 3)
        for e, et, nod in _elems(m):
            keep, s, bbar = _sec(m, e, et)
            ec = np.ascontiguousarray(xyz[nod - 1])
            u = np.ascontiguousarray(unode.reshape(-1, 3)[nod - 1]); du = np.ascontiguousarray(dunode.reshape(-1, 3)[nod - 1])
            q = np.zeros(3 * len(nod))
            L.orc_UPDATE_C3_TL(et, len(nod), d(ec), d(u), d(du), C.byref(s), bbar, d(q), d(st["e"][e]), d(st["s"][e]))
            np.add.at(qf.reshape(-1, 3), nod - 1, q.reshape(-1, 3))
        return qf

    def ass_bc(nodes, dofs, vals):
        for n, dof, v in zip(nodes, dofs, vals):
            orc.mat_ass_bc(st["A"], int(n), int(dof), float(v))

    def solve(m, A, device_resident=False):
        orc.precond_clear()
        r = orc.solve(m, A)
        orc.precond_clear()
        return r

    monkeypatch.setattr(hecmw, "mat_structure", mat_structure)
    monkeypatch.setattr(hecmw, "nl_state_reset", nl_state_reset)
    monkeypatch.setattr(hecmw, "mat_upload", lambda A, values=True, rhs=True: None)
    monkeypatch.setattr(hecmw, "fstr_StiffMatrix_NL", stiff_nl)
    monkeypatch.setattr(hecmw, "fstr_UpdateNewton", update)
    monkeypatch.setattr(hecmw, "fstr_NodalStress3D", lambda m: orc.nodal_stress_3d(m, st["e"], st["s"]))
    monkeypatch.setattr(hecmw, "hecmw_mat_ass_bc", ass_bc)
    monkeypatch.setattr(hecmw, "hecmw_solve", solve)
    return hecmw
