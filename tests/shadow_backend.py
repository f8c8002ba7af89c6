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

    def hecmw_matvec(self, X, n_internal3):
        from oracle import oracle as orc
        return orc.matvec(self._A, np.ascontiguousarray(X, float))[:n_internal3]
