"""Host-side mirrors of the two HEC-MW derived types that cross the drop-in boundary.

* ``hecmwST_matrix``      — hecmw1/src/common/hecmw_util_f.F90:433-468
* ``hecmwST_local_mesh``  — hecmw1/src/common/hecmw_util_f.F90:232-381 (fields used on the path)

Arrays keep the reference's memory layout verbatim (1-based node ids stored in
0-based numpy arrays; ``index*`` are ``0:NP`` prefix arrays) so the very same
buffers can be handed to the C-ABI (include/fistr_b200.h) that a Fortran caller
would pass through ISO_C_BINDING.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Iarray / Rarray slot numbers (1-based), hecmw1/src/solver/matrix/hecmw_matrix_misc.f90:89-121
IDX_I_ITER = 1
IDX_I_METHOD = 2
IDX_I_PRECOND = 3
IDX_I_NSET = 4
IDX_I_ITERPREMAX = 5
IDX_I_NREST = 6
IDX_I_SCALING = 7
IDX_I_METHOD2 = 8
IDX_I_ITERLOG = 21
IDX_I_TIMELOG = 22
IDX_I_NCOLOR_IN = 34
IDX_I_MAXRECYCLE_PRECOND = 35
IDX_I_FLAG_CONVERGED = 81
IDX_I_FLAG_DIVERGED = 82
IDX_I_NRECYCLE_PRECOND = 96
IDX_I_FLAG_NUMFACT = 97
IDX_I_FLAG_SYMBFACT = 98
IDX_I_SOLVER_TYPE = 99
IDX_R_RESID = 1
IDX_R_SIGMA_DIAG = 2
IDX_R_SIGMA = 3

# solver return codes, hecmw1/src/solver/init/hecmw_solve_error.f90:9-15
HECMW_SOLVER_ERROR_INCONS_PC = 1001
HECMW_SOLVER_ERROR_ZERO_DIAG = 2001
HECMW_SOLVER_ERROR_ZERO_RHS = 2002
HECMW_SOLVER_ERROR_NOCONV_MAXIT = 3001
HECMW_SOLVER_ERROR_DIVERGE_MAT = 3002
HECMW_SOLVER_ERROR_DIVERGE_PC = 3003

# 361 formulation switch, fistr1/src/lib/m_fstr.f90:68-71
kel361FI, kel361BBAR, kel361IC, kel361FBAR = 1, 2, 3, 4


@dataclass
class hecmwST_matrix:
    N: int = 0
    NP: int = 0
    NPL: int = 0
    NPU: int = 0
    NDOF: int = 3
    indexL: np.ndarray = None
    itemL: np.ndarray = None
    indexU: np.ndarray = None
    itemU: np.ndarray = None
    D: np.ndarray = None
    AL: np.ndarray = None
    AU: np.ndarray = None
    B: np.ndarray = None
    X: np.ndarray = None
    Iarray: np.ndarray = field(default_factory=lambda: np.zeros(100, np.int32))
    Rarray: np.ndarray = field(default_factory=lambda: np.zeros(100, np.float64))

    def fstr_mat_init(self):
        """Solver defaults, fistr1/src/lib/m_fstr.f90:773-805."""
        I, R = self.Iarray, self.Rarray
        I[:] = 0
        R[:] = 0.0
        I[0] = 100; I[1] = 1; I[2] = 1; I[3] = 0; I[4] = 1; I[5] = 10; I[6] = 0
        I[20] = 0; I[21] = 0; I[30] = 0; I[31] = 0; I[32] = 0; I[33] = 10
        I[12] = 3; I[13] = 0; I[34] = 3
        R[0] = 1.0e-8; R[1] = 1.0; R[2] = 0.0; R[3] = 0.1; R[4] = 0.1; R[10] = 1.0e4
        I[95] = 0; I[96] = 1; I[97] = 1; I[98] = 1
        return self

    def hecMAT_init(self):
        """Allocate value arrays, fistr1/src/lib/m_fstr.f90:807-857."""
        self.AL = np.zeros(9 * self.NPL, np.float64)
        self.AU = np.zeros(9 * self.NPU, np.float64)
        self.D = np.zeros(9 * self.NP, np.float64)
        self.B = np.zeros(3 * self.NP, np.float64)
        self.X = np.zeros(3 * self.NP, np.float64)
        return self

    def copy(self):
        m = hecmwST_matrix(self.N, self.NP, self.NPL, self.NPU, self.NDOF,
                           self.indexL, self.itemL, self.indexU, self.itemU)
        for k in ("D", "AL", "AU", "B", "X", "Iarray", "Rarray"):
            v = getattr(self, k)
            setattr(m, k, None if v is None else v.copy())
        return m

    # --- algorithmic SpMV bytes (SURVEY §8d / BASELINE.md §3) ---
    def spmv_bytes(self) -> int:
        nl = int(self.indexL[self.N]); nu = int(self.indexU[self.N]); n = self.N
        return 72 * (n + nl + nu) + 4 * (nl + nu) + 8 * (n + 1) + 48 * n

    def to_scipy(self):
        """Rows 1..N x columns 1..NP as scipy CSR (test helper)."""
        import scipy.sparse as sp
        n, npn = self.N, self.NP
        rows = []
        cols = []
        vals = []
        def add(index, item, A):
            cnt = np.diff(index[: n + 1])
            r = np.repeat(np.arange(n), cnt)
            c = item[: index[n]] - 1
            rows.append(r); cols.append(c); vals.append(A[: 9 * index[n]].reshape(-1, 3, 3))
        add(self.indexL, self.itemL, self.AL)
        add(self.indexU, self.itemU, self.AU)
        rows.append(np.arange(n)); cols.append(np.arange(n)); vals.append(self.D[: 9 * n].reshape(-1, 3, 3))
        r = np.concatenate(rows); c = np.concatenate(cols); v = np.concatenate(vals)
        a3 = np.arange(3)
        rr = np.broadcast_to(3 * r[:, None, None] + a3[None, :, None], v.shape).reshape(-1)
        cc = np.broadcast_to(3 * c[:, None, None] + a3[None, None, :], v.shape).reshape(-1)
        return sp.coo_matrix((v.reshape(-1), (rr, cc)), shape=(3 * n, 3 * npn)).tocsr()


@dataclass
class hecmwST_local_mesh:
    n_node: int = 0
    nn_internal: int = 0
    n_elem: int = 0
    n_dof: int = 3
    node: np.ndarray = None              # 3*n_node, xyz interleaved
    n_elem_type: int = 0
    elem_type_index: np.ndarray = None   # 0:n_elem_type
    elem_type_item: np.ndarray = None
    elem_node_index: np.ndarray = None   # 0:n_elem
    elem_node_item: np.ndarray = None    # 1-based local node ids (FrontISTR-internal node order)
    section_ID: np.ndarray = None        # 1-based
    sections: list = field(default_factory=list)   # [{elemopt361, E, nu}]
    n_neighbor_pe: int = 0
    neighbor_pe: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    import_index: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    import_item: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    export_index: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    export_item: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    my_rank: int = 0
    PETOT: int = 1
    zero: int = 1
    global_node_ID: np.ndarray = None
    node_ID: np.ndarray = None           # (n_node,2): (owner-local id, owner rank), dist meshes only
    node_group: dict = field(default_factory=dict)   # name -> int32 local node ids (1-based)
