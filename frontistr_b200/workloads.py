"""The benchmark workloads BASELINE.json names, as data: control-card content of the synthetic cube load
case and the headline configurations.  Used by bench.py, tools/ and the tests alike.

Load case of every synthetic cube (mirrors examples/static/exA): node group FIX = face x=0 fixed in dofs 1-3,
CLOAD -1 in dof 3 on every node of node group CL1 = face x=L; `!SOLVER, METHOD=CG, PRECOND=<p>` with
`<maxit>, <iterPREmax>` / `<resid>, 1.0, 0.0` as in fistr1's .cnt files (fstr_ctrl_common.f90: fstr_ctrl_get_SOLVER).
"""
from __future__ import annotations

# configs of BASELINE.json that are synthetic meshes: name -> (kind, edge in cells, PRECOND)
CONFIGS = {
    "H150": dict(kind="hex", n=150, precond=3, what="configs[1]: C3D8 hex cube 150^3 elements, CG + DIAG, 1 GPU"),
    "H200": dict(kind="hex", n=200, precond=3, what="configs[2]: C3D8 hex cube 200^3 elements, partitioned over 2/4/8 GPUs"),
    "T107": dict(kind="tet10", n=107, precond=3, what="configs[3]: C3D10 tetra cube, 6*107^3 elements, ~30 M DOF, block-Jacobi"),
}


def cube_cnt(tol=1.0e-8, precond=3, maxit=10000, iterpremax=1):
    """SURVEY §8d load case: face x=0 fixed in 1-3, CLOAD -1 in dof 3 on every node of face x=L."""
    return dict(boundary=[("FIX", 1, 3, 0.0)], cload=[("CL1", 3, -1.0)],
                solver=({"METHOD": "CG", "PRECOND": str(precond), "ITERLOG": "NO", "TIMELOG": "NO"},
                        [[str(maxit), str(iterpremax)], [repr(tol), "1.0", "0.0"]]))


def hex_workload_name(n, precond=3):
    pc = {1: "SSOR", 2: "SSOR", 3: "DIAG(3x3 block-Jacobi)", 10: "BILU(0)"}.get(precond, "PRECOND=%d" % precond)
    return {150: "H150", 200: "H200"}.get(n, "H%d" % n) + ": C3D8(IC) hex cube %d^3 el, %d DOF, CG+%s" % (n, 3 * (n + 1) ** 3, pc)
