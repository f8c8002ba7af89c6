"""Pins the ORACLE against the reference's own golden vectors: nodal displacement
max/min of examples/static/exA|exG {A361,A342}_correct.log, harness tolerance 1e-4 absolute
(examples/test_FrontISTR.rb:10,225-237), and against an independent sparse direct solve."""
import numpy as np
import pytest

import fb_util as util


@pytest.mark.parametrize("name", ["A361", "A342", "G361", "G342"])
def test_oracle_matches_reference_log(name):
    from oracle import oracle as orc
    mesh, cnt, meta = util.load_golden_mesh(name)
    A = util.oracle_problem(mesh, cnt)
    orc.precond_clear()
    r = orc.solve(mesh, A)
    assert r["error"] == 0 and A.Iarray[80] == 1
    U = A.X.reshape(-1, 3)
    for d, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        assert abs(U[:, d].max() - gmax) <= meta["tolerance_abs"]
        assert abs(U[:, d].min() - gmin) <= meta["tolerance_abs"]
        # the logs print 5 significant digits: match them all
        for got, want in ((U[:, d].max(), gmax), (U[:, d].min(), gmin)):
            if want != 0.0:
                assert abs(got - want) <= 0.6e-4 * abs(want) + 1e-12


@pytest.mark.parametrize("name", ["A361", "A342"])
def test_oracle_vs_direct_solve(name):
    import scipy.sparse.linalg as spl
    from oracle import oracle as orc
    mesh, cnt, _ = util.load_golden_mesh(name)
    A = util.oracle_problem(mesh, cnt)
    A.Rarray[0] = 1e-10   # 1e-12 is below the attainable true residual of the reference algorithm here
    orc.precond_clear()
    r = orc.solve(mesh, A)
    assert r["error"] == 0
    K = A.to_scipy()[:, : 3 * A.N].tocsc()
    u = spl.spsolve(K, A.B[: 3 * A.N])
    assert util.rel_l2(A.X[: 3 * A.N], u) < 1e-9
    assert abs(K - K.T).max() < 1e-9 * abs(K).max()


@pytest.mark.parametrize("name", ["I361", "I342"])
def test_nlgeom_newton_against_reference_logs(name):
    """SURVEY 8(f) rank 2 groundwork: the restated Newton loop (fstr_Newton) with total-Lagrangian B-bar C3D8 /
    C3D10 elements reproduces examples/static/exI/*_correct.log at EVERY one of the 10 sub-steps (nodal
    displacement max/min, 1e-4 absolute = the reference harness tolerance; in fact all 5 printed digits)."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    from frontistr_b200.hecmw_types import kel361BBAR
    from frontistr_b200.mshio import apply_solver_card
    g = util.golden_meta()[name]
    mesh, _, _ = util.load_golden_mesh(g["mesh"], elemopt361=kel361BBAR)   # B-bar: the 361 default under NLGEOM
    A = orc.mat_con(mesh)
    apply_solver_card(A, (g["solver"][0], g["solver"][1]))
    A.Iarray[20] = 0; A.Iarray[21] = 0
    bc = hecmw.boundary_dofs(mesh, [tuple(b) for b in g["boundary"]])
    cn, cd, cv = [], [], []
    for grp, dof, val in g["cload"]:
        ids = list(mesh.node_group[grp])
        cn += ids; cd += [dof] * len(ids); cv += [val] * len(ids)
    nl = g["nlgeom"]
    r = orc.nlgeom_static(mesh, A, bc, (cn, cd, cv), nl["substeps"], nl["converg"], nl["max_iter"])
    assert r["error"] == 0 and np.all(r["newton_iters"] >= 1) and np.all(r["newton_iters"] <= 5)
    for k, gold in enumerate(g["golden_steps"]):
        U = r["unode_steps"][k].reshape(-1, 3)
        for d, key in enumerate(("U1", "U2", "U3")):
            assert abs(U[:, d].max() - gold[key][0]) <= g["tolerance_abs"], (k, key)
            assert abs(U[:, d].min() - gold[key][1]) <= g["tolerance_abs"], (k, key)
            assert abs(U[:, d].min() - gold[key][1]) <= 6e-5 * max(abs(gold[key][1]), 1e-3)   # printed digits


def test_mooney_rivlin_is_the_derivative_of_its_energy():
    """calUpdateElasticMooneyRivlin / calElasticMooneyRivlin restated (Hyperelastic.f90): S = dW/dE and D = dS/dE for
    W = C10 (I1b - 3) + C01 (I2b - 3) + (J - 1)^2 / D, checked by central differences.  The reference ships no golden
    log with this material on the 361/342 B-bar/TL path, so this known-answer test is what pins the restatement."""
    import ctypes as C
    from oracle import oracle as orc
    L = orc.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    c = np.array([0.1486, 0.4849, 0.0789])          # tutorial/03 constants

    def W(e):
        Cm = np.array([[2 * e[0] + 1, e[3], e[5]], [e[3], 2 * e[1] + 1, e[4]], [e[5], e[4], 2 * e[2] + 1]])
        I1 = np.trace(Cm); I2 = 0.5 * (I1 ** 2 - np.trace(Cm @ Cm)); I3 = np.linalg.det(Cm)
        return c[0] * (I1 * I3 ** (-1 / 3) - 3) + c[1] * (I2 * I3 ** (-2 / 3) - 3) + (np.sqrt(I3) - 1) ** 2 / c[2]

    def S(e):
        s = np.zeros(6)
        L.orc_mooney_rivlin_stress(dp(c), dp(np.ascontiguousarray(e)), dp(s))
        return s

    rng = np.random.default_rng(1)
    h = 1e-6
    for trial in range(5):
        F = np.eye(3) + (0.0 if trial == 0 else 0.25) * rng.standard_normal((3, 3))     # a real deformation gradient
        G = 0.5 * (F.T @ F - np.eye(3))
        e0 = np.array([G[0, 0], G[1, 1], G[2, 2], 2 * G[0, 1], 2 * G[1, 2], 2 * G[2, 0]])
        D = np.zeros((6, 6))
        L.orc_mooney_rivlin_tangent(dp(c), dp(e0), dp(D))
        Sfd = np.array([(W(e0 + h * np.eye(6)[i]) - W(e0 - h * np.eye(6)[i])) / (2 * h) for i in range(6)])
        Dfd = np.array([(S(e0 + h * np.eye(6)[j]) - S(e0 - h * np.eye(6)[j])) / (2 * h) for j in range(6)]).T
        assert np.abs(S(e0) - Sfd).max() < 1e-7 and np.abs(D - Dfd).max() < 1e-6 and np.abs(D - D.T).max() < 1e-12
    assert np.abs(S(np.zeros(6))).max() < 1e-14          # stress-free reference configuration


def test_hyperelastic_cylinder_newton_loop_runs():
    """BASELINE.json configs[4] (tutorial/03_hyperelastic_cylinder: 432 C3D8 B-bar, Mooney-Rivlin, 5 sub-steps of a
    prescribed 7.0 end displacement = 26 % stretch, CG + SSOR) through the restated Newton loop: quadratic
    convergence in <= 5 iterations per sub-step, prescribed / symmetry displacements hit exactly, and the outer
    radius contracts as a nearly incompressible bar must (lambda_r ~ lambda^-1/2: -0.70 for an incompressible
    material, a little less with the finite bulk modulus 2/D).  Run-only: no golden log exists in the reference."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    from frontistr_b200.hecmw_types import kel361BBAR
    from frontistr_b200.mshio import apply_solver_card
    g = util.golden_meta()["cylinder"]
    mesh, _, _ = util.load_golden_mesh("cylinder", elemopt361=kel361BBAR)
    mesh.sections = [dict(elemopt361=kel361BBAR, E=1.0, nu=0.3, mtype=1, hyper=tuple(g["mooney_rivlin"]))]
    A = orc.mat_con(mesh)
    apply_solver_card(A, (g["solver"][0], g["solver"][1]))
    A.Iarray[20] = 0; A.Iarray[21] = 0
    bc = hecmw.boundary_dofs(mesh, [tuple(b) for b in g["boundary"]])
    nl = g["nlgeom"]
    r = orc.nlgeom_static(mesh, A, bc, ([], [], []), nl["substeps"], nl["converg"], nl["max_iter"])
    assert r["error"] == 0 and np.all(r["newton_iters"] <= 5)
    U = r["unode"].reshape(-1, 3)
    xyz = mesh.node.reshape(-1, 3)
    assert np.allclose(U[mesh.node_group["LOADS"] - 1, 2], -7.0, atol=1e-8)
    assert np.abs(U[mesh.node_group["FIX"] - 1, 2]).max() < 1e-12
    assert np.abs(U[mesh.node_group["XSYMM"] - 1, 0]).max() < 1e-12 and np.abs(U[mesh.node_group["YSYMM"] - 1, 1]).max() < 1e-12
    radial = (U[:, 0] * xyz[:, 0] + U[:, 1] * xyz[:, 1]) / np.maximum(np.hypot(xyz[:, 0], xyz[:, 1]), 1e-12)
    lam = (26.667 + 7.0) / 26.667
    incompressible = 6.413 * (1.0 - lam ** -0.5)
    assert radial.max() <= 1e-9 and 0.9 * incompressible < -radial.min() < 1.0 * incompressible
