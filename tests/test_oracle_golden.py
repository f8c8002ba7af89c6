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


@pytest.mark.parametrize("name", ["F361", "F342"])
def test_oracle_matches_reference_log_thermal_load(name):
    """exF: uniform heating 20 -> 120 degrees of the clamped block (TLOAD_C3D8IC for the default 361 IC element,
    TLOAD_C3 for 342) against the reference's golden logs."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    from frontistr_b200.mshio import apply_solver_card
    mesh, cnt, meta = util.load_golden_mesh(name)
    A = orc.mat_con(mesh)
    orc.stiffmatrix(mesh, A)
    th = meta["thermal"]
    orc.ass_thermal_load(mesh, A, np.full(mesh.n_node, th["temperature"]), np.full(mesh.n_node, th["initial"]),
                         th["alpha"], th["reftemp"])
    n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
    for i in range(len(n)):
        orc.mat_ass_bc(A, int(n[i]), int(d[i]), float(v[i]))
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    orc.precond_clear()
    r = orc.solve(mesh, A)
    orc.precond_clear()
    assert r["error"] == 0 and A.Iarray[80] == 1
    U = A.X.reshape(-1, 3)
    for dd, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        for got, want in ((U[:, dd].max(), gmax), (U[:, dd].min(), gmin)):
            assert abs(got - want) <= meta["tolerance_abs"]
            if abs(want) > 1e-8:
                assert abs(got - want) <= 1.0e-4 * abs(want)


@pytest.mark.parametrize("name", ["B361", "B342"])
def test_oracle_matches_reference_log_face_pressure(name):
    """exB: !DLOAD face pressure (P2 on hexahedra, P3 on 10-node tetrahedra) against the reference's golden logs."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    from frontistr_b200.mshio import apply_solver_card
    mesh, cnt, meta = util.load_golden_mesh(name)
    A = orc.mat_con(mesh)
    orc.stiffmatrix(mesh, A)
    pr = meta["pressure"]
    orc.ass_dload_pressure(mesh, A, pr["elems"], pr["face"], pr["value"])
    n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
    for i in range(len(n)):
        orc.mat_ass_bc(A, int(n[i]), int(d[i]), float(v[i]))
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    orc.precond_clear()
    r = orc.solve(mesh, A)
    orc.precond_clear()
    assert r["error"] == 0 and A.Iarray[80] == 1
    U = A.X.reshape(-1, 3)
    for dd, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        for got, want in ((U[:, dd].max(), gmax), (U[:, dd].min(), gmin)):
            assert abs(got - want) <= meta["tolerance_abs"]
            if abs(want) > 1e-8:
                assert abs(got - want) <= 1.0e-4 * abs(want)


@pytest.mark.parametrize("name", ["C361", "C342", "D361", "D342", "E361", "E342"])
def test_oracle_matches_reference_log_volume_loads(name):
    """exC (body force BZ), exD (gravity), exE (centrifugal): the same K / BC / CG path through three more right-hand
    sides, each against the reference's own golden log."""
    from oracle import oracle as orc
    from frontistr_b200 import hecmw
    from frontistr_b200.mshio import apply_solver_card
    mesh, cnt, meta = util.load_golden_mesh(name)
    A = orc.mat_con(mesh)
    orc.stiffmatrix(mesh, A)
    orc.ass_dload_volume(mesh, A, meta["dload"]["ltype"], meta["dload"]["params"], meta["rho"])
    n, d, v = hecmw.boundary_dofs(mesh, cnt["boundary"])
    for i in range(len(n)):
        orc.mat_ass_bc(A, int(n[i]), int(d[i]), float(v[i]))
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    orc.precond_clear()
    r = orc.solve(mesh, A)
    orc.precond_clear()
    assert r["error"] == 0 and A.Iarray[80] == 1
    U = A.X.reshape(-1, 3)
    for dd, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        assert abs(U[:, dd].max() - gmax) <= meta["tolerance_abs"] and abs(U[:, dd].min() - gmin) <= meta["tolerance_abs"]
        for got, want in ((U[:, dd].max(), gmax), (U[:, dd].min(), gmin)):
            if abs(want) > 1e-8:
                assert abs(got - want) <= 1.0e-4 * abs(want)      # 5 printed digits


@pytest.mark.parametrize("threads", [1, 2])
@pytest.mark.parametrize("name", ["A361", "A342"])
def test_oracle_ssor_matches_reference_log(name, threads):
    """The same golden logs through PRECOND=1 (SSOR) — natural ordering (1 thread) and RCM + multicolour ordering
    (>1 thread, the ordering the GPU path uses): the converged displacements do not depend on the preconditioner,
    so both SSOR restatements are pinned by the reference's own vectors too."""
    from oracle import oracle as orc
    mesh, cnt, meta = util.load_golden_mesh(name)
    cnt = dict(cnt)
    cnt["solver"] = ({**cnt["solver"][0], "PRECOND": "1"}, cnt["solver"][1])
    A = util.oracle_problem(mesh, cnt)
    orc.precond_clear()
    r = orc.solve(mesh, A, ssor_threads=threads)
    orc.precond_clear()
    assert r["error"] == 0 and A.Iarray[80] == 1
    U = A.X.reshape(-1, 3)
    for d, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        assert abs(U[:, d].max() - gmax) <= meta["tolerance_abs"] and abs(U[:, d].min() - gmin) <= meta["tolerance_abs"]


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


@pytest.mark.parametrize("name", ["I361", "I342", "SB1"])
def test_nlgeom_newton_against_reference_logs(name):
    """SURVEY 8(f) rank 2 groundwork: the restated Newton loop (fstr_Newton) with total-Lagrangian B-bar C3D8 /
    C3D10 elements reproduces examples/static/exI/*_correct.log at EVERY one of the 10 sub-steps, and
    examples/static/spring_boundary/SB_correct.log (NLSTATIC + !SPRING, CG + SSOR with iterPREmax = 2) at every one
    of its 5 (nodal displacement max/min, 1e-4 absolute = the reference harness tolerance; in fact all 5 digits)."""
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
    sn, sd, sv = [], [], []
    for grp, dof, val in g.get("spring", []):      # spring_boundary/SB1: !SPRING + SSOR with two sweeps
        ids = list(mesh.node_group[grp])
        sn += ids; sd += [dof] * len(ids); sv += [val] * len(ids)
    nl = g["nlgeom"]
    r = orc.nlgeom_static(mesh, A, bc, (cn, cd, cv), nl["substeps"], nl["converg"], nl["max_iter"], spring=(sn, sd, sv))
    assert r["error"] == 0 and np.all(r["newton_iters"] >= 1) and np.all(r["newton_iters"] <= 5)
    for k, gold in enumerate(g["golden_steps"]):
        U = r["unode_steps"][k].reshape(-1, 3)
        for d, key in enumerate(("U1", "U2", "U3")):
            assert abs(U[:, d].max() - gold[key][0]) <= g["tolerance_abs"], (k, key)
            assert abs(U[:, d].min() - gold[key][1]) <= g["tolerance_abs"], (k, key)
            assert abs(U[:, d].min() - gold[key][1]) <= 6e-5 * max(abs(gold[key][1]), 1e-3)   # printed digits
    # ... and the 26 strain / stress / von-Mises lines of every sub-step: Green-Lagrange strain and 2nd Piola-Kirchhoff
    # stress of the restated total-Lagrangian update through the restated fstr_NodalStress3D
    for k, gold in enumerate(g["golden_stress_steps"]):
        e, s = orc.nlgeom_state(mesh, k + 1)
        util.check_golden_stress(orc.nodal_stress_3d(mesh, e, s), dict(golden_stress=gold, tolerance_abs=g["tolerance_abs"]))


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


@pytest.mark.parametrize("name", ["A361", "A342"])
def test_oracle_iterpremax2_matches_reference_log(name):
    """iterPREmax = 2 (the additive-Schwarz double sweep of hecmw_precond_33_apply, precond_33.f90:86-114) with SSOR
    converges to the same golden displacements.  With DIAG the double sweep is M^-1 (2 - A M^-1), indefinite as soon
    as an eigenvalue of M^-1 A exceeds 2 — on these decks the reference algorithm then stops with
    HECMW_SOLVER_ERROR_DIVERGE_PC (3003), which is asserted as well."""
    from oracle import oracle as orc
    mesh, cnt, meta = util.load_golden_mesh(name)
    cnt = dict(cnt)
    cnt["solver"] = (cnt["solver"][0], [[cnt["solver"][1][0][0], "2"], cnt["solver"][1][1]])
    Ad = util.oracle_problem(mesh, cnt)
    orc.precond_clear()
    assert orc.solve(mesh, Ad)["error"] == 3003 and Ad.Iarray[81] == 1      # flag_diverged
    cnt["solver"] = ({**cnt["solver"][0], "PRECOND": "1"}, cnt["solver"][1])
    A = util.oracle_problem(mesh, cnt)
    assert A.Iarray[4] == 2
    orc.precond_clear()
    r = orc.solve(mesh, A, ssor_threads=2)
    orc.precond_clear()
    assert r["error"] == 0 and A.Iarray[80] == 1
    U = A.X.reshape(-1, 3)
    for d, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        assert abs(U[:, d].max() - gmax) <= meta["tolerance_abs"] and abs(U[:, d].min() - gmin) <= meta["tolerance_abs"]


@pytest.mark.parametrize("form", ["IC", "FI", "BBAR", "TET10"])
def test_patch_test_every_element_formulation(form):
    """Known-answer test independent of any log: a linear displacement field prescribed on the whole boundary of a
    distorted mesh must be reproduced exactly in the interior by every element formulation on the path (constant
    strain patch test) — also exercises hecmw_mat_ass_bc with non-zero prescribed values on many nodes."""
    from oracle import oracle as orc
    from frontistr_b200 import meshgen
    from frontistr_b200.hecmw_types import kel361BBAR, kel361FI, kel361IC
    from frontistr_b200.mshio import apply_solver_card
    if form == "TET10":
        mesh = meshgen.tet10_cube(3)
        xyz = mesh.node.reshape(-1, 3)
    else:
        mesh = meshgen.hex_cube(4, elemopt361={"IC": kel361IC, "FI": kel361FI, "BBAR": kel361BBAR}[form])
        # FrontISTR's B-bar takes the dilatational part from the CENTROID derivative (static_LIB_C3D8.f90:72-74), which
        # is only exactly consistent on parallelepipeds: it passes the patch test on the regular mesh, FI / IC also on
        # a distorted one
        if form != "BBAR":
            mesh = util.distort(mesh)
        xyz = mesh.node.reshape(-1, 3)
    G = np.array([[1.0e-3, 2.0e-4, -3.0e-4], [-1.0e-4, -5.0e-4, 4.0e-4], [2.5e-4, 1.5e-4, 7.0e-4]])
    uex = xyz @ G.T + np.array([1e-3, -2e-3, 5e-4])
    lo, hi = xyz.min(0), xyz.max(0)
    bnd = np.nonzero(np.any((np.abs(xyz - lo) < 1e-12) | (np.abs(xyz - hi) < 1e-12), axis=1))[0]
    A = orc.mat_con(mesh)
    orc.stiffmatrix(mesh, A)
    for n in bnd:
        for d in range(3):
            orc.mat_ass_bc(A, int(n) + 1, d + 1, float(uex[n, d]))
    apply_solver_card(A, util.cube_cnt(tol=1e-12)["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    orc.precond_clear()
    r = orc.solve(mesh, A)
    orc.precond_clear()
    assert r["error"] in (0, 3001) or r["resid2"] < 1e-10
    U = A.X.reshape(-1, 3)
    assert len(bnd) < mesh.n_node
    assert np.abs(U - uex).max() <= 1e-9 * np.abs(uex).max()


LINEAR_DECKS = ["A361", "A342", "B361", "B342", "C361", "C342", "D361", "D342", "E361", "E342", "F361", "F342", "G361", "G342"]


@pytest.mark.parametrize("name", LINEAR_DECKS)
def test_oracle_strain_stress_mises_match_reference_log(name):
    """fstr_UpdateNewton (INFINITE: UpdateST_C3D8IC for the default 361 IC element, UPDATE_C3 for 342) +
    fstr_NodalStress3D restated in the oracle, on the displacement the oracle's own K / load / BC / CG path
    produced: ALL 26 strain / stress / von-Mises max/min lines (nodal and elemental) of every linear golden log
    of examples/static/ex{A..G}, to the 5 printed digits.  Strains are derivatives of u, so this pins the
    displacement FIELD (not only its extrema) and, for exF, the thermal strain handling."""
    from oracle import oracle as orc
    mesh, A, meta, thermal = util.oracle_solve_deck(name)
    TT, T0, alp, ref = thermal if thermal else (None, None, 0.0, 0.0)
    e, s, q = orc.update_linear(mesh, A.X, TT, T0, alp, ref)
    out = orc.nodal_stress_3d(mesh, e, s)
    util.check_golden_stress(out, meta)
