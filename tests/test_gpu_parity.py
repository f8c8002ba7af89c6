"""GPU parity tests proper: every call goes through the C-ABI (frontistr_b200/libfistr_b200.so) and is
compared with the oracle on the same inputs.  Integer/index results must be bit-exact; floating-point
tolerances are written next to each assertion."""
import os

import numpy as np
import pytest

import fb_util as util
from frontistr_b200 import meshgen
from frontistr_b200.hecmw_types import kel361BBAR, kel361FI, kel361IC

pytestmark = pytest.mark.gpu


def _meshes():
    return {
        "hex5_ic": lambda: util.distort(meshgen.hex_cube(5)),
        "hex4_fi": lambda: util.distort(meshgen.hex_cube(4, elemopt361=kel361FI)),
        "hex4_bbar": lambda: util.distort(meshgen.hex_cube(4, elemopt361=kel361BBAR)),
        "tet3": lambda: meshgen.tet10_cube(3),
        "hex1": lambda: meshgen.hex_cube(1),                       # a single element
        "hex_ragged": lambda: meshgen.hex_cube(3, nx=7, ny=2, nz=1),  # rows of very different length
    }


@pytest.mark.parametrize("name", list(_meshes()))
def test_assembly_matches_oracle(gpu, name):
    """fstr_StiffMatrix + hecmw_mat_ass_elem.  K is compared with 1e-12 relative (per array, L2): the
    reference itself sums with order-free atomics, the GPU likewise."""
    from oracle import oracle as orc
    mesh = _meshes()[name]()
    cnt = util.cube_cnt()
    A = util.product_problem(gpu, mesh, cnt, to_host=True, apply_bc=False)
    Ao = util.oracle_problem(mesh, cnt, apply_bc=False)
    for k in ("indexL", "itemL", "indexU", "itemU"):
        assert np.array_equal(getattr(A, k), getattr(Ao, k))
    scale = np.abs(Ao.D).max()
    for k in ("D", "AL", "AU"):
        a, b = getattr(A, k), getattr(Ao, k)
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 1e-12 * scale, k


@pytest.mark.parametrize("name", ["hex5_ic", "tet3", "hex_ragged"])
def test_assembly_gather_equals_scatter_and_is_reproducible(gpu, name):
    """The deterministic gather assembly (element matrices stored, each block sums its contributions in
    ascending element order) against the atomic scatter-add path: same K to rounding; the gather result is
    bit-identical from run to run."""
    mesh = _meshes()[name]()
    cnt = util.cube_cnt()
    Ks = {}
    for mode in (1, 0, 1):
        gpu.set_option("asm_gather", mode)
        A = util.product_problem(gpu, mesh, cnt, to_host=True, apply_bc=False)
        Ks.setdefault(mode, []).append(np.concatenate([A.D, A.AL, A.AU]))
    gpu.set_option("asm_gather", 1)
    assert np.array_equal(Ks[1][0], Ks[1][1])
    scale = np.abs(Ks[0][0]).max()
    assert np.abs(Ks[1][0] - Ks[0][0]).max() <= 1e-13 * scale


def test_assembly_golden_decks(gpu):
    for name in ("A361", "A342"):
        mesh, cnt, _ = util.load_golden_mesh(name)
        A = util.product_problem(gpu, mesh, cnt, to_host=True, apply_bc=True)
        Ao = util.oracle_problem(mesh, cnt, apply_bc=True)
        scale = np.abs(Ao.D).max()
        for k in ("D", "AL", "AU", "B"):
            assert np.abs(getattr(A, k) - getattr(Ao, k)).max() <= 1e-12 * scale, (name, k)


def test_bc_elimination_matches_oracle(gpu):
    """hecmw_mat_ass_bc incl. non-zero prescribed values, several constraints per node and neighbours."""
    mesh = util.distort(meshgen.hex_cube(4))
    cnt = util.cube_cnt()
    cnt["boundary"] = [("FIX", 1, 3, 0.0), ("CL1", 1, 2, 0.125), ("CL1", 3, 3, -0.5)]
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    Ao = util.oracle_problem(mesh, cnt)
    scale = np.abs(Ao.D).max()
    for k in ("D", "AL", "AU"):
        assert np.abs(getattr(A, k) - getattr(Ao, k)).max() <= 1e-12 * scale, k
    assert np.abs(A.B - Ao.B).max() <= 1e-12 * max(scale, np.abs(Ao.B).max())
    # structural properties are exact: zeroed rows/columns, unit diagonal
    n, d, _ = gpu.boundary_dofs(mesh, cnt["boundary"])
    Dm = A.D.reshape(-1, 3, 3)
    assert np.all(Dm[n - 1, d - 1, d - 1] == 1.0)
    assert np.array_equal(A.D == 0.0, Ao.D == 0.0) and np.array_equal(A.AL == 0.0, Ao.AL == 0.0)


def test_bc_elimination_same_dof_named_twice(gpu):
    """Overlapping node groups make decks name one DOF in several !BOUNDARY lines, possibly with different values.  The
    reference eliminates one constraint at a time (fstr_AddBC.f90:54-109): the second visit of a DOF finds its column
    already zeroed, so nothing more goes to the right-hand side and only B(c) is overwritten — eliminate with the FIRST
    value, end up with the LAST.  The batched device elimination must give the same matrix and right-hand side
    (found by tests/test_host_api_fuzz.py)."""
    mesh = util.distort(meshgen.hex_cube(4))
    cnt = util.cube_cnt()
    cnt["boundary"] = [("FIX", 1, 3, 0.0), ("CL1", 1, 2, 0.125), ("CL1", 2, 3, -0.5), ("FIX", 2, 2, 0.25), ("CL1", 3, 3, 0.75)]
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    Ao = util.oracle_problem(mesh, cnt)
    scale = np.abs(Ao.D).max()
    for k in ("D", "AL", "AU"):
        assert np.abs(getattr(A, k) - getattr(Ao, k)).max() <= 1e-12 * scale, k
    assert np.abs(A.B - Ao.B).max() <= 1e-12 * max(scale, np.abs(Ao.B).max())
    fix, cl = mesh.node_group["FIX"], mesh.node_group["CL1"]
    assert np.all(A.B[3 * (fix - 1) + 1] == 0.25) and np.all(A.B[3 * (cl - 1) + 2] == 0.75) and np.all(A.B[3 * (cl - 1) + 1] == -0.5)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name", ["hex5_ic", "tet3", "hex1", "hex_ragged"])
def test_spmv_bitwise(gpu, name, variant):
    """hecmw_matvec_33: one thread per row accumulating D, L ascending, U ascending with the same fma
    chain as the oracle => bit-exact y (integer-like parity for a floating-point kernel)."""
    from oracle import oracle as orc
    mesh = _meshes()[name]()
    Ao = util.oracle_problem(mesh, util.cube_cnt())
    gpu.set_option("spmv_variant", variant)
    gpu.mat_structure(mesh, Ao)
    gpu.mat_upload(Ao)
    x = np.sin(np.arange(1, 3 * Ao.NP + 1, dtype=float))
    y = gpu.hecmw_matvec(x, 3 * Ao.N)
    yo = orc.matvec(Ao, x)[: 3 * Ao.N]
    gpu.set_option("spmv_variant", 2)
    assert np.array_equal(y, yo)


@pytest.mark.parametrize("name", ["hex5_ic", "tet3"])
def test_matresid_and_rel_resid_L2(gpu, name):
    """hecmw_matresid / hecmw_rel_resid_L2 (las_33.f90:358-415) as entry points of their own: R = B - A X bit-identical to
    B minus the oracle's matvec (same accumulation order, same subtraction), the relative residual to 1e-14; with the
    resident right-hand side and with one passed in; zero right-hand side -> norm 1."""
    from oracle import oracle as orc
    mesh = _meshes()[name]()
    cnt = util.cube_cnt()
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    n3 = 3 * A.N
    x = np.cos(0.7 * np.arange(1, 3 * A.NP + 1, dtype=float)); b2 = np.sin(np.arange(1, 3 * A.NP + 1, dtype=float))
    y = orc.matvec(A, x)                 # the oracle's SpMV on the very matrix that is resident on the device
    for b_arg, b in ((None, A.B), (b2, b2)):
        r, rel = gpu.hecmw_matresid(x, b_arg, n3)
        assert np.array_equal(r, b[:n3] - y[:n3])
        want = np.sqrt(np.dot(r, r) / np.dot(b[:n3], b[:n3]))
        assert abs(rel - want) <= 1e-14 * want
    r, rel = gpu.hecmw_matresid(x, np.zeros(3 * A.NP), n3)
    assert np.array_equal(r, -y[:n3]) and abs(rel - np.sqrt(np.dot(y[:n3], y[:n3]))) <= 1e-14 * rel


def test_precond_diag_bitwise(gpu):
    from oracle import oracle as orc
    mesh = util.distort(meshgen.hex_cube(5))
    Ao = util.oracle_problem(mesh, util.cube_cnt())
    gpu.mat_structure(mesh, Ao); gpu.mat_upload(Ao)
    gpu.hecmw_precond_setup(3, 1.0)
    z = np.cos(np.arange(3 * Ao.NP, dtype=float))
    zo = z.copy()
    gpu.hecmw_precond_apply(z)
    orc.Precond(Ao, precond=3).apply(zo)
    assert np.array_equal(z[: 3 * Ao.N], zo[: 3 * Ao.N])
    gpu.hecmw_precond_clear()


@pytest.mark.parametrize("name", ["hex5_ic", "tet3"])
def test_precond_ssor_bitwise(gpu, name):
    """Multicolour SSOR: same ordering (checked), same per-row accumulation order => bit-exact sweep."""
    from oracle import oracle as orc
    mesh = _meshes()[name]()
    Ao = util.oracle_problem(mesh, util.cube_cnt())
    gpu.mat_structure(mesh, Ao); gpu.mat_upload(Ao)
    gpu.hecmw_precond_setup(1, 1.0, 10)
    pc = orc.Precond(Ao, precond=1, ncolor_in=10, nthreads=2)
    nc, perm = gpu.precond_info(Ao.N)
    assert nc == pc.ncolor and np.array_equal(perm, pc.perm())
    z = np.cos(np.arange(3 * Ao.NP, dtype=float)); z[3 * Ao.N:] = 0
    zo = z.copy()
    gpu.hecmw_precond_apply(z)
    pc.apply(zo)
    assert np.array_equal(z[: 3 * Ao.N], zo[: 3 * Ao.N])
    gpu.hecmw_precond_clear()


@pytest.mark.parametrize("name", ["A361", "A342"])
def test_solve_golden_decks(gpu, name):
    """End-to-end through the C-ABI on the reference's own decks: displacement max/min of the golden
    logs (1e-4 absolute, the reference harness tolerance) and iteration count within +-2 of the oracle."""
    from oracle import oracle as orc
    mesh, cnt, meta = util.load_golden_mesh(name)
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    r = gpu.hecmw_solve(mesh, A)
    Ao = util.oracle_problem(mesh, cnt)
    orc.precond_clear()
    ro = orc.solve(mesh, Ao)
    assert r["error"] == 0 and A.Iarray[80] == 1 and A.Iarray[81] == 0
    # A361 (70 iterations): +-2 of the oracle.  A342 needs ~720 iterations for 1575 DOF: CG has long lost
    # orthogonality there and the count depends on the summation order of the dot products at the 1e-16
    # level - the REFERENCE ALGORITHM ITSELF gives 706..734 depending on how the same system is decomposed into
    # subdomains (every decomposition adds the same terms in another order).  The window is not a constant of this
    # test: it is computed here, deterministically (one oracle thread), over 25 decompositions - RCB along 11 axis
    # strings and KMETIS / PMETIS with 2..8 domains (tests/fb_util.py: oracle_iteration_spread; also
    # tests/test_oracle_units.py::test_iteration_count_spread_of_the_reference_algorithm).  The GPU count must lie
    # inside the reference's own spread (+-2); the residual history is pinned while the recurrences still agree.
    # 26 executions sample the reference's range, they do not bound it: the oracle itself, run with 16 OpenMP threads on
    # A342, has given 739 against a sampled 706..734.  The window is therefore the sampled range extended by its own width
    # on either side (never less than +-2): zero-width samples (A361: every decomposition gives the same count) keep the
    # north_star gate, chaotic ones get the room the reference algorithm takes for itself.
    counts = [ro["iter"]] + util.oracle_iteration_spread(mesh, cnt)
    slack = max(2, max(counts) - min(counts))
    assert min(counts) - slack <= r["iter"] <= max(counts) + slack, (r["iter"], counts)
    if name == "A361":
        assert abs(r["iter"] - ro["iter"]) <= 2
    orc.precond_clear()
    roh = orc.solve(mesh, util.oracle_problem(mesh, cnt), hist_len=64)
    assert np.allclose(r["hist"][:20], roh["hist"][:20], rtol=1e-8)
    assert r["resid2"] < A.Rarray[0]
    U = A.X.reshape(-1, 3)
    for d, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        assert abs(U[:, d].max() - gmax) <= 1e-4 and abs(U[:, d].min() - gmin) <= 1e-4


@pytest.mark.parametrize("name", ["B361", "B342", "C361", "C342", "D361", "D342", "E361", "E342", "F361", "F342"])
def test_solve_golden_decks_distributed_loads(gpu, name):
    """exB (face pressure), exC (BZ), exD (GRAV), exE (CENT), exF (thermal): assembly, BC elimination and CG on the
    device; the right-hand side comes from the restated load routines DL_C3 / TLOAD_C3 / TLOAD_C3D8IC (caller side of
    the path, test infrastructure); displacement max/min of the reference's golden logs."""
    from oracle import oracle as orc
    from frontistr_b200.mshio import apply_solver_card
    mesh, cnt, meta = util.load_golden_mesh(name)
    A = gpu.hecmw_mat_con(mesh)
    A.fstr_mat_init().hecMAT_init()
    gpu.mat_structure(mesh, A)
    gpu.fstr_StiffMatrix(mesh, A, to_host=False)
    if "dload" in meta:
        orc.ass_dload_volume(mesh, A, meta["dload"]["ltype"], meta["dload"]["params"], meta["rho"])
    elif "pressure" in meta:
        orc.ass_dload_pressure(mesh, A, meta["pressure"]["elems"], meta["pressure"]["face"], meta["pressure"]["value"])
    else:
        th = meta["thermal"]
        orc.ass_thermal_load(mesh, A, np.full(mesh.n_node, th["temperature"]), np.full(mesh.n_node, th["initial"]),
                             th["alpha"], th["reftemp"])
    gpu.mat_upload(A, values=False, rhs=True)
    n, d, v = gpu.boundary_dofs(mesh, cnt["boundary"])
    gpu.hecmw_mat_ass_bc(n, d, v)
    apply_solver_card(A, cnt["solver"])
    A.Iarray[20] = 0; A.Iarray[21] = 0
    r = gpu.hecmw_solve(mesh, A, device_resident=True)
    assert r["error"] == 0 and A.Iarray[80] == 1 and r["resid2"] < A.Rarray[0]
    U = A.X.reshape(-1, 3)
    for dd, key in enumerate(("U1", "U2", "U3")):
        gmax, gmin = meta["golden"][key]
        assert abs(U[:, dd].max() - gmax) <= 1e-4 and abs(U[:, dd].min() - gmin) <= 1e-4       # the harness tolerance
        for got, want in ((U[:, dd].max(), gmax), (U[:, dd].min(), gmin)):                    # ... and the printed digits
            if abs(want) > 1e-8:
                assert abs(got - want) <= 1.0e-4 * abs(want)


@pytest.mark.parametrize("precond,fused", [(3, 1), (3, 0), (1, 0)])
@pytest.mark.parametrize("meshname", ["hex12", "tet4"])
def test_solve_matches_oracle(gpu, meshname, precond, fused):
    """CG parity (north_star): iterations within +-2, final true residual below RESID, displacements within
    1e-9 relative L2 of the oracle.  Run at RESID=1e-10 (1e-8 leaves too much slack for a 1e-9 comparison;
    1e-12 is below the attainable true residual of the reference algorithm on these meshes)."""
    from oracle import oracle as orc
    mesh = util.distort(meshgen.hex_cube(12)) if meshname == "hex12" else meshgen.tet10_cube(4)
    cnt = util.cube_cnt(tol=1e-10, precond=precond)
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    gpu.set_option("cg_fused", fused)
    r = gpu.hecmw_solve(mesh, A)
    gpu.set_option("cg_fused", 1)
    Ao = util.oracle_problem(mesh, cnt)
    orc.precond_clear()
    ro = orc.solve(mesh, Ao, hist_len=10000, ssor_threads=2)
    assert r["error"] == 0 and ro["error"] == 0
    assert r["iter"] > 50                      # exercises the every-50-iterations true-residual branch
    assert abs(r["iter"] - ro["iter"]) <= 2
    assert r["resid2"] < 1e-10 and A.Iarray[80] == 1
    assert util.rel_l2(A.X[: 3 * A.N], Ao.X[: 3 * A.N]) <= 1e-9
    # residual histories: identical to 1e-8 early on, then rounding differences of the dot products grow
    n = min(len(r["hist"]), len(ro["hist"])) - 3
    assert np.allclose(r["hist"][:40], ro["hist"][:40], rtol=1e-8)
    assert np.allclose(r["hist"][: n // 2], ro["hist"][: n // 2], rtol=1e-3)
    # tail: same decay, rounding has decorrelated.  How much is not a constant of this test: the reference algorithm's
    # own histories on 2 / 2 / 4 / 8 subdomains of the same system drift 2-5 % apart from the undivided run over these
    # iterations (DIAG; the subdomain-local SSOR changes the preconditioner itself and is compared on the first half only)
    if precond == 3:
        own = util.oracle_history_spread(mesh, cnt, ro["hist"], n)
        dev = float(np.max(np.abs(r["hist"][:n] / ro["hist"][:n] - 1.0)))
        assert dev <= 3.0 * own + 1e-6, (dev, own)


def test_solve_iterpremax2(gpu):
    """iterPREmax=2 exercises the additive-Schwarz multi-sweep branch (precond_33.f90:110-113)."""
    from oracle import oracle as orc
    mesh = util.distort(meshgen.hex_cube(8))
    cnt = util.cube_cnt(tol=1e-9, precond=1, iterpremax=2)
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    r = gpu.hecmw_solve(mesh, A)
    Ao = util.oracle_problem(mesh, cnt)
    orc.precond_clear()
    ro = orc.solve(mesh, Ao, ssor_threads=2)
    assert r["error"] == 0 and abs(r["iter"] - ro["iter"]) <= 2
    assert util.rel_l2(A.X[: 3 * A.N], Ao.X[: 3 * A.N]) <= 1e-8


def test_solve_device_resident_pipeline(gpu):
    """stiffmatrix -> bc -> solve without the matrix ever visiting the host == host-array pipeline."""
    mesh = util.distort(meshgen.hex_cube(8))
    cnt = util.cube_cnt(tol=1e-9)
    A = util.product_problem(gpu, mesh, cnt, to_host=False)
    r = gpu.hecmw_solve(mesh, A, device_resident=True)
    x_dev = A.X.copy()
    B = util.product_problem(gpu, mesh, cnt, to_host=True)
    r2 = gpu.hecmw_solve(mesh, B)
    assert r["error"] == 0 and r2["error"] == 0 and abs(r["iter"] - r2["iter"]) <= 1
    assert util.rel_l2(x_dev, B.X) <= 1e-9


def test_solver_status_codes(gpu):
    """Error/status contract (hecmw_solve_error.f90:9-15, Iterative.f90:145-170)."""
    mesh = meshgen.hex_cube(4)
    cnt = util.cube_cnt(tol=1e-8)
    A = util.product_problem(gpu, mesh, cnt, to_host=True)
    # zero RHS -> 2002, X = 0
    Z = A.copy(); Z.B[:] = 0.0; Z.X[:] = 1.0
    r = gpu.hecmw_solve(mesh, Z)
    assert r["error"] == 2002 and np.all(Z.X == 0.0)
    # MAXIT too small -> 3001, not converged
    M = A.copy(); M.Iarray[0] = 5
    r = gpu.hecmw_solve(mesh, M)
    assert r["error"] == 3001 and r["iter"] == 6 and M.Iarray[80] == 0
    # negative definite matrix -> p.q <= 0 -> 3002 and flag_diverged
    Nn = A.copy(); Nn.D *= -1; Nn.AL *= -1; Nn.AU *= -1
    r = gpu.hecmw_solve(mesh, Nn)
    assert r["error"] in (3002, 3003) and Nn.Iarray[81] == 1
    # NIER set to "unlimited" must not cost a history buffer of that size
    Big = A.copy(); Big.Iarray[0] = 2 ** 31 - 1
    r = gpu.hecmw_solve(mesh, Big)
    assert r["error"] == 0 and Big.Iarray[80] == 1 and len(r["hist"]) == r["iter"]
    # zero diagonal entry -> 2001
    Zd = A.copy(); Zd.D[4] = 0.0
    r = gpu.hecmw_solve(mesh, Zd)
    assert r["error"] == 2001
    # METHOD other than CG is not on this path -> 1001
    Me = A.copy(); Me.Iarray[1] = 2
    r = gpu.hecmw_solve(mesh, Me)
    assert r["error"] == 1001
    # an option of the reference path that is not implemented is refused, not ignored: USEJAD (Iarray(33)); and after a
    # divergence METHOD2 > 1 would switch to another Krylov method (Iterative.f90:150-153)
    Uj = A.copy(); Uj.Iarray[32] = 1
    assert gpu.hecmw_solve(mesh, Uj)["error"] == 1001
    M2 = Nn.copy(); M2.Iarray[7] = 2; M2.X[:] = 0.0
    assert gpu.hecmw_solve(mesh, M2)["error"] == 1001
    M2ok = A.copy(); M2ok.Iarray[7] = 2            # METHOD2 is only looked at after a divergence
    assert gpu.hecmw_solve(mesh, M2ok)["error"] == 0
    # same call sequence with the oracle gives the same codes
    from oracle import oracle as orc
    for mat, want in ((A.copy(), 0), (Uj.copy(), 1001), (M2.copy(), 1001)):
        orc.precond_clear()
        mat.X[:] = 0.0
        assert orc.solve(mesh, mat)["error"] == want


def test_matrix_dump_files(gpu):
    import pathlib
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        old = os.environ.get("FISTR_B200_DUMP_DIR")
        os.environ["FISTR_B200_DUMP_DIR"] = tmp
        try:
            _matrix_dump_files(gpu, pathlib.Path(tmp))
        finally:
            if old is None:
                del os.environ["FISTR_B200_DUMP_DIR"]
            else:
                os.environ["FISTR_B200_DUMP_DIR"] = old


def _matrix_dump_files(gpu, tmp_path):
    """`!SOLVER, DUMPTYPE=MM|CSR|BSR` (hecmw_mat_dump / _rhs / _solution, matrix/hecmw_matrix_dump.f90): the files the solve
    writes - names, headers, entry order (L, D, U per row; all NP rows), Fortran E20.12E3 numbers - read back to the matrix,
    right-hand side and solution that went through the solve; DUMPEXIT returns without solving."""
    from frontistr_b200 import mshio
    from frontistr_b200._lib import FistrB200Error
    mesh = util.distort(meshgen.hex_cube(3))
    A = util.product_problem(gpu, mesh, util.cube_cnt(tol=1e-10), to_host=True)
    import glob
    got = {}
    for dtype, ext in ((1, "mm"), (2, "csr"), (3, "bsr")):
        M = A.copy(); M.Iarray[30] = dtype; M.X[:] = 0.0; M.Iarray[96] = 1
        assert gpu.hecmw_solve(mesh, M)["error"] == 0
        files = sorted(glob.glob(str(tmp_path / ("dump_matrix_*_0.%s" % ext))))
        assert len(files) == 1
        stem = files[0][: -len(ext) - 1]
        got[ext] = mshio.read_mat_dump(files[0])
        rhs, sol = mshio.read_mat_dump(stem + ".rhs"), mshio.read_mat_dump(stem + ".sol")
        assert len(rhs) == 3 * A.NP and np.allclose(rhs, A.B, rtol=1e-11, atol=0) and np.allclose(sol, M.X, rtol=1e-11, atol=1e-300)
        with open(files[0]) as f:
            first = f.readline().rstrip("\n")
        assert first == {"mm": "%%MatrixMarket matrix coordinate real general", "csr": "%%CSR matrix real general",
                         "bsr": "%%Block-CSR matrix real general"}[ext]
        with open(stem + ".rhs") as f:
            ln = f.readline().rstrip("\n")
        assert len(ln) == 20 and ln[1:3] == "0." and ln[15] == "E" and ln[16] in "+-"      # E20.12E3
    # dense reconstruction from each format equals the hecmwST_matrix
    n3 = 3 * A.NP
    K = np.zeros((n3, n3))
    for i in range(A.NP):
        K[3 * i: 3 * i + 3, 3 * i: 3 * i + 3] = A.D[9 * i: 9 * i + 9].reshape(3, 3)
        for k in range(A.indexL[i], A.indexL[i + 1]):
            j = A.itemL[k] - 1; K[3 * i: 3 * i + 3, 3 * j: 3 * j + 3] = A.AL[9 * k: 9 * k + 9].reshape(3, 3)
        for k in range(A.indexU[i], A.indexU[i + 1]):
            j = A.itemU[k] - 1; K[3 * i: 3 * i + 3, 3 * j: 3 * j + 3] = A.AU[9 * k: 9 * k + 9].reshape(3, 3)
    tol = 1e-11 * np.abs(K).max()
    mm = got["mm"]; Kmm = np.zeros_like(K); Kmm[mm["row"] - 1, mm["col"] - 1] = mm["value"]
    assert mm["nrow"] == n3 and np.abs(Kmm - K).max() <= tol
    csr = got["csr"]; Kc = np.zeros_like(K)
    for r in range(n3):
        Kc[r, csr["item"][csr["index"][r]: csr["index"][r + 1]] - 1] = csr["value"][csr["index"][r]: csr["index"][r + 1]]
    assert np.abs(Kc - K).max() <= tol and np.all(np.diff(csr["item"][csr["index"][0]: csr["index"][1]]) > 0)
    bsr = got["bsr"]; Kb = np.zeros_like(K)
    for i in range(A.NP):
        for k in range(bsr["index"][i], bsr["index"][i + 1]):
            j = bsr["item"][k] - 1; Kb[3 * i: 3 * i + 3, 3 * j: 3 * j + 3] = bsr["value"][k]
    assert bsr["ndof"] == 3 and bsr["nrow"] == A.NP and np.abs(Kb - K).max() <= tol
    # DUMPEXIT: the dump is written, nothing is solved
    E = A.copy(); E.Iarray[30] = 3; E.Iarray[31] = 1; E.X[:] = 7.0
    with pytest.raises(FistrB200Error):
        gpu.hecmw_solve(mesh, E)
    assert np.all(E.X == 7.0)


def test_assembly_error_contract(gpu):
    """A malformed mesh is an error code, never a device fault: node ids outside 1..n_node, an element with the wrong
    number of nodes, a section id out of range, an element option outside 1..3, an element type outside the path; and an
    element whose node pair is missing from the matrix graph is the reference's "cannot find connectivity"
    (hecmw_mat_ass.f90:86-90) = FISTR_B200_ERR_CONNECTIVITY."""
    from frontistr_b200._lib import FistrB200Error

    def fresh():
        m = meshgen.hex_cube(3)
        A_ = gpu.hecmw_mat_con(m); A_.fstr_mat_init().hecMAT_init()
        gpu.mat_structure(m, A_)
        return m, A_
    cases = []
    m, A = fresh(); m.elem_node_item = m.elem_node_item.copy(); m.elem_node_item[5] = m.n_node + 1; cases.append((m, A, -1))
    m, A = fresh(); m.elem_node_item = m.elem_node_item.copy(); m.elem_node_item[9] = 0; cases.append((m, A, -1))
    m, A = fresh(); m.elem_node_index = m.elem_node_index.copy(); m.elem_node_index[2] += 1; cases.append((m, A, -1))
    m, A = fresh(); m.section_ID = m.section_ID.copy(); m.section_ID[0] = 2; cases.append((m, A, -1))
    m, A = fresh(); m.sections = [dict(m.sections[0], elemopt361=4)]; cases.append((m, A, -1))
    m, A = fresh(); m.elem_type_item = np.array([351], np.int32); cases.append((m, A, -1))
    # valid ids, but the pair (1, n_node) is not an edge of the matrix graph the structure was built from
    m, A = fresh(); m.elem_node_item = m.elem_node_item.copy(); m.elem_node_item[0] = m.n_node; cases.append((m, A, -5))
    for m, A, code in cases:
        with pytest.raises(FistrB200Error) as ei:
            gpu.fstr_StiffMatrix(m, A, to_host=False)
        assert ei.value.code == code, (ei.value.code, code, str(ei.value))
    # and the library is fine afterwards
    mesh = util.distort(meshgen.hex_cube(3))
    A = util.product_problem(gpu, mesh, util.cube_cnt(), to_host=True)
    assert gpu.hecmw_solve(mesh, A)["error"] == 0


def test_structure_size_limits(gpu):
    """Maximum sizes: blocks are addressed by 32-bit slots ([D | AL | AU], csrc/assembly.cuh), so a structure with
    NP + NPL + NPU >= 2^32 - 2 is refused before anything is allocated; so are negative / inconsistent sizes."""
    import ctypes as C
    from frontistr_b200 import _lib
    L = _lib.lib()
    one = np.zeros(2, np.int32)
    ip_ = one.ctypes.data_as(C.POINTER(C.c_int))
    big = 2 ** 31 - 1
    for N, NP, NPL, NPU in ((1, 1, big, big), (2, 1, 0, 0), (-1, 1, 0, 0), (1, 1, -5, 0)):
        rc = L.fistr_b200_mat_structure(N, NP, NPL, NPU, ip_, ip_, ip_, ip_, 0, None, None, None, None, None)
        assert rc == -1, (N, NP, NPL, NPU, rc)
    assert b"32-bit slots" in L.fistr_b200_last_error() or b"bad sizes" in L.fistr_b200_last_error()
    # a malformed graph or halo table is refused before it reaches a kernel: column out of range, L entry above the
    # diagonal, unsorted row, index array that does not end at NPL, import entry that names an internal node
    from frontistr_b200._lib import FistrB200Error
    m0 = meshgen.hex_cube(2)
    A0 = gpu.hecmw_mat_con(m0)

    def declare(A_, m_):
        A_.fstr_mat_init().hecMAT_init()
        gpu.mat_structure(m_, A_)
    bad = []
    Ab = A0.copy(); Ab.itemL = A0.itemL.copy(); Ab.itemL[3] = A0.NP + 7; bad.append((Ab, m0))
    Ab = A0.copy(); Ab.itemL = A0.itemL.copy(); Ab.itemL[0] = A0.NP; bad.append((Ab, m0))
    Ab = A0.copy(); Ab.itemU = A0.itemU.copy(); Ab.itemU[[0, 1]] = Ab.itemU[[1, 0]]; bad.append((Ab, m0))
    Ab = A0.copy(); Ab.indexU = A0.indexU.copy(); Ab.indexU[-1] -= 1; bad.append((Ab, m0))
    from frontistr_b200 import partition
    mh = partition.local_mesh(m0, partition.rcb_domains(m0, "x"), 0, 2)
    mb = partition.local_mesh(m0, partition.rcb_domains(m0, "x"), 0, 2); mb.import_item = mb.import_item.copy(); mb.import_item[0] = 1
    bad.append((gpu.hecmw_mat_con(mb), mb))
    for Ab, mb_ in bad:
        with pytest.raises(FistrB200Error) as ei:
            declare(Ab, mb_)
        assert ei.value.code == -1 and "mat_structure" in str(ei.value)
    declare(gpu.hecmw_mat_con(mh), mh)            # the untouched local mesh of rank 0 is accepted
    # the library is still usable afterwards
    mesh = meshgen.hex_cube(2)
    A = util.product_problem(gpu, mesh, util.cube_cnt(), to_host=True)
    assert gpu.hecmw_solve(mesh, A)["error"] == 0


def test_precond_recycle_flags(gpu):
    """Iarray(96/97/98) policy (matrix_misc.f90:678-697): numeric set-up reused up to 3 solves."""
    mesh = meshgen.hex_cube(4)
    A = util.product_problem(gpu, mesh, util.cube_cnt(), to_host=True)
    assert A.Iarray[97] == 1
    gpu.hecmw_solve(mesh, A)
    assert A.Iarray[97] == 0 and A.Iarray[96] == 0 and A.Iarray[95] == 0
    A.Iarray[96] = 1                       # FrontISTR sets numfact=1 on later Newton iterations
    gpu.hecmw_solve(mesh, A)
    assert A.Iarray[95] == 1 and A.Iarray[96] == 0


def test_large_properties(gpu):
    """Size-independent properties at a size the oracle cannot cross-check quickly (64^3, 0.8 M DOF):
    symmetry x.Ay == y.Ax, solve reaches the tolerance on the TRUE residual, and SpMV variants agree bitwise."""
    mesh = meshgen.hex_cube(64)
    cnt = util.cube_cnt(tol=1e-8)
    gpu.set_option("ic_inverse", 0)        # 2.6e5 IC elements would start the library's kernel selection (a test of its own,
    try:                                   # tests/test_gpu_z_unvalidated.py): this test keeps to the kernel that has run on hardware
        A = util.product_problem(gpu, mesh, cnt, to_host=False)
    finally:
        gpu.set_option("ic_inverse", -1)
    n3 = 3 * A.N
    x = np.sin(np.arange(1, n3 + 1, dtype=float)); y = np.cos(0.37 * np.arange(1, n3 + 1, dtype=float))
    ys = {}
    for v in (0, 1, 2):
        gpu.set_option("spmv_variant", v)
        ys[v] = gpu.hecmw_matvec(x, n3)
    gpu.set_option("spmv_variant", 2)
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[1], ys[2])
    Ay = gpu.hecmw_matvec(y, n3)
    assert abs(x @ Ay - y @ ys[2]) <= 1e-10 * abs(x @ Ay)
    r = gpu.hecmw_solve(mesh, A, device_resident=True)
    assert r["error"] == 0 and r["resid2"] < 1e-8 and A.Iarray[80] == 1
    assert np.all(A.X.reshape(-1, 3)[mesh.node_group["FIX"] - 1] == 0.0)
