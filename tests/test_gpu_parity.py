"""Parity of the CUDA path (through the C ABI) against the CPU oracle. Needs a B200."""
import ctypes as C

import numpy as np
import pytest

from helpers import RTOL, UTOL, bar_problem, distorted, grid_problem, node_dofs, rel_err

pytestmark = pytest.mark.gpu


# ---- T1: the reference's 1-D path -------------------------------------------------------
@pytest.mark.parametrize("p,q", [(1, 1), (1, 2), (2, 3), (3, 4), (3, 5), (4, 5), (7, 5)])
@pytest.mark.parametrize("matname", ["linear", "exponential"])
def test_bar_residual_jacobian(nls, orc, p, q, matname):
    mat = nls.LinearElasticity1D(E=2.5) if matname == "linear" else nls.ExponentialMaterial(1.0, 2.0)
    fem, om = bar_problem(nls, orc, 37, p, q, mat, area=0.7)
    rng = np.random.default_rng(p * 10 + q)
    U = rng.uniform(-0.05, 0.05, fem.ndof)
    K = nls.CSRMatrix(fem)
    rp, ci = om.pattern()
    assert np.array_equal(rp, K.rowptr) and np.array_equal(ci, K.colind)
    R = np.zeros(fem.ndof)
    Ud = U.copy()
    nls.Residual(mat)(fem, Ud, R)
    nls.Jacobian(mat)(fem, Ud, K)
    Uo, Ro = om.residual(U)
    assert np.array_equal(Ud, Uo)                      # Dirichlet overwrite of U
    assert rel_err(R, Ro) <= RTOL
    Ko = om.jacobian_csr(U)
    assert rel_err(K.vals, Ko) <= RTOL
    Kd = om.jacobian_dense(U)                          # the reference's dense K has nothing outside the pattern
    assert rel_err(K.toarray(), Kd) <= RTOL
    # fused pass == separate passes
    R2, K2 = np.zeros(fem.ndof), nls.CSRMatrix(fem)
    nls.residual_jacobian(mat, fem, U.copy(), R2, K2)
    assert rel_err(R2, R) <= 1e-14 and rel_err(K2.vals, K.vals) <= 1e-14


@pytest.mark.parametrize("nel,p,q", [(4, 1, 2), (3, 2, 3), (100, 3, 4)])
def test_kat3_exponential_bar_newton(nls, orc, nel, p, q):
    """KAT-3: tip displacement ln2/2, 5 iterations, CONVERGED_RELATIVE."""
    mat = nls.ExponentialMaterial(1.0, 2.0)
    fem, om = bar_problem(nls, orc, nel, p, q, mat)
    s = nls.NewtonSolver(fem, rtol=1e-10)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve()
    ref = om.newton(rtol=1e-10)                        # faithful: dense K + LU
    assert s.converged_reason is nls.REASON_CONVERGED_RELATIVE and ref["reason"] == "CONVERGED_RELATIVE"
    assert s.iterations == ref["iterations"] == 5
    assert abs(s.U[-1] - 0.34657359027997264) <= 1e-10
    assert rel_err(s.U, ref["U"]) <= UTOL
    np.testing.assert_allclose(s.norm_history[:5], ref["hist"][:5], rtol=1e-6, atol=1e-12)


def test_newton_semantics(nls, orc):
    """maxits quirk, dtol, :modified, absolute tolerance -- newton_fem.jl:145-187."""
    mat = nls.ExponentialMaterial(1.0, 2.0)
    fem, om = bar_problem(nls, orc, 8, 2, 3, mat)
    mk = lambda **kw: _solver(nls, fem, mat, **kw)
    s = mk(rtol=1e-10, maxits=5); s.solve()
    assert s.converged_reason is nls.REASON_DIVERGED_MAXITS and s.iterations == 5   # converged on the last allowed it
    assert om.newton(rtol=1e-10, maxits=5)["reason"] == "DIVERGED_MAXITS"
    s = mk(rtol=1e-10, dtol=1e-3); s.solve()
    assert s.converged_reason is nls.REASON_DIVERGED_STEP and s.iterations == 0
    s = mk(rtol=1e-30, atol=1e-9, maxits=50); s.solve()
    ref = om.newton(rtol=1e-30, atol=1e-9, maxits=50)
    assert s.converged_reason is nls.REASON_CONVERGED_ABSOLUTE and s.iterations == ref["iterations"]
    s = mk(rtol=1e-8, maxits=200, type="modified"); s.solve()
    ref = om.newton(rtol=1e-8, maxits=200, modified=True)
    assert repr(s.converged_reason).startswith("Converged") and s.iterations == ref["iterations"] > 6
    assert rel_err(s.U, ref["U"]) <= 1e-8
    assert not nls.is_converged(mk(maxits=0)) 


def _solver(nls, fem, mat, **kw):
    s = nls.NewtonSolver(fem, **kw)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    return s


@pytest.mark.parametrize("p,q", [(1, 2), (2, 3), (3, 4)])
def test_plastic_bar_load_stepping(nls, orc, p, q):
    """KAT-5 + the PseudoTimeStepper protocol: trial state in the residual pass, tangent read by the
    Jacobian pass, commit on poststep! (linearelastoplasticity.jl:60-103, pseudotime.jl:73-93).
    Force-driven loading 1.6 sin(2.5 t): elastic, yield with tangent Eep, then elastic unloading."""
    mat = nls.LinearElastoPlasticity1D(E=100.0, H_kappa=10.0, H_alpha=5.0, kappa0=1.0, alpha0=0.0)
    assert abs(mat.Eep - 100.0 * 15.0 / 115.0) < 1e-14
    mesh = nls.bar_mesh(6, p)
    fem = nls.FEMModel(mesh, q, p, data={"A": 1.0})
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    fem.addboundary(nls.TimeDependent(nls.Neumann([mesh.nn - 1], [0.0]), lambda t: [1.6 * np.sin(2.5 * t)]))
    nls.initialize_state(mat, fem)
    s = _solver(nls, fem, mat, rtol=1e-10, atol=1e-12, maxits=30)
    ts = nls.PseudoTimeStepper(s, dt=0.125, maxtime=1.0)
    ts.solve()
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_PLASTIC, mat=mat.params(), area=1.0)
    om.set_dirichlet([0], [0.0])
    U = np.zeros(fem.ndof)
    its = []
    for k in range(1, 9):
        om.set_neumann([mesh.nn - 1], [1.6 * np.sin(2.5 * 0.125 * k)])
        r = om.newton(U0=U, rtol=1e-10, atol=1e-12, maxits=30)
        assert r["reason"].startswith("CONVERGED")
        U = r["U"]
        om.commit_state(U)
        its.append(r["iterations"])
        assert rel_err(ts.U[k], U) <= UTOL, k
        assert ts.num_its[k, 0] == r["iterations"], k
    assert ts.step == 8 and its == [1, 1, 2, 2, 2, 1, 1, 1]
    for name in ("sig", "alp", "kap", "gam", "dsig", "sig_n", "alp_n", "kap_n", "gam_n"):
        got, want = nls.getstate(fem, name), om.state(name)
        assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max()), name
    assert om.state("gam").max() > 0.03      # the path did yield


def test_timestepper_aborts_on_nonconvergence(nls, orc):
    """A displacement-driven p=2 plastic bar makes the reference's Newton 2-cycle (the oracle shows
    the same); the timestepper must stop at the first non-converged step and trim (pseudotime.jl:83-86)."""
    mat = nls.LinearElastoPlasticity1D(E=100.0, H_kappa=10.0, H_alpha=5.0, kappa0=1.0, alpha0=0.0)
    mesh = nls.bar_mesh(6, 2)
    fem = nls.FEMModel(mesh, 3, 2, data={"A": 1.0})
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    fem.addboundary(nls.TimeDependent(nls.Dirichlet([mesh.nn - 1], [0.0]), lambda t: [0.03 * np.sin(2.5 * t)]))
    nls.initialize_state(mat, fem)
    s = _solver(nls, fem, mat, rtol=1e-10, atol=1e-12, maxits=12)
    ts = nls.PseudoTimeStepper(s, dt=0.125, maxtime=1.0)
    ts.solve()
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=2, nq=3, mat_kind=orc.MAT_PLASTIC, mat=mat.params(), area=1.0)
    om.set_dirichlet([0, mesh.nn - 1], [0.0, 0.03 * np.sin(2.5 * 0.125)])
    ref = om.newton(rtol=1e-10, atol=1e-12, maxits=12)
    assert ref["reason"] == "DIVERGED_MAXITS"
    assert s.converged_reason is nls.REASON_DIVERGED_MAXITS and ts.step == 1
    assert ts.residual.shape[0] == 2                     # trimmed to the steps taken
    np.testing.assert_allclose(s.norm_history, ref["hist"], rtol=1e-9)


def test_plastic_below_yield_equals_linear(nls, orc):
    """KAT-5: below yield the plastic bar is LinearElasticity1D."""
    E = 50.0
    mp = nls.LinearElastoPlasticity1D(E=E, H_kappa=1.0, H_alpha=1.0, kappa0=10.0, alpha0=0.0)
    ml = nls.LinearElasticity1D(E=E)
    outs = []
    for mat in (mp, ml):
        fem, _ = bar_problem(nls, orc, 5, 1, 2, mat, tip_force=0.3)
        if mat.has_state:
            nls.initialize_state(mat, fem)
        s = _solver(nls, fem, mat, rtol=1e-12, atol=1e-13, maxits=10); s.solve()
        outs.append(s.U.copy())
    assert rel_err(outs[0], outs[1]) <= 1e-12


# ---- T2: Q1 Neo-Hookean -----------------------------------------------------------------
@pytest.mark.parametrize("dim,n", [(2, 9), (2, 24), (3, 5), (3, 9)])
@pytest.mark.parametrize("mesh_kind", ["affine", "distorted"])
def test_q1_residual_jacobian(nls, orc, dim, n, mesh_kind):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, om = grid_problem(nls, orc, n, dim, mat, compress=-0.2 / (n - 1))   # < element size: no inversion
    if mesh_kind == "distorted":
        X = distorted(fem.mesh)
        fem.mesh.coords[:] = X
        om.coords[:] = X
    U = nls.smooth_field(fem.mesh.coords, amp=0.02)
    K = nls.CSRMatrix(fem)
    rp, ci = om.pattern()
    assert np.array_equal(rp, K.rowptr) and np.array_equal(ci, K.colind)      # bit-exact
    R, Ud = np.zeros(fem.ndof), U.copy()
    nls.residual_jacobian(mat, fem, Ud, R, K)
    Uo, Ro = om.residual(U)
    Ko = om.jacobian_csr(U)
    assert np.array_equal(Ud, Uo)
    assert rel_err(R, Ro) <= RTOL, rel_err(R, Ro)
    assert rel_err(K.vals, Ko) <= RTOL, rel_err(K.vals, Ko)
    # separate passes agree with the fused one
    R1, K1 = np.zeros(fem.ndof), nls.CSRMatrix(fem)
    nls.Residual(mat)(fem, U.copy(), R1); nls.Jacobian(mat)(fem, U.copy(), K1)
    assert rel_err(R1, R) <= 1e-13 and rel_err(K1.vals, K.vals) <= 1e-13


@pytest.mark.parametrize("dim,n,compress", [(2, 17, -0.03), (3, 7, -0.1)])
def test_q1_newton(nls, orc, dim, n, compress):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, om = grid_problem(nls, orc, n, dim, mat, compress=compress)
    s = _solver(nls, fem, mat, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-12)
    s.solve()
    ref = om.newton(rtol=1e-10, atol=1e-14, maxits=25, linear="pcg", lin_rtol=1e-12)
    assert repr(s.converged_reason).startswith("Converged") and ref["reason"].startswith("CONVERGED")
    assert s.iterations == ref["iterations"]
    assert rel_err(s.U, ref["U"]) <= UTOL
    if fem.ndof <= 1500:      # also against the faithful dense-LU variant
        ref2 = om.newton(rtol=1e-10, atol=1e-14, maxits=25, linear="dense")
        assert s.iterations == ref2["iterations"] and rel_err(s.U, ref2["U"]) <= UTOL


def test_krylov_hooks(nls, orc):
    """SpMV, dot, axpy and the PCG solve against the oracle on the same CSR values."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, om = grid_problem(nls, orc, 8, 3, mat, compress=-0.01)     # mild state: the tangent stays SPD
    U = nls.smooth_field(fem.mesh.coords, amp=0.002)
    vals = om.jacobian_csr(U)
    H = fem.handle(mat)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    H.call("nls_set_values", dp(vals))
    rng = np.random.default_rng(3)
    x = rng.standard_normal(fem.ndof)
    y = np.zeros(fem.ndof)
    H.call("nls_spmv", dp(x), dp(y))
    assert rel_err(y, om.spmv(vals, x)) <= 1e-13
    z = rng.standard_normal(fem.ndof)
    out = C.c_double(0)
    H.call("nls_dot", fem.ndof, dp(x), dp(z), C.byref(out))
    assert abs(out.value - float(x @ z)) <= 1e-12 * np.abs(x * z).sum()
    z2 = z.copy()
    H.call("nls_axpy", fem.ndof, 0.37, dp(x), dp(z2))
    assert np.abs(z2 - (z + 0.37 * x)).max() <= 1e-15
    b = rng.standard_normal(fem.ndof)
    xs = np.zeros(fem.ndof)
    its, st, rr = C.c_int32(0), C.c_int32(0), C.c_double(0)
    H.call("nls_linear_solve", 1e-12, 5000, dp(b), dp(xs), C.byref(its), C.byref(rr), C.byref(st))
    rc, xo, ito, rro = om.pcg(vals, b, rtol=1e-12, maxits=5000)
    assert st.value == 0 and rc == 0
    assert abs(its.value - ito) <= 2
    assert rel_err(xs, xo) <= 1e-9
    con = om.constrained_mask().astype(bool)
    assert np.all(xs[con] == 0.0)
    # maxits exhaustion is reported, not hidden
    H.call("nls_linear_solve", 1e-12, 3, dp(b), dp(xs), C.byref(its), C.byref(rr), C.byref(st))
    assert st.value == 1 and its.value == 3


def test_symmetric_product_option(nls, orc):
    """The experimental upper-blocks-only SpMV of the Krylov solvers (option sym_spmv, off by default) gives the same
    PCG and MINRES solutions as the full product."""
    for dim, n in ((2, 12), (3, 7)):
        mat = nls.NeoHookean(E=1.0, nu=0.3)
        fem, om = grid_problem(nls, orc, n, dim, mat, compress=-0.01)
        vals = om.jacobian_csr(nls.smooth_field(fem.mesh.coords, amp=0.002))
        H = fem.handle(mat)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        H.call("nls_set_option", b"pcg_small", 0)
        b = np.random.default_rng(11).standard_normal(fem.ndof)
        sols = {}
        for sym in (0, 1):
            H.call("nls_set_option", b"sym_spmv", sym)
            H.call("nls_set_values", dp(vals))
            for name in ("nls_linear_solve", "nls_linear_solve_indefinite"):
                x = np.zeros(fem.ndof)
                its, st, rr = C.c_int32(0), C.c_int32(0), C.c_double(0)
                H.call(name, 1e-12, 20000, dp(b), dp(x), C.byref(its), C.byref(rr), C.byref(st))
                assert st.value == 0
                sols[(sym, name)] = (x, its.value)
        for name in ("nls_linear_solve", "nls_linear_solve_indefinite"):
            assert rel_err(sols[(1, name)][0], sols[(0, name)][0]) <= 1e-9
            assert abs(sols[(1, name)][1] - sols[(0, name)][1]) <= 3


def test_errors_are_loud(nls):
    H = nls.Handle(0)
    with pytest.raises(nls.NLSError):
        H.call("nls_build_pattern")                      # no mesh yet
    mesh = nls.grid_mesh(4, 2)
    fem = nls.FEMModel(mesh)
    with pytest.raises(nls.NLSError):
        fem.handle(nls.LinearElasticity1D(1.0)).call("nls_dev_assemble", 1, 1)   # 1-D law on a 2-D mesh
    with pytest.raises(ValueError):
        nls.gauss_quadrature(6)
    bad = nls.Mesh(2, np.array([[0, 1, 2, 99]], dtype=np.int32), coords=np.zeros((4, 2)))
    with pytest.raises(nls.NLSError):
        nls.FEMModel(bad).handle()


def test_empty_and_duplicate_boundaries(nls, orc):
    """No BCs at all (singular K is the solver's problem, assembly must still work), and repeated
    Dirichlet DoFs where the last value wins like sequential apply!."""
    mat = nls.ExponentialMaterial(1.0, 2.0)
    mesh = nls.bar_mesh(5, 1)
    fem = nls.FEMModel(mesh, 2, 1, data={"A": 1.0})
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=1, nq=2, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=1.0)
    U = np.linspace(0, 0.1, fem.ndof)
    R = np.zeros(fem.ndof); nls.Residual(mat)(fem, U.copy(), R)
    assert rel_err(R, om.residual(U)[1]) <= RTOL
    fem.addboundary(nls.Dirichlet([0, 5, 0], [1.0, 2.0, 3.0]))
    fem.addboundary(nls.Neumann([2, 2], [0.25, 0.5]))
    om.set_dirichlet([0, 5, 0], [1.0, 2.0, 3.0]); om.set_neumann([2, 2], [0.25, 0.5])
    Ud = U.copy(); nls.Residual(mat)(fem, Ud, R)
    Uo, Ro = om.residual(U)
    assert Ud[0] == 3.0 and np.array_equal(Ud, Uo) and rel_err(R, Ro) <= RTOL


@pytest.mark.parametrize("n,kind", [(24, "affine"), (37, "distorted"), (66, "shuffled")])
def test_gather_vs_atomic_assembly_2d(nls, orc, n, kind):
    """The two 2-D assembly kernels (gather/row-owner = default, atomic scatter = option 0) agree with
    each other and with the oracle, also on an unstructured numbering; the gather kernel is
    bit-reproducible run to run."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mesh = nls.grid_mesh(n, 2)
    if kind == "distorted":
        mesh.coords[:] = distorted(mesh)
    if kind == "shuffled":
        rng = np.random.default_rng(5)
        perm = rng.permutation(mesh.nn)
        conn = perm[mesh.conn][rng.permutation(mesh.nel)].astype(np.int32)
        coords = np.zeros_like(mesh.coords); coords[perm] = mesh.coords
        mesh = nls.Mesh(2, conn, coords=coords)
    fem = nls.FEMModel(mesh)
    fem.addboundary(nls.Dirichlet([0, 1, 7], [0.0, 0.001, -0.002]))
    fem.addboundary(nls.Neumann([5, 9], [0.25, -0.5]))
    om = orc.OracleModel(2, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet([0, 1, 7], [0.0, 0.001, -0.002]); om.set_neumann([5, 9], [0.25, -0.5])
    U = nls.smooth_field(mesh.coords, amp=0.2 / (n - 1))    # well below the element size: no inversion at the BC nodes
    H = fem.handle(mat)
    out = {}
    for opt in (1, 0, 1):
        H.call("nls_set_option", b"assembly", opt)
        R, K = np.full(fem.ndof, np.nan), nls.CSRMatrix(fem)
        K.vals[:] = np.nan
        nls.residual_jacobian(mat, fem, U.copy(), R, K)
        if opt == 1 and 1 in out:
            assert np.array_equal(out[1][0], R) and np.array_equal(out[1][1], K.vals)   # deterministic
        out[opt] = (R, K.vals.copy())
        R1, K1 = np.zeros(fem.ndof), nls.CSRMatrix(fem)
        nls.Residual(mat)(fem, U.copy(), R1); nls.Jacobian(mat)(fem, U.copy(), K1)
        assert rel_err(R1, R) <= 1e-13 and rel_err(K1.vals, K.vals) <= 1e-13
    _, Ro = om.residual(U); Ko = om.jacobian_csr(U)
    for opt in (0, 1):
        assert rel_err(out[opt][0], Ro) <= RTOL and rel_err(out[opt][1], Ko) <= RTOL, opt


@pytest.mark.parametrize("dim,n,kind,scratch_mb", [(2, 37, "distorted", 0), (2, 66, "shuffled", 1), (3, 9, "distorted", 0),
                                                     (3, 13, "affine", 1), (3, 12, "shuffled", 1)])
def test_assembly_schemes_agree(nls, orc, dim, n, kind, scratch_mb):
    """Every assembly scheme (0 = atomic scatter, 1 = default row-owner scheme of the dimension, 2 = element
    records + row owners) against the oracle and against each other, on structured, distorted and randomly
    renumbered meshes; scratch_mb = 1 forces the record scratch into several node chunks. The row-owner
    schemes sum in element order, so they are bit-reproducible and residual-only / Jacobian-only passes
    reproduce the fused pass exactly."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mesh = nls.grid_mesh(n, dim)
    if kind == "distorted":
        mesh.coords[:] = distorted(mesh)
    if kind == "shuffled":
        rng = np.random.default_rng(5)
        perm = rng.permutation(mesh.nn)
        conn = perm[mesh.conn][rng.permutation(mesh.nel)].astype(np.int32)
        coords = np.zeros_like(mesh.coords); coords[perm] = mesh.coords
        mesh = nls.Mesh(dim, conn, coords=coords)
    fem = nls.FEMModel(mesh)
    fem.addboundary(nls.Dirichlet([0, 1, 7], [0.0, 0.001, -0.002]))
    fem.addboundary(nls.Neumann([5, 9], [0.25, -0.5]))
    om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet([0, 1, 7], [0.0, 0.001, -0.002]); om.set_neumann([5, 9], [0.25, -0.5])
    U = nls.smooth_field(mesh.coords, amp=0.2 / (n - 1))
    H = fem.handle(mat)
    if scratch_mb:
        H.call("nls_set_option", b"scratch_mb", scratch_mb)
    _, Ro = om.residual(U); Ko = om.jacobian_csr(U)
    out = {}
    for opt in (2, 0, 1, 2):
        H.call("nls_set_option", b"assembly", opt)
        R, K = np.full(fem.ndof, np.nan), nls.CSRMatrix(fem)
        K.vals[:] = np.nan
        nls.residual_jacobian(mat, fem, U.copy(), R, K)
        assert rel_err(R, Ro) <= RTOL and rel_err(K.vals, Ko) <= RTOL, opt
        if opt == 2 and 2 in out:
            assert np.array_equal(out[2][0], R) and np.array_equal(out[2][1], K.vals)   # deterministic
        out[opt] = (R, K.vals.copy())
        R1, K1 = np.full(fem.ndof, np.nan), nls.CSRMatrix(fem)
        K1.vals[:] = np.nan
        nls.Residual(mat)(fem, U.copy(), R1); nls.Jacobian(mat)(fem, U.copy(), K1)
        assert rel_err(R1, R) <= 1e-13 and rel_err(K1.vals, K.vals) <= 1e-13
    assert rel_err(out[2][1], out[0][1]) <= 1e-13 and rel_err(out[2][1], out[1][1]) <= 1e-13


def test_distributed_force(nls, orc):
    """DistributedForce on the device (distributedforce.jl:7-27, residual.jl:19-25): 1-D with a function of
    time evaluated through the timestepper context, and constant 2-D / 3-D body loads."""
    E, A = 3.0, 0.5
    mat = nls.LinearElasticity1D(E=E)
    mesh = nls.bar_mesh(9, 3)
    fem = nls.FEMModel(mesh, 4, 3, data={"A": A})
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    s = _solver(nls, fem, mat, rtol=1e-11, atol=1e-14, maxits=5)
    nls.set_residual_fn(s, nls.Residual(mat, [nls.DistributedForce(lambda t: 2.0 * t)]))
    ts = nls.PseudoTimeStepper(s, dt=0.25, maxtime=0.5)
    nls.set_ctx(s, ts)
    ts.solve()
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=3, nq=4, mat_kind=orc.MAT_LINEAR, mat=(E,), area=A)
    om.set_dirichlet([0], [0.0])
    U = np.zeros(fem.ndof)
    for k, t in [(1, 0.25), (2, 0.5)]:
        om.set_body_force([2.0 * t])
        r = om.newton(U0=U, rtol=1e-11, atol=1e-14, maxits=5)
        U = r["U"]
        assert rel_err(ts.U[k], U) <= UTOL and ts.num_its[k, 0] == r["iterations"]
    for dim, n, body in [(2, 11, [0.3, -0.7]), (3, 6, [0.1, 0.2, -0.5])]:
        mat2 = nls.NeoHookean(E=1.0, nu=0.3)
        fem2, om2 = grid_problem(nls, orc, n, dim, mat2, compress=-0.1 / (n - 1))
        X = distorted(fem2.mesh); fem2.mesh.coords[:] = X; om2.coords[:] = X
        om2.set_body_force(body)
        U = nls.smooth_field(X, amp=0.1 / (n - 1))
        R = np.zeros(fem2.ndof)
        nls.Residual(mat2, [nls.DistributedForce(body)])(fem2, U.copy(), R)
        assert rel_err(R, om2.residual(U)[1]) <= RTOL
        R0 = np.zeros(fem2.ndof)
        nls.Residual(mat2)(fem2, U.copy(), R0)                    # removing the force restores the plain residual
        om2.set_body_force(None)
        assert rel_err(R0, om2.residual(U)[1]) <= RTOL and np.abs(R0 - R).max() > 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_compute_internalforce_is_pure(nls, orc, dim):
    """The material-level dispatch point (defaults.jl:29-46, 68-82): F_int and dF_int/dU alone -- U untouched, Dirichlet
    values not applied, Neumann / body loads not subtracted, mode add/set, invalid mode raises."""
    if dim == 1:
        mat = nls.ExponentialMaterial(sigma_sat=1.0, b=2.0)
        mesh = nls.bar_mesh(7, 2)
        fem = nls.FEMModel(mesh, 3, 2, data={"A": 1.0})
        om = orc.OracleModel(1, mesh.conn, mesh.coords, p=2, nq=3, mat_kind=mat.kind, mat=mat.params(), area=1.0)
        U = 0.01 * np.arange(mesh.nn, dtype=np.float64) ** 1.5
    else:
        mat = nls.NeoHookean(E=1.0, nu=0.3)
        mesh = nls.grid_mesh(5, dim)
        mesh.coords[:] = distorted(mesh)
        fem = nls.FEMModel(mesh)
        om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
        U = nls.smooth_field(mesh.coords, amp=0.03)
    fem.addboundary(nls.Dirichlet([0, 1], [0.5, -0.25]))          # must NOT be applied by the pure loops
    fem.addboundary(nls.Neumann([fem.ndof - 1], [3.0]))           # must NOT be subtracted
    _, Ro = om.residual(U); Ko = om.jacobian_csr(U)                # the oracle model carries no boundary: pure F_int
    U0 = U.copy()
    F = np.full(fem.ndof, 2.0)
    nls.compute_internalforce(mat, fem, U, F, mode="add")
    assert np.array_equal(U, U0)
    assert rel_err(F - 2.0, Ro) <= RTOL
    nls.compute_internalforce(mat, fem, U, F, mode="set")
    assert rel_err(F, Ro) <= RTOL
    K = nls.CSRMatrix(fem)
    K.vals[:] = 1.0
    nls.compute_dinternalforce(mat, fem, U, K, mode="add")
    assert np.array_equal(U, U0) and rel_err(K.vals - 1.0, Ko) <= RTOL
    nls.compute_dinternalforce(mat, fem, U, K, mode="set")
    assert rel_err(K.vals, Ko) <= RTOL
    with pytest.raises(ValueError):
        nls.compute_internalforce(mat, fem, U, F, mode="replace")
    # the functors on top still apply the boundaries
    R = np.zeros(fem.ndof)
    U1 = U.copy()
    nls.Residual(mat)(fem, U1, R)
    assert U1[0] == 0.5 and U1[1] == -0.25
