"""CPU tests (no GPU): pin the oracle against everything the reference's own tests assert for this
path (test/runtests.jl:54-207), the known answers KAT-1..5 (tests/golden/reference_kats.json),
and -- for the builder-defined Q1 Neo-Hookean extension -- an independent torch-autograd
differentiation of the strain energy."""
import json
import os

import numpy as np
import pytest

from helpers import distorted, grid_problem, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POINTS = np.linspace(-1, 1, 10001)   # range(-1, 1; length=10001), runtests.jl:55


def approx(a, b):   # Julia's default isapprox on arrays: norm(a-b) <= sqrt(eps)*max(norm(a), norm(b))
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) <= np.sqrt(np.finfo(float).eps) * max(np.linalg.norm(a), np.linalg.norm(b))


def fns1(x): return 0.5 * np.array([1 - x, 1 + x])
def dfns1(x): return 0.5 * np.array([-1.0, 1.0])
def fns2(x): return np.array([-0.5 * x * (1 - x), 1 - x ** 2, 0.5 * x * (1 + x)])
def dfns2(x): return np.array([x - 0.5, -2 * x, x + 0.5])
def fns3(x): return 9 / 16 * np.array([-(1 - x) * (1 / 3 + x) * (1 / 3 - x), 3 * (1 - x) * (1 + x) * (1 / 3 - x),
                                       3 * (1 - x) * (1 + x) * (1 / 3 + x), -(1 + x) * (1 / 3 + x) * (1 / 3 - x)])


class _OracleShape:
    def __init__(self, orc): self.o = orc
    def lagrange(self, p, kind):
        x = self.o.lagrange_nodes(p, kind); return (x, self.o.lagrange_weights(x))
    def N(self, s, xi): return self.o.shape_N(s[0], s[1], xi)
    def dN(self, s, xi): return self.o.shape_dN(s[0], s[1], xi)
    def nodes(self, s): return s[0]


class _ProductShape:
    def __init__(self, nls): self.n = nls
    def lagrange(self, p, kind): return self.n.lagrange(p, kind)
    def N(self, s, xi): return self.n.N(s, xi)
    def dN(self, s, xi): return self.n.dN(s, xi)
    def nodes(self, s): return s.nodes


@pytest.fixture(params=["oracle", "product-host"])
def shape_api(request, orc, nls):
    return _OracleShape(orc) if request.param == "oracle" else _ProductShape(nls)


# ---- runtests.jl:54-97 "ShapeFunctions" ---------------------------------------------------
@pytest.mark.parametrize("p", [1, 2, 3])
def test_lagrange_chebyshev2(shape_api, p):
    s = shape_api.lagrange(p, "chebyshev2")
    assert len(shape_api.nodes(s)) - 1 == p
    assert approx(shape_api.nodes(s), [np.cos(i * np.pi / p) for i in range(p, -1, -1)])
    pts = POINTS[::7] if p == 3 else POINTS
    Ns = np.array([shape_api.N(s, x) for x in pts])
    assert approx(Ns.sum(axis=1), np.ones(len(pts)))                  # partition of unity
    if p <= 2:
        f, df = (fns1, dfns1) if p == 1 else (fns2, dfns2)
        assert approx(Ns, np.array([f(x) for x in pts]))
        assert approx(np.array([shape_api.dN(s, x) for x in pts]), np.array([df(x) for x in pts]))


@pytest.mark.parametrize("p", [1, 2, 3])
def test_lagrange_equidistant(shape_api, p):
    s = shape_api.lagrange(p, "equidistant")
    assert approx(shape_api.nodes(s), np.linspace(-1, 1, p + 1))
    pts = POINTS[::5]
    Ns = np.array([shape_api.N(s, x) for x in pts])
    assert approx(Ns.sum(axis=1), np.ones(len(pts)))
    f = {1: fns1, 2: fns2, 3: fns3}[p]
    assert approx(Ns, np.array([f(x) for x in pts]))
    if p <= 2:
        df = dfns1 if p == 1 else dfns2
        assert approx(np.array([shape_api.dN(s, x) for x in pts]), np.array([df(x) for x in pts]))


def test_kat4_chebyshev_p2_rounding(shape_api):
    """KAT-4: cos(pi/2) != 0, so xi = 0.0 takes the general (off-node) branch."""
    k = json.load(open(os.path.join(GOLD, "reference_kats.json")))["KAT4"]
    s = shape_api.lagrange(2, "chebyshev2")
    assert shape_api.nodes(s)[1] == k["center_node"]
    w = s[1] if isinstance(s, tuple) else s.w
    assert list(w) == k["weights"]
    assert list(shape_api.dN(s, 0.0)) == k["dN_at_0"]


def test_oracle_and_product_tables_bitwise_equal(orc, nls):
    """The product tabulates its shape data on the host with its own code; it must match the
    oracle bit for bit (both restate shapefunctions.jl / quadrature.jl)."""
    for p in range(1, 8):
        for kind in ("chebyshev2", "equidistant"):
            x = orc.lagrange_nodes(p, kind); w = orc.lagrange_weights(x)
            s = nls.lagrange(p, kind)
            assert np.array_equal(x, s.nodes) and np.array_equal(w, s.w)
            for xi in list(np.linspace(-1, 1, 41)) + list(x):
                assert np.array_equal(orc.shape_N(x, w, xi), nls.N(s, xi))
                assert np.array_equal(orc.shape_dN(x, w, xi), nls.dN(s, xi))
    for q in range(1, 6):
        xo, wo = orc.gauss(q); g = nls.gauss_quadrature(q)
        assert np.array_equal(xo, g.points) and np.array_equal(wo, g.weights)


# ---- runtests.jl:99-137 "Quadrature" -----------------------------------------------------
def test_gauss_exact(orc, nls):
    polys = {1: (lambda x: x - 1, -2.0), 2: (lambda x: x**3 - 2 * x**2 + 3 * x - 4, -28 / 3),
             3: (lambda x: x**5 - 2 * x**4 + 3 * x**3 - 4 * x**2 + 5 * x - 6, -232 / 15),
             4: (lambda x: x**7 - 2 * x**6 + 3 * x**5 - 4 * x**4 + 5 * x**3 - 6 * x**2 + 7 * x - 8, -776 / 35),
             5: (lambda x: x**9 - 2 * x**8 + 3 * x**7 - 4 * x**6 + 5 * x**5 - 6 * x**4 + 7 * x**3 - 8 * x**2 + 9 * x - 10, -9236 / 315)}
    for q, (f, exact) in polys.items():
        xo, _ = orc.gauss(q)
        assert np.isclose(orc.integrate(q, [f(x) for x in xo]), exact, rtol=1.5e-8, atol=0)
        assert np.isclose(nls.integrate(nls.gauss_quadrature(q), f), exact, rtol=1.5e-8, atol=0)
    with pytest.raises(ValueError):
        orc.gauss(6)
    with pytest.raises(ValueError):
        nls.gauss_quadrature(6)


def test_gauss_nonlinear(orc, nls):
    f1, e1 = (lambda x: np.tanh(1 + x)), -2 - np.log(2) + np.log(1 + np.exp(4))
    f2, e2 = (lambda x: np.sin(np.exp(x ** 2))), 1.77247907969601871352278360619085214927452859669568635245
    for q, r1, r2 in [(3, 5e-4, 5e-2), (4, 5e-5, 5e-3), (5, 5e-6, 6e-4)]:
        Q = nls.gauss_quadrature(q)
        assert abs(nls.integrate(Q, f1) - e1) <= r1 * abs(e1)
        assert abs(nls.integrate(Q, f2) - e2) <= r2 * abs(e2)
        xo, _ = orc.gauss(q)
        assert abs(orc.integrate(q, f1(xo)) - e1) <= r1 * abs(e1)
    # vector-valued integrands and :index mode (quadrature.jl:58-65)
    Q = nls.gauss_quadrature(3)
    v = nls.integrate(Q, lambda x, i: np.array([1.0, x, float(i)]), "index")
    assert np.allclose(v, [2.0, 0.0, 0 * 5 / 9 + 1 * 8 / 9 + 2 * 5 / 9])


# ---- runtests.jl:139-207 "Boundaries" (0-based here) ----------------------------------------
def test_boundaries(nls):
    bc = nls.Dirichlet([0, 2], [1.0, 3.0])
    vec = np.zeros(4); bc.apply(vec)
    assert approx(vec, [1.0, 0.0, 3.0, 0.0])
    el = nls.Element([1, 2], [0.0, 0.5, 1.0])
    for b in (nls.Dirichlet([0, 2], [1.0, 3.0]), nls.Dirichlet([2], [3.0])):
        eb = nls.ElementBoundary(el, b)
        assert list(eb.nodes) == [1] and list(eb.values) == [3.0] and eb.element is el
        ev = np.zeros(2); eb.apply(ev)
        assert approx(ev, [0.0, 3.0])
    tb = nls.TimeDependent(nls.Dirichlet([0, 2], [0.0, 0.0]), lambda _, t: [1.0 * t, 3.0 * t], pass_context=True)
    vec = np.zeros(4); tb.apply(vec); assert approx(vec + 1, np.ones(4))
    tb.update(0.5); tb.apply(vec); assert approx(vec, [0.5, 0.0, 1.5, 0.0])
    tb.update(1.0); tb.apply(vec); assert approx(vec, [1.0, 0.0, 3.0, 0.0])
    el = nls.Element([1, 2], [0.0, 1 / 3, 2 / 3, 1.0])
    tb = nls.TimeDependent(nls.Dirichlet([0, 2], [0.0, 0.0]), lambda t: [1.0 * t, 3.0 * t])
    eb = nls.ElementBoundary(el, tb)
    assert list(eb.nodes) == [1] and list(eb.values) == [0.0]
    ev = np.zeros(2); eb.apply(ev); assert approx(ev + 1, [1.0, 1.0])
    tb.update(0.5); eb.apply(ev); assert approx(ev, [0.0, 1.5])
    tb.update(1.0); eb.apply(ev); assert approx(ev, [0.0, 3.0])
    nb = nls.Neumann([1, 1], [0.5, 0.25]); F = np.zeros(3); nb.apply(F); assert F[1] == 0.75


# ---- KAT-1/2: element closed forms (defaults.jl:2-26) ---------------------------------------
@pytest.mark.parametrize("p,q", [(1, 1), (1, 2), (2, 2), (2, 3), (3, 4), (5, 5)])
def test_kat1_linear_element(orc, nls, p, q):
    E, A, L = 3.0, 0.25, 0.8
    mesh = nls.bar_mesh(1, p, length=L)
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_LINEAR, mat=(E,), area=A)
    u = 0.01 * (orc.lagrange_nodes(p) + 1) / 2        # linear in xi at the Chebyshev-2 nodes: constant strain
    fe, Ke = om.element(0, u)
    eps = 0.01 / L
    assert np.allclose(Ke @ u, fe, rtol=1e-12, atol=1e-15)
    assert np.isclose(fe[0], -E * eps * A, rtol=1e-12) and np.isclose(fe[-1], E * eps * A, rtol=1e-12)
    assert np.abs(fe[1:-1]).max(initial=0) <= 1e-14     # interior nodes carry no net force
    if p == 1:
        assert np.allclose(Ke, E * A / L * np.array([[1, -1], [-1, 1]]), rtol=1e-13)
    assert np.allclose(Ke.sum(axis=1), 0, atol=1e-12)   # rigid translation


def test_kat2_exponential_element(orc, nls):
    s_sat, b, A, L = 1.3, 2.2, 0.6, 0.5
    mesh = nls.bar_mesh(1, 1, length=L)
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=1, nq=2, mat_kind=orc.MAT_EXPONENTIAL, mat=(s_sat, b), area=A)
    u = np.array([0.01, 0.04]); eps = (u[1] - u[0]) / L
    fe, Ke = om.element(0, u)
    assert np.allclose(fe, s_sat * (1 - np.exp(-b * eps)) * A * np.array([-1, 1]), rtol=1e-13)
    assert np.allclose(Ke, b * s_sat * np.exp(-b * eps) * (A / L) * np.array([[1, -1], [-1, 1]]), rtol=1e-13)


def test_kat3_newton(orc, nls):
    k = json.load(open(os.path.join(GOLD, "reference_kats.json")))["KAT3"]
    for nel, p, q in k["meshes"]:
        mesh = nls.bar_mesh(nel, p)
        om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_EXPONENTIAL,
                             mat=(k["sigma_sat"], k["b"]), area=k["A"])
        om.set_dirichlet([0], [0.0]); om.set_neumann([mesh.nn - 1], [k["tip_force"]])
        r = om.newton(rtol=k["rtol"])
        assert r["reason"] == k["reason"] and r["iterations"] == k["iterations"]
        assert abs(r["U"][-1] - k["tip"]) <= 2e-15
        assert np.allclose(r["U"], mesh.coords.ravel() * np.log(2) / 2, atol=1e-14)   # exact u = x ln2/2
        assert np.allclose(r["hist"][:5], k["hist"], rtol=1e-7)
        assert r["hist"][5] <= 6e-15
        # Krylov variant gives the same iterate sequence
        r2 = om.newton(rtol=k["rtol"], linear="pcg")
        assert r2["iterations"] == k["iterations"] and np.allclose(r2["U"], r["U"], atol=1e-14)


def test_kat5_plasticity(orc, nls):
    E, Hk, Ha = 100.0, 10.0, 5.0
    Eep = E * (Hk + Ha) / (E + Hk + Ha)
    mesh = nls.bar_mesh(3, 1)
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=1, nq=2, mat_kind=orc.MAT_PLASTIC,
                         mat=(E, Hk, Ha, 1.0, 0.0, Eep), area=1.0)
    om.set_dirichlet([0], [0.0]); om.set_neumann([3], [0.9])          # below yield: elastic
    r = om.newton(rtol=1e-12, atol=1e-13, maxits=10)
    assert r["reason"].startswith("CONVERGED") and np.allclose(r["U"], 0.9 / E * mesh.coords.ravel(), rtol=1e-12)
    assert np.all(om.state("dsig") == E) and np.all(om.state("gam") == 0)
    om.commit_state(r["U"])
    om.set_neumann([3], [1.3])                                        # above yield: tangent Eep
    r = om.newton(U0=r["U"], rtol=1e-12, atol=1e-13, maxits=10)
    assert r["reason"].startswith("CONVERGED") and np.all(om.state("dsig") == Eep)
    assert np.allclose(om.state("sig"), 1.3, rtol=1e-12)
    dg = (1.3 - 1.0) / (Hk + Ha)        # consistency: sigma = kappa + alpha after return mapping
    assert np.allclose(om.state("gam"), dg, rtol=1e-10)
    assert np.allclose(om.state("kap") + om.state("alp"), 1.3, rtol=1e-12)


def test_newton_semantics_oracle(orc, nls):
    mesh = nls.bar_mesh(8, 2)
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=2, nq=3, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=1.0)
    om.set_dirichlet([0], [0.0]); om.set_neumann([mesh.nn - 1], [0.5])
    assert om.newton(rtol=1e-10, maxits=5)["reason"] == "DIVERGED_MAXITS"      # the maxits quirk
    assert om.newton(rtol=1e-10, maxits=6)["reason"] == "CONVERGED_RELATIVE"
    r = om.newton(rtol=1e-10, dtol=1e-3)
    assert r["reason"] == "DIVERGED_STEP" and r["iterations"] == 0
    assert om.newton(rtol=1e-30, atol=1e-9)["reason"] == "CONVERGED_ABSOLUTE"
    assert om.newton()["reason"] == "DIVERGED_MAXITS"                          # default 1e-16 tolerances
    om2 = orc.OracleModel(1, mesh.conn, mesh.coords, p=2, nq=3, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=1.0)
    om2.set_neumann([mesh.nn - 1], [0.5])                                      # no Dirichlet: singular K
    # in floating point the pivot is tiny rather than exactly 0 (LAPACK behaves the same): either the
    # solve is flagged singular or the step explodes past dtol
    assert om2.newton(rtol=1e-10)["reason"] in ("DIVERGED_LINEAR_SOLVE", "DIVERGED_STEP")
    rc, _ = orc.dense_solve([[1.0, 2.0], [2.0, 4.0]], [1.0, 1.0])
    assert rc == 1
    rc, x = orc.dense_solve([[0.0, 2.0], [3.0, 1.0]], [4.0, 5.0])              # needs pivoting
    assert rc == 0 and np.allclose(x, [1.0, 2.0])


# ---- T2: Q1 Neo-Hookean pinned by autograd ---------------------------------------------------
def _q1_table(dim):
    g = 1 / np.sqrt(3)
    out = []
    for q in range(2 ** dim):
        xi = [(-g if ((q >> d) & 1) == 0 else g) for d in range(dim)]
        dN = np.zeros((2 ** dim, dim))
        for a in range(2 ** dim):
            s = [(-1 if ((a >> d) & 1) == 0 else 1) for d in range(dim)]
            for D in range(dim):
                v = 1.0
                for d in range(dim):
                    v *= (0.5 * s[d]) if d == D else 0.5 * (1 + s[d] * xi[d])
                dN[a, D] = v
        out.append(dN)
    return out


def _energy(U, X, conn, dim, mu, lam):
    import torch
    tot = 0
    Xt, cn = torch.tensor(X), torch.tensor(conn.astype(np.int64))
    u = U.reshape(-1, dim)
    for dN in _q1_table(dim):
        dN = torch.tensor(dN)
        J = torch.einsum("ead,aD->edD", Xt[cn], dN)
        G = torch.einsum("aE,eED->eaD", dN, torch.linalg.inv(J))
        F = torch.eye(dim, dtype=torch.float64) + torch.einsum("eai,eaJ->eiJ", u[cn], G)
        Jd = torch.linalg.det(F)
        I1 = (F * F).sum((1, 2)) + (1.0 if dim == 2 else 0.0)        # plane strain: F33 = 1
        psi = mu / 2 * (I1 - 3) - mu * torch.log(Jd) + lam / 2 * torch.log(Jd) ** 2
        tot = tot + (psi * torch.linalg.det(J)).sum()
    return tot


@pytest.mark.parametrize("dim,n", [(2, 4), (3, 3)])
def test_neohookean_matches_autograd(orc, nls, dim, n):
    import torch
    mu, lam = 0.4, 0.6
    mesh = nls.grid_mesh(n, dim)
    X = distorted(mesh)
    om = orc.OracleModel(dim, mesh.conn, X, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mu, lam))
    U = nls.smooth_field(X, 0.05)
    Ut = torch.tensor(U, requires_grad=True)
    g, = torch.autograd.grad(_energy(Ut, X, mesh.conn, dim, mu, lam), Ut, create_graph=True)
    H = torch.stack([torch.autograd.grad(g[i], Ut, retain_graph=True)[0] for i in range(len(U))]).numpy()
    _, R = om.residual(U)
    K = om.jacobian_dense(U)
    assert rel_err(R, g.detach().numpy()) <= 1e-12
    assert rel_err(K, H) <= 1e-12
    assert np.abs(K - K.T).max() <= 1e-14 * np.abs(K).max()           # symmetry
    rp, ci = om.pattern()
    vals = om.jacobian_csr(U)
    dense = np.zeros_like(K)
    for i in range(len(U)):
        dense[i, ci[rp[i]:rp[i + 1]]] = vals[rp[i]:rp[i + 1]]
    assert np.array_equal(dense, K)                                    # nothing outside the pattern, same sums
    assert np.count_nonzero(K) <= len(ci)


def test_neohookean_patch_and_small_strain(orc, nls):
    """Uniform F reproduces the analytic P (patch test); small strain reduces to linear elasticity."""
    mu, lam = 0.5, 0.75
    mesh = nls.grid_mesh(4, 3)
    X = distorted(mesh, 0.2, seed=3)
    om = orc.OracleModel(3, mesh.conn, X, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mu, lam))
    Hm = np.array([[0.05, 0.02, 0.0], [-0.01, 0.03, 0.04], [0.0, 0.01, -0.02]])
    U = (X @ Hm.T).ravel()
    _, R = om.residual(U)
    interior = np.all((mesh.coords > 1e-9) & (mesh.coords < 1 - 1e-9), axis=1)
    assert np.abs(R.reshape(-1, 3)[interior]).max() <= 1e-14           # divergence-free stress
    F = np.eye(3) + Hm
    P = mu * (F - np.linalg.inv(F).T) + lam * np.log(np.linalg.det(F)) * np.linalg.inv(F).T
    top = np.isclose(mesh.coords[:, 2], 1.0)
    assert np.allclose(R.reshape(-1, 3)[top].sum(axis=0), P @ np.array([0, 0, 1.0]), rtol=1e-12)
    K0 = om.jacobian_dense(np.zeros(om.ndof))
    eps_u = 1e-6 * (X @ Hm.T).ravel()
    _, Rs = om.residual(eps_u)
    assert np.allclose(K0 @ eps_u, Rs, rtol=0, atol=1e-6 * np.abs(Rs).max())   # O(strain) linearisation error


def test_oracle_regression_vectors(orc, nls):
    """Committed oracle outputs (tests/golden/oracle_cases.json) still reproduce bit for bit."""
    cases = json.load(open(os.path.join(GOLD, "oracle_cases.json")))
    for name, c in cases.items():
        if c["dim"] == 1:
            mesh = nls.bar_mesh(c["nel"], c["p"])
            om = orc.OracleModel(1, mesh.conn, mesh.coords, p=c["p"], nq=c["q"], mat_kind=orc.MAT_EXPONENTIAL,
                                 mat=(1.0, 2.0), area=c["area"])
            om.set_dirichlet([0], [0.0]); om.set_neumann([mesh.nn - 1], [0.5])
        else:
            conn = np.array(c["conn"], dtype=np.int32).reshape(-1, 2 ** c["dim"])
            om = orc.OracleModel(c["dim"], conn, np.array(c["coords"]), mat_kind=orc.MAT_NEOHOOKEAN, mat=(c["mu"], c["lam"]))
            om.set_dirichlet(c["dir_dofs"], c["dir_vals"])
        rp, ci = om.pattern()
        assert rp.tolist() == c["rowptr"] and ci.tolist() == c["colind"], name
        Uo, R = om.residual(np.array(c["U"]))
        assert Uo.tolist() == c["U_out"] and R.tolist() == c["R"], name
        assert om.jacobian_csr(np.array(c["U"])).tolist() == c["vals"], name


def test_threaded_oracle_matches_serial(orc, nls):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    _, om = grid_problem(nls, orc, 12, 2, mat, compress=-0.01)
    U = nls.smooth_field(om.coords)
    _, R1 = om.residual(U); K1 = om.jacobian_csr(U)
    om.set_threads(4)
    _, R4 = om.residual(U); K4 = om.jacobian_csr(U)
    assert rel_err(R4, R1) <= 1e-14 and rel_err(K4, K1) <= 1e-14
    b = np.random.default_rng(0).standard_normal(om.ndof)
    rc4, x4, it4, _ = om.pcg(K1, b)
    om.set_threads(1)
    rc1, x1, it1, _ = om.pcg(K1, b)
    assert rc1 == rc4 == 0 and abs(it1 - it4) <= 1 and rel_err(x4, x1) <= 1e-9


def test_distributed_force_oracle(orc, nls):
    """DistributedForce (gallery/distributedforce.jl:7-27): total load, and the exact solution of a
    uniformly loaded linear bar u(x) = f/(E A) (L x - x^2/2), nodally exact for p >= 2."""
    E, A, f, L = 3.0, 0.5, 1.25, 1.0
    for p, q in [(2, 3), (3, 4)]:
        mesh = nls.bar_mesh(6, p)
        om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_LINEAR, mat=(E,), area=A)
        om.set_dirichlet([0], [0.0])
        om.set_body_force([f])
        _, R = om.residual(np.zeros(mesh.nn))
        assert abs(-R.sum() - f * L) <= 1e-14
        # element nodes sit at the Chebyshev-2 points of each element (geometry is affine between end nodes)
        r = om.newton(rtol=1e-12, atol=1e-14, maxits=5)
        assert r["reason"].startswith("CONVERGED") and r["iterations"] == 1
        ends = mesh.conn[:, [0, -1]].ravel()
        x = mesh.coords.ravel()[ends]
        assert np.allclose(r["U"][ends], f / (E * A) * (L * x - x ** 2 / 2), rtol=1e-12, atol=1e-15)
    m2 = nls.grid_mesh(5, 3)
    om = orc.OracleModel(3, m2.conn, distorted(m2), mat_kind=orc.MAT_NEOHOOKEAN, mat=(1.0, 1.0))
    om.set_body_force([0.5, -2.0, 0.25])
    _, R = om.residual(np.zeros(m2.nn * 3))
    assert np.allclose(-R.reshape(-1, 3).sum(axis=0), [0.5, -2.0, 0.25], rtol=1e-13)   # unit cube volume


# ---- dynamics: consistent mass and Newmark restatement ---------------------------------------
def test_mass_known_answers(orc):
    """mass.jl:2-17 on one 2-node element: rho A L / 6 [[2,1],[1,2]]; any p with exact quadrature: entries sum to
    rho A L (partition of unity); Q1: sum of the node-level mass = rho |Omega|."""
    conn = np.array([[0, 1]], dtype=np.int32)
    om = orc.OracleModel(1, conn, np.array([0.0, 2.0]), p=1, nq=2, mat_kind=orc.MAT_LINEAR, mat=(1.0,), area=1.0)
    assert np.allclose(om.mass_dense(3.0, 0.5), 3.0 * 0.5 * 2.0 / 6.0 * np.array([[2, 1], [1, 2]]), rtol=1e-15, atol=0)
    for p, q in [(2, 3), (3, 4), (4, 5)]:
        nel = 3
        nn = nel * p + 1
        conn = np.array([[e * p + a for a in range(p + 1)] for e in range(nel)], dtype=np.int32)
        X = np.linspace(0.0, 1.5, nn)
        om = orc.OracleModel(1, conn, X, p=p, nq=q, mat_kind=orc.MAT_LINEAR, mat=(1.0,), area=1.0)
        M = om.mass_dense(2.0, 0.25)
        assert abs(M.sum() - 2.0 * 0.25 * 1.5) <= 1e-14 and np.allclose(M, M.T, atol=1e-16)
    n = 4
    g = np.arange(n * n).reshape(n, n)
    conn = np.array([[g[j, i], g[j, i + 1], g[j + 1, i], g[j + 1, i + 1]] for j in range(n - 1) for i in range(n - 1)], dtype=np.int32)
    xs = np.linspace(0, 1, n)
    X = np.array([[x, y] for y in xs for x in xs])
    om = orc.OracleModel(2, conn, X, mat_kind=orc.MAT_NEOHOOKEAN, mat=(1.0, 1.0))
    M = om.mass_dense(1.5)
    assert abs(M.sum() - 1.5 * 2) <= 1e-14 and np.allclose(M, M.T, atol=1e-16)


def test_newmark_linear_energy_conservation(orc):
    """Average-acceleration Newmark (beta = 1/4, gamma = 1/2) conserves 1/2 v'Mv + 1/2 u'Ku exactly for a linear
    undamped system; the restated stepper (newmark.jl:109-155) must too, and the first Newton iteration must
    converge the linear problem (1 iteration per step)."""
    nel, E, rho, A = 5, 4.0, 1.5, 1.0
    conn = np.array([[e, e + 1] for e in range(nel)], dtype=np.int32)
    X = np.linspace(0.0, 1.0, nel + 1)
    om = orc.OracleModel(1, conn, X, p=1, nq=2, mat_kind=orc.MAT_LINEAR, mat=(E,), area=A)
    om.set_dirichlet([0], [0.0])
    # released from rest position with an initial velocity (zero at the clamp): R(U0) = 0, so a0 = 0 on every DoF and the
    # clamped entry of the acceleration, which the reference never solves for, stays zero
    out = om.newmark(rho, 0.01, 0.2, area=A, Ud0=0.01 * X, rtol=1e-9, atol=1e-14, maxits=5)
    assert out["reason"] == "CONVERGED" and out["steps"] == 20
    assert np.all(out["num_its"][1:] == 1)
    M = om.mass_dense(rho, A)
    K = om.jacobian_dense(np.zeros(nel + 1))
    free = np.arange(1, nel + 1)
    en = [0.5 * v[free] @ M[np.ix_(free, free)] @ v[free] + 0.5 * u[free] @ K[np.ix_(free, free)] @ u[free]
          for u, v in zip(out["U"], out["Ud"])]
    assert max(en) - min(en) <= 1e-12 * en[0] and en[0] > 0


def test_arclength_restatement_snap_through(orc):
    """newtonarclength_fem.jl restated: on a shallow arch the load factor passes a limit point, decreases on the unstable
    branch and rises again; the reference's det(K) sign rule and the stiffness-parameter rule (what the device path uses)
    give the same path; every converged state satisfies lambda F_ext = F_int(d) on the free DoFs."""
    import __graft_entry__ as g
    nls = g.load_package()
    nx, ny, R, th0, t = 25, 3, 10.0, 0.25, 0.08
    m = nls.grid_mesh((nx, ny), 2)
    s, yy = m.coords[:, 0], m.coords[:, 1] * (nx - 1) / (ny - 1)
    th, r = (2 * s - 1) * th0, R + (yy - 0.5) * t
    X = np.stack([r * np.sin(th), r * np.cos(th) - R * np.cos(th0)], axis=1)
    E, nu = 1000.0, 0.3
    om = orc.OracleModel(2, m.conn, X, mat_kind=orc.MAT_NEOHOOKEAN, mat=(E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))))
    ends = np.concatenate([np.arange(ny) * nx, np.arange(ny) * nx + nx - 1])
    dofs = (ends[:, None] * 2 + np.arange(2)[None, :]).ravel()
    om.set_dirichlet(dofs, np.zeros(len(dofs)))
    apex = (ny - 1) * nx + nx // 2
    F = np.zeros(om.ndof); F[2 * apex + 1] = -0.05
    kw = dict(b=0.5, da=0.4, rtol=1e-7, ftol=1e-7, maxsteps=24, maxits=30)
    a = orc.newtonarclength(om, F, sign_rule="det", **kw)
    b = orc.newtonarclength(om, F, sign_rule="stiffness", **kw)
    assert a["num_steps"] == b["num_steps"] == 24 and np.array_equal(a["num_its"], b["num_its"])
    assert np.array_equal(a["lam"], b["lam"]) and np.array_equal(a["d"], b["d"])
    lam = a["lam"]
    k = int(np.argmax(lam[:14]))
    assert 2 < k < 13 and lam[k + 4] < lam[k] and np.all(a["num_its"][1:] < 30), (k, lam)
    free = ~om.constrained_mask().astype(bool)
    for s_ in (k, k + 4, 23):
        _, Fi = om.residual(a["d"][s_])
        assert np.linalg.norm((lam[s_] * F - Fi)[free]) <= 1e-6 * np.linalg.norm(F)


def test_dynamics_golden_regression(orc):
    """The committed dynamics / continuation vectors (tests/golden/oracle_dynamics_cases.json) are what the oracle produces."""
    import json
    import os
    import __graft_entry__ as g
    nls = g.load_package()
    here = os.path.dirname(os.path.abspath(__file__))
    dyn = json.load(open(os.path.join(here, "golden", "oracle_dynamics_cases.json")))
    cases = json.load(open(os.path.join(here, "golden", "oracle_cases.json")))
    c = cases[dyn["mass_q1_2d_n4_distorted"]["case"]]
    om = orc.OracleModel(2, np.array(c["conn"], dtype=np.int32).reshape(-1, 4), np.array(c["coords"]).reshape(-1, 2),
                         mat_kind=orc.MAT_NEOHOOKEAN, mat=(c["mu"], c["lam"]))
    assert np.allclose(om.mass_csr(dyn["mass_q1_2d_n4_distorted"]["rho"]), dyn["mass_q1_2d_n4_distorted"]["vals"], rtol=1e-14, atol=0)
    nm = dyn["newmark_bar_exp_p2_q3"]
    mesh = nls.bar_mesh(nm["nel"], nm["p"])
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=nm["p"], nq=nm["q"], mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=nm["area"])
    om.set_dirichlet([0], [0.0]); om.set_neumann([mesh.nn - 1], [0.0])
    out = om.newmark(nm["rho"], nm["dt"], nm["maxtime"], area=nm["area"], bc_update=lambda t: (None, [nm["tip_rate"] * t]),
                     rtol=1e-11, atol=1e-13, maxits=20)
    assert out["num_its"].tolist() == nm["num_its"] and np.allclose(out["U"], nm["U"], rtol=1e-12, atol=1e-16)


def test_j2_oracle_known_answers():
    """Pins for the J2 restatement (no reference counterpart): yield strain of uniaxial strain, stress on the yield surface
    after the return, consistent tangent = finite difference of the internal force, elastic unloading after commit."""
    import importlib
    orc = importlib.import_module("oracle.oracle")
    E, nu, sy, H = 100.0, 0.3, 0.5, 10.0
    K, G = E / (3 * (1 - 2 * nu)), E / (2 * (1 + nu))
    X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]], dtype=float)
    o = orc.J2Oracle(3, np.arange(8)[None, :], X, K, G, sy, H)
    for ezz, plastic in ((0.9 * sy / (2 * G), False), (3.0 * sy / (2 * G), True)):
        U = np.zeros((8, 3)); U[:, 2] = ezz * X[:, 2]
        o.residual(U.ravel())
        assert bool(o.alpha.max() > 0) == plastic
        eps = np.zeros((3, 3)); eps[2, 2] = ezz
        ee = eps - o.epsp[0, 0]
        s = 2 * G * (ee - np.trace(ee) / 3 * np.eye(3))
        if plastic:      # on the (hardened) yield surface
            assert abs(np.sqrt((s * s).sum()) - np.sqrt(2.0 / 3.0) * (sy + H * o.alpha[0, 0])) <= 1e-13
            assert abs(np.trace(o.epsp[0, 0])) <= 1e-16          # plastic flow is isochoric
    rng = np.random.default_rng(1)
    U = 0.02 * rng.standard_normal(24)
    _, R = o.residual(U)
    Kd = o.jacobian_dense(U)
    Kfd = np.zeros((24, 24))
    for j in range(24):
        Up, Um = U.copy(), U.copy()
        Up[j] += 1e-7; Um[j] -= 1e-7
        Kfd[:, j] = (o.residual(Up)[1] - o.residual(Um)[1]) / 2e-7
    assert np.abs(Kd - Kfd).max() <= 1e-8 * np.abs(Kd).max() and np.abs(Kd - Kd.T).max() <= 1e-14 * np.abs(Kd).max()
    o.residual(U); o.commit_state()
    o.residual(0.999 * U)                                        # small step back: elastic
    assert np.array_equal(o.alpha, o.alpha_n)
