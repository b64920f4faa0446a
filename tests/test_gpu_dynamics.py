"""SURVEY 8f-3 on the device: consistent mass (gallery/mass.jl), the dynamic Residual/Jacobian methods
(gallery/residual.jl:48-54, gallery/jacobian.jl:42-51) and NewmarkTimeStepper (timesteppers/newmark.jl), against the
oracle's restatement on the same inputs. Tolerances: operator entries 1e-12 relative, step counts exact,
histories 1e-10 relative (north_star)."""
import ctypes as C

import numpy as np
import pytest

from helpers import RTOL, UTOL, distorted, grid_problem, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,q", [(1, 2), (2, 3), (3, 5)])
def test_mass_bar(nls, orc, p, q):
    rho, A = 2.5, 0.7
    mat = nls.LinearElasticity1D(E=3.0, rho=rho)
    mesh = nls.bar_mesh(7, p)
    mesh.coords[:] = mesh.coords ** 1.3            # non-uniform element lengths
    fem = nls.FEMModel(mesh, q, p, data={"A": A})
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_LINEAR, mat=(3.0,), area=A)
    M = nls.CSRMatrix(fem)
    nls.assemblemass(mat, fem, M)
    Mo = om.mass_csr(rho, A)
    assert rel_err(M.vals, Mo) <= RTOL
    if q > p:       # exact quadrature: total mass = rho A L
        assert abs(M.vals.sum() - rho * A * (mesh.coords[-1] - mesh.coords[0])) <= 1e-13 * rho * A
    x = np.random.default_rng(1).standard_normal(fem.ndof)
    y = np.linspace(-1, 1, fem.ndof)
    y0 = y.copy()
    nls.applymass(mat, fem, x, y)                  # R += M a
    assert rel_err(y - y0, om.mass_dense(rho, A) @ x) <= RTOL


@pytest.mark.parametrize("dim,n", [(2, 9), (3, 6)])
def test_mass_and_dynamic_operators_q1(nls, orc, dim, n):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mat.rho = 1.7
    fem, om = grid_problem(nls, orc, n, dim, mat, compress=-0.1 / (n - 1))
    X = distorted(fem.mesh); fem.mesh.coords[:] = X; om.coords[:] = X
    M = nls.CSRMatrix(fem)
    nls.assemblemass(mat, fem, M)
    assert rel_err(M.vals, om.mass_csr(1.7)) <= RTOL
    assert abs(M.vals.sum() - 1.7 * dim) <= 1e-12 * dim       # sum_ab m_ab = rho |Omega|, once per component
    rng = np.random.default_rng(2)
    U = nls.smooth_field(X, amp=0.1 / (n - 1))
    a = rng.standard_normal(fem.ndof)
    R, Ud = np.zeros(fem.ndof), U.copy()
    nls.Residual(mat)(fem, Ud, np.zeros(fem.ndof), a, R, None)          # dynamic method
    Uo, Ro = om.dynamic_residual(U, a, 1.7)
    assert np.array_equal(Ud, Uo) and rel_err(R, Ro) <= RTOL

    class Ctx:
        def dUdd_dU(self):
            return 0.37

        def time(self):
            return 0.0
    K = nls.CSRMatrix(fem)
    nls.Jacobian(mat)(fem, U.copy(), np.zeros(fem.ndof), a, K, Ctx())
    Ko = om.dynamic_jacobian_dense(U, 0.37, 1.7)
    rp, ci = K.rowptr, K.colind
    Kd = np.concatenate([Ko[i, ci[rp[i]:rp[i + 1]]] for i in range(fem.ndof)])
    assert rel_err(K.vals, Kd) <= RTOL


def _newmark(nls, fem, mat, dt, maxtime, **kw):
    s = nls.NewtonSolver(fem, rtol=1e-11, atol=1e-13, maxits=20, lin_rtol=1e-13)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    ts = nls.NewmarkTimeStepper(s, mat, dt=dt, maxtime=maxtime, **kw)
    nls.set_ctx(s, ts)
    return ts


def test_newmark_exponential_bar(nls, orc):
    """Bar clamped at node 0, tip force ramped in time (TimeDependent Neumann), nonlinear material: step-by-step
    parity of U, Ud, Udd, Newton iteration counts and residuals with the oracle."""
    p, q, nel, rho, A = 2, 3, 6, 2.0, 0.5
    mat = nls.ExponentialMaterial(1.0, 2.0, rho=rho)
    mesh = nls.bar_mesh(nel, p)
    fem = nls.FEMModel(mesh, q, p, data={"A": A})
    tip = mesh.nn - 1
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    fem.addboundary(nls.TimeDependent(nls.Neumann([tip], [0.0]), lambda t: [0.2 * t]))
    ts = _newmark(nls, fem, mat, dt=0.05, maxtime=0.5)
    ts.solve()
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=A)
    om.set_dirichlet([0], [0.0]); om.set_neumann([tip], [0.0])
    ref = om.newmark(rho, 0.05, 0.5, area=A, bc_update=lambda t: (None, [0.2 * t]), rtol=1e-11, atol=1e-13, maxits=20)
    assert ref["reason"] == "CONVERGED" and ts.step == ref["steps"] == 10
    assert np.array_equal(ts.num_its[:ts.step + 1, 0], ref["num_its"])
    for mine, theirs in [(ts.U, ref["U"]), (ts.Ud, ref["Ud"]), (ts.Udd, ref["Udd"])]:
        assert rel_err(mine[:ts.step + 1], theirs) <= UTOL
    assert np.abs(ts.U[-1]).max() > 1e-4          # something actually moved


def test_newmark_q1_initial_velocity(nls, orc):
    """2-D Neo-Hookean block, bottom clamped, released with an initial velocity field and a consistent initial
    acceleration (M a0 = -R(U0)); also the reference's literal a0 = 0 start."""
    n, dim = 6, 2
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mat.rho = 1.3
    mesh = nls.grid_mesh(n, dim)
    fem = nls.FEMModel(mesh)
    bottom = nls.face_nodes(n, dim, dim - 1, 0)
    db = (bottom[:, None] * dim + np.arange(dim)[None, :]).ravel()
    fem.addboundary(nls.Dirichlet(db, np.zeros(len(db))))
    om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet(db, np.zeros(len(db)))
    U0 = nls.smooth_field(mesh.coords, amp=0.01); U0[db] = 0.0
    V0 = 0.05 * nls.smooth_field(mesh.coords[:, ::-1].copy(), amp=1.0); V0[db] = 0.0
    for consistent in (True, False):
        ts = _newmark(nls, fem, mat, dt=0.02, maxtime=0.1, consistent_a0=consistent)
        ts.solve(U0, V0)
        ref = om.newmark(1.3, 0.02, 0.1, U0=U0, Ud0=V0, rtol=1e-11, atol=1e-13, maxits=20, consistent_a0=consistent)
        assert ref["reason"] == "CONVERGED" and ts.step == ref["steps"] == 5
        assert np.array_equal(ts.num_its[:ts.step + 1, 0], ref["num_its"])
        assert rel_err(ts.Udd[0], ref["Udd"][0]) <= 1e-9 if consistent else np.all(ts.Udd[0] == 0.0)
        for mine, theirs in [(ts.U, ref["U"]), (ts.Ud, ref["Ud"]), (ts.Udd, ref["Udd"])]:
            assert rel_err(mine[:ts.step + 1], theirs) <= 1e-9
