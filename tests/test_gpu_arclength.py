"""SURVEY 8f-4 on the device: Jacobi-preconditioned MINRES for the symmetric indefinite tangent and the arc-length
continuation of nonlinearsolvers/newtonarclength_fem.jl, against the oracle's dense restatement (bordered `\\`)."""
import ctypes as C

import numpy as np
import pytest

from helpers import grid_problem, rel_err

pytestmark = pytest.mark.gpu


def _arch(nls, nx=41, ny=3, R=10.0, th0=0.25, t=0.08):
    """Shallow circular arch (rise 0.31, thickness 0.08) meshed with nx x ny Q1 nodes; snaps through under an apex load."""
    m = nls.grid_mesh((nx, ny), 2)
    s, yy = m.coords[:, 0], m.coords[:, 1] * (nx - 1) / (ny - 1)
    th, r = (2 * s - 1) * th0, R + (yy - 0.5) * t
    X = np.stack([r * np.sin(th), r * np.cos(th) - R * np.cos(th0)], axis=1)
    ends = np.concatenate([np.arange(ny) * nx, np.arange(ny) * nx + nx - 1])
    dofs = (ends[:, None] * 2 + np.arange(2)[None, :]).ravel()
    apex = (ny - 1) * nx + nx // 2
    return nls.Mesh(2, m.conn, coords=X), dofs, apex


def test_minres_definite_and_indefinite(nls, orc):
    """MINRES equals the dense solve on an SPD tangent and on a symmetric indefinite one (K - shift * M-like diagonal)."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, om = grid_problem(nls, orc, 7, 2, mat, compress=-0.01)
    U = nls.smooth_field(fem.mesh.coords, amp=0.002)
    K = om.jacobian_dense(U)
    free = ~om.constrained_mask().astype(bool)
    rp, ci = om.pattern()
    rows = np.repeat(np.arange(om.ndof), np.diff(rp))
    H = fem.handle(mat)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(7)
    b = rng.standard_normal(om.ndof)
    ev = np.linalg.eigvalsh(K[np.ix_(free, free)])
    for shift in (0.0, 0.5 * (ev[2] + ev[3])):             # second case: three negative eigenvalues
        Ks = K - shift * np.eye(om.ndof)
        vals = Ks[rows, ci]
        H.call("nls_set_values", dp(np.ascontiguousarray(vals)))
        x = np.zeros(om.ndof)
        its, st, rr = C.c_int32(0), C.c_int32(0), C.c_double(0)
        H.call("nls_linear_solve_indefinite", 1e-13, 20000, dp(b), dp(x), C.byref(its), C.byref(rr), C.byref(st))
        xo = np.zeros(om.ndof); xo[free] = np.linalg.solve(Ks[np.ix_(free, free)], b[free])
        assert st.value == 0 and np.all(x[~free] == 0.0)
        assert rel_err(x, xo) <= 1e-8, (shift, rel_err(x, xo), its.value)
        if shift > 0:                                       # PCG must refuse this matrix, not return garbage
            H.call("nls_linear_solve", 1e-13, 20000, dp(b), dp(x), C.byref(its), C.byref(rr), C.byref(st))
            assert st.value != 0


def test_arclength_snap_through_arch(nls, orc):
    """The arch's load factor rises to a limit point, falls (indefinite tangent) and rises again; the device path follows
    the oracle step by step: same iteration counts, load factors and displacements."""
    mesh, dofs, apex = _arch(nls)
    E, nu = 1000.0, 0.3
    mat = nls.NeoHookean(E=E, nu=nu)
    fem = nls.FEMModel(mesh)
    fem.addboundary(nls.Dirichlet(dofs, np.zeros(len(dofs))))
    om = orc.OracleModel(2, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet(dofs, np.zeros(len(dofs)))
    F = np.zeros(fem.ndof); F[2 * apex + 1] = -0.05
    kw = dict(b=0.5, da=0.25, rtol=1e-6, ftol=1e-6, maxsteps=26, maxits=30)
    res = nls.newtonarclength(fem, mat, F, lin_rtol=1e-13, **kw)
    ref = orc.newtonarclength(om, F, sign_rule="det", **kw)
    assert res.reason == "completed" and res.num_steps == ref["num_steps"] == 26
    assert np.array_equal(res.num_its, ref["num_its"])
    assert rel_err(res.lam, ref["lam"]) <= 1e-8 and rel_err(res.d, ref["d"]) <= 1e-8
    lam = res.lam
    up1 = int(np.argmax(lam[:15]))
    assert 3 < up1 < 12 and lam[up1] > lam[up1 + 4] and lam[-1] > lam[up1 + 12]        # limit point, descent, second rise
    assert np.all(np.diff(res.d[:, 2 * apex + 1]) < 0)                                   # the apex keeps moving down
    assert res.lin_iterations > 0
