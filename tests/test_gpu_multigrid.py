"""Multigrid-preconditioned Krylov solve (csrc/kernels_mg.cu) on grid-topology 2-D and 3-D meshes: the solution of K x = b equals
the oracle's Jacobi-PCG solution of the same system (ragged grids whose node counts exercise the odd / even coarsening rule,
component-wise Dirichlet masks), with far fewer iterations; the Newton iterates do not depend on the preconditioner; the
single-precision smoother products are an option, not a change of result."""
import ctypes as C

import numpy as np
import pytest

from helpers import grid_problem, rel_err

pytestmark = pytest.mark.gpu
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))      # noqa: E731


def _solve(H, n, b, rtol=1e-12, maxits=20000):
    x = np.zeros(n)
    its, st, rr = C.c_int32(0), C.c_int32(0), C.c_double(0)
    H.call("nls_linear_solve", rtol, maxits, dp(b), dp(x), C.byref(its), C.byref(rr), C.byref(st))
    return x, its.value, st.value, rr.value


@pytest.mark.parametrize("shape", [(9, 9, 9), (12, 7, 10), (17, 10, 12), (6, 14, 21), (24, 24, 24), (9, 9, 25), (25, 31), (40, 18), (7, 50), (64, 64)])
def test_multigrid_solve_matches_oracle(nls, orc, shape):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, om = grid_problem(nls, orc, shape, len(shape), mat, compress=-0.01)
    vals = om.jacobian_csr(nls.smooth_field(fem.mesh.coords, amp=0.002))
    H = fem.handle(mat)
    H.call("nls_set_values", dp(vals))
    b = np.random.default_rng(5).standard_normal(fem.ndof)
    rc, xo, ito, _ = om.pcg(vals, b, rtol=1e-12, maxits=20000)
    assert rc == 0
    con = om.constrained_mask().astype(bool)
    res = {}
    for pc in (0, 1, 2):
        H.call("nls_set_option", b"precond", pc)
        H.call("nls_set_option", b"pcg_small", 0)
        x, its, st, rr = _solve(H, fem.ndof, b)
        assert st == 0 and rr <= 1e-12
        assert rel_err(x, xo) <= 1e-9
        assert np.all(x[con] == 0.0)
        res[pc] = its
    assert abs(res[0] - ito) <= 2
    assert res[2] <= 40 and res[2] * 3 <= res[0], "the V-cycle must cut the iteration count: %r" % (res,)
    # the FP32 smoother products are an implementation detail of the preconditioner
    H.call("nls_set_option", b"precond", 2)
    H.call("nls_set_option", b"mg_fp32", 0)
    x64, its64, st, _ = _solve(H, fem.ndof, b)
    assert st == 0 and rel_err(x64, xo) <= 1e-9 and abs(its64 - res[2]) <= 2
    # maxits exhaustion is reported
    _, its, st, _ = _solve(H, fem.ndof, b, maxits=2)
    assert st == 1 and its == 2


@pytest.mark.parametrize("shape", [(14, 11, 13), (45, 38)])
def test_newton_independent_of_preconditioner(nls, orc, shape):
    """Same Newton iteration count and displacement (1e-10 relative, the north star's bar) with Jacobi and with multigrid."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    sols = {}
    for pc in (0, 2):
        fem, om = grid_problem(nls, orc, shape, len(shape), mat, compress=-0.02 if len(shape) == 3 else -0.005, with_oracle=False)   # under half an element height
        H = fem.handle(mat)
        H.call("nls_set_option", b"precond", pc)
        H.call("nls_set_option", b"pcg_small", 0)
        H.call("nls_set_option", b"direct", 0)          # 6 006 DoF: Newton would otherwise take the exact banded solve
        s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-12)
        nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
        s.solve()
        sols[pc] = (s.U.copy(), s.iterations, repr(s.converged_reason), s.lin_iterations)
    assert sols[0][2].startswith("Converged") and sols[2][2].startswith("Converged")
    assert sols[0][1] == sols[2][1]
    assert rel_err(sols[2][0], sols[0][0]) <= 1e-10
    assert sols[2][3] * 3 <= sols[0][3]
