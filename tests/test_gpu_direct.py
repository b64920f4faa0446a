"""Exact banded LU for Newton's linear solves of small systems (csrc/kernels_direct.cu; replaces `dofs(K) \\ dofs(-R)`,
newton_fem.jl:158, at the reference's own problem sizes): residual histories that end at rounding level like the reference's
sparse direct solve, the same iterates as the Krylov path, and a zero pivot reported as DIVERGED_LINEAR_SOLVE."""
import numpy as np
import pytest

from helpers import bar_problem, grid_problem, rel_err

pytestmark = pytest.mark.gpu


def _newton(nls, fem, mat, direct, **kw):
    fem.handle(mat).call("nls_set_option", b"direct", direct)
    s = nls.NewtonSolver(fem, **kw)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve()
    return s


def test_kat3_history_ends_at_rounding_level(nls, orc):
    """With the exact solve the residual history follows the oracle's dense-LU Newton down to its floor: ~1e-16 on the small
    KAT-3 bar (10 p = 2 elements), ~3e-13 at C1 (100 p = 3 elements, 301 DoF: the floor of the reference algorithm itself at
    this conditioning -- the oracle stalls there too)."""
    mat = nls.ExponentialMaterial(1.0, 2.0)
    for nel, p, q, floor in ((10, 2, 3, 5e-15), (100, 3, 4, 1e-12)):
        fem, om = bar_problem(nls, orc, nel, p, q, mat)
        ref = om.newton(rtol=1e-16, atol=1e-16, maxits=8)
        s = _newton(nls, fem, mat, 1, rtol=1e-16, atol=1e-16, maxits=8)
        assert s.iterations == ref["iterations"] == 8
        assert s.lin_iterations == s.iterations              # one "iteration" per exact solve
        assert max(ref["hist"][-3:]) <= floor and max(s.norm_history[-3:]) <= 3 * max(ref["hist"][-3:]) + 1e-16
        np.testing.assert_allclose(s.norm_history[:5], ref["hist"][:5], rtol=1e-6, atol=10 * floor)
        assert rel_err(s.U, ref["U"]) <= 1e-12
        k = _newton(nls, fem, mat, 0, rtol=1e-10, maxits=25)
        assert k.lin_iterations > k.iterations and rel_err(k.U, s.U) <= 1e-9


@pytest.mark.parametrize("dim,n", [(2, 12), (3, 6)])
def test_direct_and_krylov_newton_agree(nls, orc, dim, n):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    sols = []
    for direct in (1, 0):
        fem, _ = grid_problem(nls, orc, n, dim, mat, compress=-0.05, with_oracle=False)
        s = _newton(nls, fem, mat, direct, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-13)
        assert repr(s.converged_reason).startswith("Converged")
        sols.append(s)
    assert sols[0].iterations == sols[1].iterations
    assert sols[0].lin_iterations == sols[0].iterations < sols[1].lin_iterations
    assert rel_err(sols[0].U, sols[1].U) <= 1e-10


def test_zero_pivot_is_a_diverged_linear_solve(nls, orc):
    """A bar without a Dirichlet boundary: K is exactly singular, the elimination meets a zero pivot, and the Newton loop
    stops with DIVERGED_LINEAR_SOLVE (newton_fem.jl:157-167: the reference catches the SingularException of `\\`)."""
    mat = nls.LinearElasticity1D(2.0)
    mesh = nls.bar_mesh(4, 1)
    fem = nls.FEMModel(mesh, 2, 1, data={"A": 1.0})
    fem.addboundary(nls.Neumann([mesh.nn - 1], [0.25]))
    s = _newton(nls, fem, mat, 1, rtol=1e-10, maxits=5)
    assert s.converged_reason is nls.REASON_DIVERGED_LINEAR_SOLVE
