"""Edge cases of the hot path through the C ABI: empty and ragged inputs, isolated nodes, irregular valence,
strongly graded meshes, fully constrained models. Every case is checked against the oracle (1e-12 relative,
patterns bit-exact) for every assembly scheme the mesh admits."""
import ctypes as C

import numpy as np
import pytest

from helpers import RTOL, UTOL, rel_err

pytestmark = pytest.mark.gpu


def _check_all_schemes(nls, orc, mesh, U, dirichlet=None, neumann=None, expect_scheme=None):
    dim = mesh.dim
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem = nls.FEMModel(mesh)
    om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    if dirichlet:
        fem.addboundary(nls.Dirichlet(*dirichlet)); om.set_dirichlet(*dirichlet)
    if neumann:
        fem.addboundary(nls.Neumann(*neumann)); om.set_neumann(*neumann)
    K = nls.CSRMatrix(fem)
    rp, ci = om.pattern()
    assert np.array_equal(rp, K.rowptr) and np.array_equal(ci, K.colind)
    Uo, Ro = om.residual(U)
    Ko = om.jacobian_csr(U)
    H = fem.handle(mat)
    seen = set()
    for opt in (1, 2, 0):
        H.call("nls_set_option", b"assembly", opt)
        seen.add((opt, nls.lib().nls_assembly_scheme(H.h)))
        R, Ud = np.full(fem.ndof, np.nan), U.copy()
        K.vals[:] = np.nan
        nls.residual_jacobian(mat, fem, Ud, R, K)
        assert np.array_equal(Ud, Uo), opt
        assert rel_err(R, Ro) <= RTOL, (opt, rel_err(R, Ro))
        if len(Ko):
            assert rel_err(K.vals, Ko) <= RTOL, (opt, rel_err(K.vals, Ko))
    if expect_scheme is not None:
        assert (1, expect_scheme) in seen, seen
    return fem, om, mat


def test_no_elements(nls, orc):
    """A model with nodes but no elements: empty pattern, R = -Fext, nothing crashes."""
    mesh = nls.Mesh(2, np.zeros((0, 4), dtype=np.int32), coords=np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    fem = nls.FEMModel(mesh)
    fem.addboundary(nls.Neumann([1, 4], [0.5, -0.25]))
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    K = nls.CSRMatrix(fem)
    assert len(K.colind) == 0 and np.all(K.rowptr == 0)
    R = np.full(fem.ndof, np.nan)
    nls.Residual(mat)(fem, np.zeros(fem.ndof), R)
    expect = np.zeros(fem.ndof); expect[1] = -0.5; expect[4] = 0.25
    assert np.array_equal(R, expect)


@pytest.mark.parametrize("dim", [2, 3])
def test_single_element_and_isolated_nodes(nls, orc, dim):
    """One element plus nodes that belong to no element (empty CSR rows)."""
    nen = 1 << dim
    cube = np.array([[(a >> d) & 1 for d in range(dim)] for a in range(nen)], dtype=np.float64)
    coords = np.vstack([cube * [1.0, 0.8, 1.2][:dim], np.full((2, dim), 5.0) + np.arange(2)[:, None]])
    perm = np.array([nen + 1, *range(nen), nen])[::-1].copy()           # isolated nodes first and last, ids shuffled
    inv = np.argsort(perm)
    conn = inv[np.arange(nen)][None, :].astype(np.int32)
    mesh = nls.Mesh(dim, conn, coords=coords[perm])
    U = 0.05 * np.sin(np.arange(mesh.nn * dim) + 1.0)
    _check_all_schemes(nls, orc, mesh, U, neumann=([0, 3], [0.1, -0.2]))


def test_l_shape_with_hole_2d(nls, orc):
    """Ragged 2-D domain: an L-shape with a hole cut out of a structured grid (partially filled 15x7 tiles, tiles
    without elements, isolated nodes inside the hole)."""
    n = 40
    mesh = nls.grid_mesh(n, 2)
    cx = mesh.coords[mesh.conn].mean(axis=1)
    keep = ~((cx[:, 0] > 0.55) & (cx[:, 1] > 0.6)) & ~((np.abs(cx[:, 0] - 0.3) < 0.1) & (np.abs(cx[:, 1] - 0.3) < 0.08))
    mesh = nls.Mesh(2, mesh.conn[keep].copy(), coords=mesh.coords)
    U = nls.smooth_field(mesh.coords, amp=0.2 / (n - 1))
    _check_all_schemes(nls, orc, mesh, U, dirichlet=([0, 1, 2, 3], [0.0, 0.0, 0.001, 0.0]), neumann=([50, 51], [0.3, -0.1]),
                       expect_scheme=1)


def test_graded_mesh_2d(nls, orc):
    """Strongly graded coordinates: the planner's bins (about one node per bin on average) hold many nodes where the
    mesh is fine, so tiles overflow and are split into several clusters."""
    n = 48
    mesh = nls.grid_mesh(n, 2)
    X = mesh.coords.copy()
    X = X ** 3.0                                  # element sizes vary by ~3 n^2 across the domain
    mesh = nls.Mesh(2, mesh.conn, coords=X)
    U = 0.01 * X.ravel() * np.sin(np.arange(X.size))          # displacement scaled with the local element size
    _check_all_schemes(nls, orc, mesh, U, dirichlet=([0, 1], [0.0, 0.0]), expect_scheme=1)


@pytest.mark.parametrize("valence,scheme", [(3, 1), (5, 1), (8, 1), (9, 2)])
def test_irregular_valence_star_2d(nls, orc, valence, scheme):
    """`valence` quads around one node (unstructured topology). Up to 8 the shared-memory gather kernel's packed
    source lists hold them (4 for the fast path, 8 for the general one); beyond that the model falls back to
    the element-record scheme."""
    k = valence
    ang = 2 * np.pi * np.arange(k) / k
    a = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    b = 1.7 * np.stack([np.cos(ang + np.pi / k), np.sin(ang + np.pi / k)], axis=1)
    coords = np.vstack([[0.0, 0.0], a, b])
    conn = np.array([[0, 1 + i, 1 + (i + 1) % k, 1 + k + i] for i in range(k)], dtype=np.int32)
    mesh = nls.Mesh(2, conn, coords=coords)
    U = 0.03 * np.cos(1.3 * np.arange(mesh.nn * 2))
    _check_all_schemes(nls, orc, mesh, U, neumann=([0, 1], [0.2, 0.1]), expect_scheme=scheme)


def test_fully_constrained_newton(nls, orc):
    """Every DoF prescribed: the free block is empty, ||R|| over it is 0, norm_r0 falls back to 1 (newton_fem.jl:143)
    and the solve reports relative convergence at iteration 0 with U = the prescribed values."""
    n, dim = 5, 2
    mesh = nls.grid_mesh(n, dim)
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem = nls.FEMModel(mesh)
    dofs = np.arange(mesh.nn * dim)
    vals = 0.01 * np.sin(dofs + 0.5)
    fem.addboundary(nls.Dirichlet(dofs, vals))
    s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=5)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve()
    om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet(dofs, vals)
    ref = om.newton(rtol=1e-10, atol=1e-14, maxits=5, linear="pcg")
    assert s.iterations == ref["iterations"] == 0
    assert repr(s.converged_reason).startswith("Converged") and ref["reason"] == "CONVERGED_RELATIVE"
    assert np.array_equal(s.U, vals) and np.array_equal(ref["U"], vals)


def test_bar_high_order_many_gauss_points_ragged_lengths(nls, orc):
    """1-D: highest supported order and quadrature (p = 7, q = 5), element lengths spanning three decades."""
    p, q, nel = 7, 5, 9
    mesh = nls.bar_mesh(nel, p)
    ends = np.concatenate([[0.0], np.cumsum(np.logspace(-3, 0, nel))])
    X = np.concatenate([ends[e] + (ends[e + 1] - ends[e]) * np.linspace(0, 1, p + 1)[(0 if e == 0 else 1):] for e in range(nel)])
    mesh.coords[:] = X.reshape(mesh.coords.shape)
    mat = nls.ExponentialMaterial(1.0, 2.0)
    fem = nls.FEMModel(mesh, q, p, data={"A": 0.3})
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=0.3)
    U = 0.01 * np.sin(3 * X)
    R, K = np.zeros(fem.ndof), nls.CSRMatrix(fem)
    nls.residual_jacobian(mat, fem, U.copy(), R, K)
    assert rel_err(R, om.residual(U)[1]) <= RTOL and rel_err(K.vals, om.jacobian_csr(U)) <= RTOL
