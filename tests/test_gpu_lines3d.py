"""Pair-owner ("lines") assembly of 3-D grid-topology meshes (csrc/kernels_lines3d.cu) against the oracle: whole grids of
ragged shapes (tiles of 8 x 8 lines and x-chunks that do not divide the mesh), distorted coordinates, slab-local meshes
with ghost planes (the multi-GPU numbering, checked here on one GPU), determinism, and the fall-back to the atomic
scatter on meshes without grid topology."""
import ctypes as C

import numpy as np
import pytest

from helpers import RTOL, distorted, entry_err, node_dofs, rel_err

pytestmark = pytest.mark.gpu


def _scheme(nls, H):
    return nls.lib().nls_assembly_scheme(H.h)


def _assemble(nls, mat, fem, U):
    R, K = np.full(fem.nrows, np.nan), nls.CSRMatrix(fem)
    K.vals[:] = np.nan
    Uio = U.copy()
    nls.residual_jacobian(mat, fem, Uio, R, K)
    return Uio, R, K


@pytest.mark.parametrize("shape,kind", [((2, 2, 2), "affine"), ((3, 2, 2), "distorted"), ((5, 4, 9), "distorted"),
                                        ((9, 9, 9), "distorted"), ((17, 10, 12), "affine"), ((12, 19, 8), "distorted"),
                                        ((61, 9, 10), "distorted"), ((8, 17, 18), "affine")])
def test_lines_assembly_matches_oracle(nls, orc, shape, kind):
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mesh = nls.grid_mesh(shape, 3)
    if kind == "distorted":
        rng = np.random.default_rng(3)
        h = 1.0 / (shape[0] - 1)
        mesh.coords[:] = mesh.coords + rng.uniform(-0.15 * h, 0.15 * h, size=mesh.coords.shape)
    fem = nls.FEMModel(mesh)
    dd = np.array([0, 1, 7, mesh.nn * 3 - 1]); dv = np.array([0.0, 0.001, -0.002, 0.003])
    fem.addboundary(nls.Dirichlet(dd, dv))
    fem.addboundary(nls.Neumann([5, 9], [0.25, -0.5]))
    om = orc.OracleModel(3, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet(dd, dv); om.set_neumann([5, 9], [0.25, -0.5])
    U = nls.smooth_field(mesh.coords, amp=0.2 / (shape[0] - 1))
    H = fem.handle(mat)
    assert _scheme(nls, H) == 3, "grid topology must select the pair-owner scheme"
    Uo, Ro = om.residual(U); Ko = om.jacobian_csr(U)
    Rabs, Kabs = om.contribution_scale(U)
    Uio, R, K = _assemble(nls, mat, fem, U)
    assert np.array_equal(Uio, Uo)
    rp, ci = om.pattern()
    assert np.array_equal(K.rowptr, rp) and np.array_equal(K.colind, ci)
    assert not np.isnan(R).any() and not np.isnan(K.vals).any(), "every row is written"
    assert rel_err(R, Ro) <= RTOL and rel_err(K.vals, Ko) <= RTOL
    neu = np.zeros(len(Rabs)); neu[[5, 9]] = [0.25, 0.5]
    assert entry_err(R, Ro, Rabs + neu) <= RTOL and entry_err(K.vals, Ko, Kabs) <= RTOL
    # bit-reproducible run to run; the single-operator passes are separate instantiations (last-bit differences allowed)
    _, R2, K2 = _assemble(nls, mat, fem, U)
    assert np.array_equal(R, R2) and np.array_equal(K.vals, K2.vals)
    R1, K1 = np.full(fem.ndof, np.nan), nls.CSRMatrix(fem)
    K1.vals[:] = np.nan
    nls.Residual(mat)(fem, U.copy(), R1); nls.Jacobian(mat)(fem, U.copy(), K1)
    assert rel_err(R1, R) <= 1e-14 and rel_err(K1.vals, K.vals) <= 1e-14
    # against the atomic scatter
    H.call("nls_set_option", b"assembly", 0)
    assert _scheme(nls, H) == 0
    _, R0, K0 = _assemble(nls, mat, fem, U)
    assert rel_err(R0, R) <= 1e-13 and rel_err(K0.vals, K.vals) <= 1e-13


@pytest.mark.parametrize("shape,nranks", [((5, 4, 9), 2), ((9, 11, 20), 3), ((12, 9, 33), 4)])
def test_lines_assembly_on_slab_local_meshes(nls, orc, shape, nranks):
    """The local meshes of a slab partition (owned planes first, then the ghost planes: nls_partition's numbering) on one
    GPU, ghost values supplied by hand: owned rows of R and K equal the rows of the single-domain oracle."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    gm = nls.grid_mesh(shape, 3)
    om = orc.OracleModel(3, gm.conn, gm.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    db = node_dofs(nls.face_nodes(shape, 3, 2, 0), 3)
    om.set_dirichlet(db, np.zeros(len(db)))
    Ug = nls.smooth_field(gm.coords, 0.02)
    Uo, Ro = om.residual(Ug); Ko = om.jacobian_csr(Ug); rpo, cio = om.pattern()
    sR, sK = np.abs(Ro).max(), np.abs(Ko).max()
    for rank in range(nranks):
        sl = nls.slab_local_mesh(shape, 3, nranks, rank)
        fem = nls.FEMModel(sl["mesh"])
        fem.set_partition(sl["n_owned"], sl["neigh"], sl["send_ptr"], sl["send_nodes"], sl["recv_ptr"], sl["recv_nodes"])
        l2g, no = sl["l2g"], sl["n_owned"]
        g2l = {int(gid): i for i, gid in enumerate(l2g)}
        keep = [k for k in range(len(db)) if int(db[k] // 3) in g2l and g2l[int(db[k] // 3)] < no]
        ld = np.array([g2l[int(db[k] // 3)] * 3 + int(db[k] % 3) for k in keep], dtype=np.int64)
        fem.addboundary(nls.Dirichlet(ld, np.zeros(len(ld))))
        H = fem.handle(mat)
        assert _scheme(nls, H) == 3
        Ul = Uo.reshape(-1, 3)[l2g].ravel().copy()
        _, R, K = _assemble(nls, mat, fem, Ul)
        lo = sl["node_off"][rank] * 3
        assert np.abs(R - Ro[lo:lo + no * 3]).max() <= 1e-12 * sR
        for lrow in range(no * 3):
            cols = K.colind[K.rowptr[lrow]:K.rowptr[lrow + 1]]
            gcols = l2g[cols // 3] * 3 + cols % 3
            order = np.argsort(gcols)
            assert np.array_equal(gcols[order], cio[rpo[lo + lrow]:rpo[lo + lrow + 1]])
            assert np.abs(K.vals[K.rowptr[lrow]:K.rowptr[lrow + 1]][order] - Ko[rpo[lo + lrow]:rpo[lo + lrow + 1]]).max() <= 1e-12 * sK


def test_lines_scheme_needs_grid_topology(nls, orc):
    """A renumbered mesh has no grid topology: the plan is refused and the atomic scatter stays in use."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mesh = nls.grid_mesh(6, 3)
    rng = np.random.default_rng(5)
    perm = rng.permutation(mesh.nn)
    conn = perm[mesh.conn][rng.permutation(mesh.nel)].astype(np.int32)
    coords = np.zeros_like(mesh.coords); coords[perm] = mesh.coords
    fem = nls.FEMModel(nls.Mesh(3, conn, coords=coords))
    H = fem.handle(mat)
    assert _scheme(nls, H) == 0


def test_lines_newton_matches_oracle(nls, orc):
    """Newton with the pair-owner Jacobian: same iteration count and displacement as the oracle's dense-LU Newton."""
    from helpers import UTOL, grid_problem
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, om = grid_problem(nls, orc, (7, 6, 5), 3, mat, compress=-0.04)
    s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-12)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve(np.zeros(fem.ndof))
    assert _scheme(nls, fem.handle(mat)) == 3
    ref = om.newton(rtol=1e-10, atol=1e-14, maxits=25)
    assert s.iterations == ref["iterations"]
    assert rel_err(s.U, ref["U"]) <= UTOL
