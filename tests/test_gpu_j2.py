"""Small-strain J2 plasticity on Q1 quads / hexahedra (csrc/kernels_j2.cu): the stateful-material protocol of
the reference (trial state and tangent data written by the internal-force pass, read by the tangent pass, committed on
poststep!) against its CPU restatement oracle.J2Oracle -- residual, CSR Jacobian and every per-qp state array, elastic and
plastic; a load / unload history driven by PseudoTimeStepper."""
import numpy as np
import pytest

from helpers import distorted, node_dofs, rel_err

pytestmark = pytest.mark.gpu
E, NU, SY, HH = 100.0, 0.3, 0.5, 10.0
NAMES = ["epsp_xx", "epsp_yy", "epsp_zz", "epsp_xy", "epsp_yz", "epsp_xz", "alpha"]


def _dense(K, n):
    A = np.zeros((n, n))
    for r in range(n):
        A[r, K.colind[K.rowptr[r]:K.rowptr[r + 1]]] = K.vals[K.rowptr[r]:K.rowptr[r + 1]]
    return A


def _state(nls, fem, suffix=""):
    return {nm: nls.getstate(fem, nm + suffix) for nm in NAMES}


def _oracle_state(o, committed=False):
    ep, al = (o.epsp_n, o.alpha_n) if committed else (o.epsp, o.alpha)
    return {"epsp_xx": ep[..., 0, 0], "epsp_yy": ep[..., 1, 1], "epsp_zz": ep[..., 2, 2], "epsp_xy": ep[..., 0, 1],
            "epsp_yz": ep[..., 1, 2], "epsp_xz": ep[..., 0, 2], "alpha": al}


@pytest.mark.parametrize("dim,n,amp", [(2, 5, 1e-4), (2, 6, 0.02), (3, 4, 1e-4), (3, 4, 0.02)])
def test_j2_residual_jacobian_and_state(nls, orc, dim, n, amp):
    mat = nls.J2Plasticity(E=E, nu=NU, sigma_y=SY, H=HH)
    mesh = nls.grid_mesh(n, dim)
    mesh.coords[:] = distorted(mesh, amp=0.1, seed=2)
    fem = nls.FEMModel(mesh)
    db = node_dofs(nls.face_nodes(n, dim, dim - 1, 0), dim)
    fem.addboundary(nls.Dirichlet(db, np.zeros(len(db))))
    nls.initialize_state(mat, fem)
    o = orc.J2Oracle(dim, mesh.conn, mesh.coords, mat.K, mat.G, SY, HH)
    o.set_dirichlet(db, np.zeros(len(db)))
    rng = np.random.default_rng(7)
    for rep in range(2):                 # second round: from a committed plastic state
        U = amp * rng.standard_normal(fem.ndof)
        Ug = U.copy()
        R, K = np.zeros(fem.ndof), nls.CSRMatrix(fem)
        nls.residual_jacobian(mat, fem, Ug, R, K)
        Uo, Ro = o.residual(U)
        Ko = o.jacobian_dense(U)
        assert np.array_equal(Ug, Uo)
        assert rel_err(R, Ro) <= 1e-12
        assert rel_err(_dense(K, fem.ndof), Ko) <= 1e-12
        got, want = _state(nls, fem), _oracle_state(o)
        for nm in NAMES:
            assert np.abs(got[nm] - want[nm]).max() <= 1e-13 * max(1.0, np.abs(want[nm]).max()), nm
        plastic = int((o.alpha > o.alpha_n).sum())
        assert (plastic > 0) == (amp > 1e-3)
        # the tangent pass alone reads what the internal-force pass stored (linearelastoplasticity.jl:88-93)
        K2 = nls.CSRMatrix(fem)
        nls.Jacobian(mat)(fem, U.copy(), K2)
        assert rel_err(K2.vals, K.vals) <= 1e-14
        # nothing is committed until save_state!
        gotn = _state(nls, fem, "_n")
        for nm in NAMES:
            assert np.abs(gotn[nm] - _oracle_state(o, True)[nm]).max() <= 1e-13
        nls.save_state(mat, fem, Ug)
        o.commit_state()
        gotn = _state(nls, fem, "_n")
        for nm in NAMES:
            assert np.abs(gotn[nm] - _oracle_state(o, True)[nm]).max() <= 1e-13 * max(1.0, np.abs(want[nm]).max()), nm


@pytest.mark.parametrize("dim,n", [(2, 7), (3, 4)])
def test_j2_load_unload_history(nls, orc, dim, n):
    """Displacement-driven compression into the plastic range and back (PseudoTimeStepper, pseudotime.jl:73-93): Newton
    iteration counts, displacements and the committed state of every step against the oracle path."""
    mat = nls.J2Plasticity(E=E, nu=NU, sigma_y=SY, H=HH)
    mesh = nls.grid_mesh(n, dim)
    fem = nls.FEMModel(mesh)
    db = node_dofs(nls.face_nodes(n, dim, dim - 1, 0), dim)
    dt = node_dofs(nls.face_nodes(n, dim, dim - 1, 1), dim, [dim - 1])
    # peak strain 1.5 % >> yield strain 0.65 %, then 0.5 % of elastic unloading (full unloading re-yields in reverse, where
    # plain Newton without a line search cycles -- in the oracle too; the reference's bar does the same, see
    # test_timestepper_aborts_on_nonconvergence)
    load = lambda t: -0.03 * t if t <= 0.5 else -0.015 + 0.01 * (t - 0.5)          # noqa: E731
    fem.addboundary(nls.Dirichlet(db, np.zeros(len(db))))
    fem.addboundary(nls.TimeDependent(nls.Dirichlet(dt, np.zeros(len(dt))), lambda t: np.full(len(dt), load(t))))
    nls.initialize_state(mat, fem)
    s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-13, maxits=30, lin_rtol=1e-13)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    ts = nls.PseudoTimeStepper(s, dt=0.125, maxtime=1.0)
    ts.solve()
    assert ts.step == 8
    o = orc.J2Oracle(dim, mesh.conn, mesh.coords, mat.K, mat.G, SY, HH)
    U = np.zeros(fem.ndof)
    yielded = False
    for k in range(1, 9):
        o.set_dirichlet(np.concatenate([db, dt]), np.concatenate([np.zeros(len(db)), np.full(len(dt), load(0.125 * k))]))
        U, its, hist = o.newton(U, rtol=1e-10, atol=1e-13, maxits=30)
        o.commit_state()
        yielded = yielded or o.alpha_n.max() > 0
        assert ts.num_its[k, 0] == its, (k, ts.num_its[k, 0], its)
        assert rel_err(ts.U[k], U) <= 1e-10, k
    assert yielded and o.alpha_n.max() > 0.005
    gotn = _state(nls, fem, "_n")
    for nm in NAMES:
        want = _oracle_state(o, True)[nm]
        assert np.abs(gotn[nm] - want).max() <= 1e-10 * max(1.0, np.abs(want).max()), nm
    assert np.abs(gotn["epsp_zz" if dim == 3 else "epsp_yy"]).max() > 1e-3      # plastic strain stays after the unloading
