"""Full-size (BASELINE.json C2 / 3-D) checks through size-independent properties, golden fixtures,
and the 2-rank NCCL path. Needs a B200."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import RTOL, UTOL, grid_problem, node_dofs, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def test_golden_fixtures(nls):
    """Committed oracle vectors (tests/golden/oracle_cases.json) reproduced by the CUDA path."""
    cases = json.load(open(os.path.join(GOLD, "oracle_cases.json")))
    for name, c in cases.items():
        if c["dim"] == 1:
            mesh = nls.bar_mesh(c["nel"], c["p"])
            fem = nls.FEMModel(mesh, c["q"], c["p"], data={"A": c["area"]})
            fem.addboundary(nls.Dirichlet([0], [0.0])); fem.addboundary(nls.Neumann([mesh.nn - 1], [0.5]))
            mat = nls.ExponentialMaterial(1.0, 2.0)
        else:
            conn = np.array(c["conn"], dtype=np.int32).reshape(-1, 2 ** c["dim"])
            fem = nls.FEMModel(nls.Mesh(c["dim"], conn, coords=np.array(c["coords"])))
            fem.addboundary(nls.Dirichlet(c["dir_dofs"], c["dir_vals"]))
            mat = nls.NeoHookean(c["mu"], c["lam"])
        K = nls.CSRMatrix(fem)
        assert K.rowptr.tolist() == c["rowptr"] and K.colind.tolist() == c["colind"], name   # bit-exact
        U, R = np.array(c["U"]), np.zeros(fem.ndof)
        nls.residual_jacobian(mat, fem, U, R, K)
        assert U.tolist() == c["U_out"], name
        assert rel_err(R, np.array(c["R"])) <= RTOL and rel_err(K.vals, np.array(c["vals"])) <= RTOL, name


@pytest.mark.parametrize("dim,n", [(2, 708), (3, 70)])
def test_fullsize_properties(nls, orc, dim, n):
    """C2 (1 002 528 DoF) and the 1.03 M-DoF cube: patch test, translation invariance, symmetry,
    and interior rows against the oracle evaluated on a small mesh with the same spacing."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    mesh = nls.grid_mesh(n, dim)
    fem = nls.FEMModel(mesh)
    H = fem.handle(mat)
    Hm = np.array([[0.03, 0.01, 0.0], [-0.02, 0.02, 0.01], [0.0, 0.015, -0.025]])[:dim, :dim]
    U = (mesh.coords @ Hm.T).ravel()
    R, K = np.zeros(fem.ndof), nls.CSRMatrix(fem)
    nls.residual_jacobian(mat, fem, U, R, K)
    assert K.rowptr[-1] == (4 * (3 * n - 2) ** 2 if dim == 2 else 9 * (3 * n - 2) ** 3)
    interior = np.all((mesh.coords > 1e-9) & (mesh.coords < 1 - 1e-9), axis=1)
    scale = np.abs(R).max()
    assert np.abs(R.reshape(-1, dim)[interior]).max() <= 1e-11 * scale          # patch test
    F = np.eye(dim) + Hm
    Fi = np.linalg.inv(F)
    P = mat.mu * (F - Fi.T) + mat.lam * np.log(np.linalg.det(F)) * Fi.T
    top = np.isclose(mesh.coords[:, dim - 1], 1.0)
    nvec = np.zeros(dim); nvec[dim - 1] = 1.0
    assert np.allclose(R.reshape(-1, dim)[top].sum(axis=0), P @ nvec, rtol=1e-9)  # traction resultant
    # translation invariance: K t = 0 ; symmetry: x.(K y) = y.(K x)
    rng = np.random.default_rng(0)
    y = np.zeros(fem.ndof)
    for c in range(dim):
        t = np.zeros((mesh.nn, dim)); t[:, c] = 1.0
        H.call("nls_spmv", dp(t.ravel()), dp(y))
        assert np.abs(y).max() <= 1e-11 * np.abs(K.vals).max()
    a, b = rng.standard_normal(fem.ndof), rng.standard_normal(fem.ndof)
    Ka, Kb = np.zeros(fem.ndof), np.zeros(fem.ndof)
    H.call("nls_spmv", dp(a), dp(Ka)); H.call("nls_spmv", dp(b), dp(Kb))
    assert abs(b @ Ka - a @ Kb) <= 1e-10 * abs(b @ Ka)
    # uniform F on a uniform mesh: every interior block row equals the oracle's interior row on a
    # 5^dim mesh with the same spacing
    h = 1.0 / (n - 1)
    small = nls.grid_mesh(5, dim)
    Xs = small.coords * (4 * h)
    om = orc.OracleModel(dim, small.conn, Xs, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    Ks = om.jacobian_csr((Xs @ Hm.T).ravel())
    rps, _ = om.pattern()
    mid_s = sum(2 * 5 ** d for d in range(dim))                                 # node (2,2[,2])
    want = np.concatenate([Ks[rps[mid_s * dim + i]:rps[mid_s * dim + i + 1]] for i in range(dim)])
    nodes = np.flatnonzero(interior)[:: max(1, interior.sum() // 997)]
    for a_ in nodes:
        got = K.vals[K.rowptr[a_ * dim]:K.rowptr[a_ * dim + dim]]
        assert rel_err(got, want) <= 1e-11, a_


def test_fullsize_newton_c2_slice(nls, orc):
    """Newton on a 708 x 24 strip of the C2 mesh (34k DoF): same iteration count and displacement as
    the oracle's CSR+PCG variant."""
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    shape = (708, 24)
    mesh = nls.grid_mesh(shape, 2)
    fem = nls.FEMModel(mesh)
    db = node_dofs(nls.face_nodes(shape, 2, 1, 0), 2)
    dt = node_dofs(nls.face_nodes(shape, 2, 1, 1), 2, [1])
    c = -0.3 / 707
    fem.addboundary(nls.Dirichlet(db, np.zeros(len(db)))); fem.addboundary(nls.Dirichlet(dt, np.full(len(dt), c)))
    om = orc.OracleModel(2, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam), nthreads=8)
    om.set_dirichlet(np.concatenate([db, dt]), np.concatenate([np.zeros(len(db)), np.full(len(dt), c)]))
    s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-12)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve()
    ref = om.newton(rtol=1e-10, atol=1e-14, maxits=25, linear="pcg", lin_rtol=1e-12)
    assert repr(s.converged_reason).startswith("Converged") and s.iterations == ref["iterations"]
    assert rel_err(s.U, ref["U"]) <= UTOL


_NCCL_WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import __graft_entry__ as g
nls = g.load_package()
from oracle import oracle as orc
from helpers import node_dofs, rel_err
rank, world, lr = nls.distributed.env_rank()
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
# third case: interior + boundary clusters (overlapped halo); fourth: the multigrid preconditioner forced on (rank-local V-cycles)
for dim, shape, comp, precond in [(2, (9, 14), -0.02, -1), (3, (5, 4, 9), -0.05, -1), (2, (40, 46), -0.004, -1), (3, (11, 9, 23), -0.02, 2), (2, (27, 44), -0.01, 2)]:
    uid = nls.distributed.broadcast_unique_id(dist, rank, device="cuda")   # one NCCL id per communicator
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem, part = nls.distributed.slab_model(nls.FEMModel, shape, dim, world, rank, lr, uid)
    fem.handle(mat).call("nls_set_option", b"precond", precond)
    gm = nls.grid_mesh(shape, dim)
    l2g, no = part["l2g"], part["n_owned"]
    g2l = {int(gid): i for i, gid in enumerate(l2g)}
    db = node_dofs(nls.face_nodes(shape, dim, dim - 1, 0), dim)
    dt = node_dofs(nls.face_nodes(shape, dim, dim - 1, 1), dim, [dim - 1])
    gd = np.concatenate([db, dt]); gv = np.concatenate([np.zeros(len(db)), np.full(len(dt), comp)])
    keep = [k for k in range(len(gd)) if int(gd[k] // dim) in g2l and g2l[int(gd[k] // dim)] < no]
    ld = np.array([g2l[int(gd[k] // dim)] * dim + int(gd[k] % dim) for k in keep], dtype=np.int64)
    fem.addboundary(nls.Dirichlet(ld, gv[keep]))
    om = orc.OracleModel(dim, gm.conn, gm.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam))
    om.set_dirichlet(gd, gv)
    Ug = nls.smooth_field(gm.coords, 0.02)
    Ul = Ug.reshape(-1, dim)[l2g].copy()
    Ul[no:] = 123.0                                   # ghosts must come from the halo exchange
    Ul = Ul.ravel()
    R = np.zeros(no * dim); K = nls.CSRMatrix(fem)
    nls.residual_jacobian(mat, fem, Ul, R, K)
    Uo, Ro = om.residual(Ug); Ko = om.jacobian_csr(Ug); rpo, cio = om.pattern()
    assert np.array_equal(Ul.reshape(-1, dim), Uo.reshape(-1, dim)[l2g]), "halo/dirichlet"
    sR, sK = np.abs(Ro).max(), np.abs(Ko).max()
    lo = part["node_off"][rank] * dim
    assert np.abs(R - Ro[lo:lo + no * dim]).max() <= 1e-12 * sR
    for lrow in range(no * dim):                       # owned rows == global rows (columns mapped)
        cols = K.colind[K.rowptr[lrow]:K.rowptr[lrow + 1]]
        gcols = l2g[cols // dim] * dim + cols % dim
        order = np.argsort(gcols)
        assert np.array_equal(gcols[order], cio[rpo[lo + lrow]:rpo[lo + lrow + 1]])
        assert np.abs(K.vals[K.rowptr[lrow]:K.rowptr[lrow + 1]][order] - Ko[rpo[lo + lrow]:rpo[lo + lrow + 1]]).max() <= 1e-12 * sK
    s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-12)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve(np.zeros(fem.ndof))
    ref = om.newton(rtol=1e-10, atol=1e-14, maxits=25, linear="pcg", lin_rtol=1e-12)
    assert repr(s.converged_reason).startswith("Converged"), s.converged_reason
    assert s.iterations == ref["iterations"], (s.iterations, ref["iterations"])
    assert rel_err(s.U, ref["U"].reshape(-1, dim)[l2g].ravel()) <= 1e-10
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_nccl_parity(tmp_path):
    """2 ranks / 2 GPUs: slab partition, halo exchange + Krylov reductions (peer-memory windows, or NCCL with
    NLS_B200_P2P=0), the 2-D assembly with the exchange overlapped with its interior clusters; owned rows of R and
    the CSR Jacobian and the Newton solution equal the single-domain oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2


def test_golden_dynamics_fixtures(nls):
    """Committed dynamics / continuation vectors (tests/golden/oracle_dynamics_cases.json: consistent mass, Newmark
    histories, an arc-length path through a limit point) reproduced by the CUDA path."""
    dyn = json.load(open(os.path.join(GOLD, "oracle_dynamics_cases.json")))
    cases = json.load(open(os.path.join(GOLD, "oracle_cases.json")))
    # consistent mass
    g = dyn["mass_q1_2d_n4_distorted"]
    c = cases[g["case"]]
    fem = nls.FEMModel(nls.Mesh(2, np.array(c["conn"], dtype=np.int32).reshape(-1, 4), coords=np.array(c["coords"])))
    mat = nls.NeoHookean(c["mu"], c["lam"]); mat.rho = g["rho"]
    M = nls.CSRMatrix(fem)
    nls.assemblemass(mat, fem, M)
    assert rel_err(M.vals, np.array(g["vals"])) <= RTOL
    # Newmark
    g = dyn["newmark_bar_exp_p2_q3"]
    mat = nls.ExponentialMaterial(1.0, 2.0, rho=g["rho"])
    mesh = nls.bar_mesh(g["nel"], g["p"])
    fem = nls.FEMModel(mesh, g["q"], g["p"], data={"A": g["area"]})
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    fem.addboundary(nls.TimeDependent(nls.Neumann([mesh.nn - 1], [0.0]), lambda t: [g["tip_rate"] * t]))
    s = nls.NewtonSolver(fem, rtol=1e-11, atol=1e-13, maxits=20, lin_rtol=1e-13)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    ts = nls.NewmarkTimeStepper(s, mat, dt=g["dt"], maxtime=g["maxtime"])
    nls.set_ctx(s, ts)
    ts.solve()
    assert ts.num_its[:ts.step + 1, 0].astype(int).tolist() == g["num_its"]
    for mine, theirs in [(ts.U, g["U"]), (ts.Ud, g["Ud"]), (ts.Udd, g["Udd"])]:
        assert rel_err(mine[:ts.step + 1], np.array(theirs)) <= 1e-10
    # arc-length continuation
    g = dyn["arclength_arch_25x3"]
    nx, ny = g["nx"], g["ny"]
    m = nls.grid_mesh((nx, ny), 2)
    s_, yy = m.coords[:, 0], m.coords[:, 1] * (nx - 1) / (ny - 1)
    th, r = (2 * s_ - 1) * g["th0"], g["R"] + (yy - 0.5) * g["t"]
    X = np.stack([r * np.sin(th), r * np.cos(th) - g["R"] * np.cos(g["th0"])], axis=1)
    fem = nls.FEMModel(nls.Mesh(2, m.conn, coords=X))
    ends = np.concatenate([np.arange(ny) * nx, np.arange(ny) * nx + nx - 1])
    dofs = (ends[:, None] * 2 + np.arange(2)[None, :]).ravel()
    fem.addboundary(nls.Dirichlet(dofs, np.zeros(len(dofs))))
    apex = (ny - 1) * nx + nx // 2
    F = np.zeros(fem.ndof); F[2 * apex + 1] = g["load"]
    res = nls.newtonarclength(fem, nls.NeoHookean(E=g["E"], nu=g["nu"]), F, b=g["b"], da=g["da"], rtol=g["rtol"], ftol=g["ftol"],
                              maxsteps=g["maxsteps"], maxits=g["maxits"], lin_rtol=1e-13)
    assert res.num_its.tolist() == g["num_its"]
    assert rel_err(res.lam, np.array(g["lam"])) <= 1e-8 and rel_err(res.d[:, 2 * apex + 1], np.array(g["apex_v"])) <= 1e-8
