"""Shared problem builders: the same seeded/deterministic inputs for the CUDA path and the oracle."""
import numpy as np

RTOL = 1e-12   # residual / Jacobian entries (north_star): |a-b| <= RTOL * max|b| over the array
UTOL = 1e-10   # converged displacement, relative


def rel_err(a, b):
    """Error relative to the scale of the reference array (entries that cancel to ~0 are judged
    against the magnitude of the contributions they are made of)."""
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)


def entry_err(a, b, scale):
    """Entry-wise relative error: every entry judged against the sum of the absolute element contributions it is made
    of (oracle.contribution_scale) -- the north star's 'entries within 1e-12 relative', robust to entries that cancel."""
    s = np.where(scale > 0, scale, 1.0)
    return float((np.abs(a - b) / s).max()) if len(a) else 0.0


def node_dofs(nodes, dim, comps=None):
    comps = range(dim) if comps is None else comps
    return (np.asarray(nodes)[:, None] * dim + np.asarray(list(comps))[None, :]).ravel()


def bar_problem(nls, orc, nel, p, q, material, area=1.0, tip_force=0.5, device=0):
    """KAT-3 style bar: node 0 clamped, tip force on the last node."""
    mesh = nls.bar_mesh(nel, p)
    fem = nls.FEMModel(mesh, q, p, data={"A": area}, device=device)
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    fem.addboundary(nls.Neumann([mesh.nn - 1], [tip_force]))
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=material.kind,
                         mat=material.params(), area=area)
    om.set_dirichlet([0], [0.0])
    om.set_neumann([mesh.nn - 1], [tip_force])
    return fem, om


def grid_problem(nls, orc, n, dim, material, compress=-0.05, device=0, with_oracle=True):
    """Unit square/cube: face (slowest axis)=0 clamped, opposite face pushed by `compress`."""
    mesh = nls.grid_mesh(n, dim)
    fem = nls.FEMModel(mesh, device=device)
    bottom = nls.face_nodes(n, dim, dim - 1, 0)
    top = nls.face_nodes(n, dim, dim - 1, 1)
    db = node_dofs(bottom, dim)
    dt = node_dofs(top, dim, [dim - 1])
    fem.addboundary(nls.Dirichlet(db, np.zeros(len(db))))
    fem.addboundary(nls.Dirichlet(dt, np.full(len(dt), compress)))
    om = None
    if with_oracle:
        om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(material.mu, material.lam))
        om.set_dirichlet(np.concatenate([db, dt]), np.concatenate([np.zeros(len(db)), np.full(len(dt), compress)]))
    return fem, om


def distorted(mesh, amp=0.15, seed=0):
    """Move interior nodes by a seeded random fraction of the cell size (non-affine elements)."""
    rng = np.random.default_rng(seed)
    X = mesh.coords.copy()
    n = int(round(len(X) ** (1.0 / mesh.dim)))
    h = 1.0 / (n - 1)
    interior = np.all((X > 1e-12) & (X < 1 - 1e-12), axis=1)
    X[interior] += rng.uniform(-amp * h, amp * h, size=(interior.sum(), mesh.dim))
    return X
