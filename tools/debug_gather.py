import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
n = 66
mesh = nls.grid_mesh(n, 2)
rng = np.random.default_rng(5)
perm = rng.permutation(mesh.nn)
conn = perm[mesh.conn][rng.permutation(mesh.nel)].astype(np.int32)
coords = np.zeros_like(mesh.coords); coords[perm] = mesh.coords
mesh = nls.Mesh(2, conn, coords=coords)
mat = nls.NeoHookean(E=1.0, nu=0.3)
fem = nls.FEMModel(mesh)
U = nls.smooth_field(mesh.coords, amp=0.02)
R, K = np.full(fem.ndof, np.nan), nls.CSRMatrix(fem)
K.vals[:] = np.nan
nls.residual_jacobian(mat, fem, U.copy(), R, K)
bad = np.flatnonzero(np.isnan(R))
print("nan R entries", len(bad), "nan K", np.isnan(K.vals).sum())
print("bad nodes coords*(n-1):", (mesh.coords[bad[::2] // 2] * (n - 1))[:20])
H = fem.handle(mat); H.call("nls_set_option", b"assembly", 0)
R0, K0 = np.zeros(fem.ndof), nls.CSRMatrix(fem)
nls.residual_jacobian(mat, fem, U.copy(), R0, K0)
ok = ~np.isnan(R)
print("max diff on written rows", np.abs(R[ok] - R0[ok]).max(), np.nanmax(np.abs(K.vals - K0.vals)))
