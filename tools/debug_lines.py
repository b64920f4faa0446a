"""Scratch: where the pair-owner 3-D assembly differs from the atomic scatter. Usage: debug_lines.py nx,ny,nz ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
for arg in sys.argv[1:]:
    shape = tuple(int(v) for v in arg.split(","))
    mesh = nls.grid_mesh(shape, 3)
    rng = np.random.default_rng(3)
    mesh.coords[:] = mesh.coords + rng.uniform(-0.05, 0.05, size=mesh.coords.shape) / (shape[0] - 1)
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    fem = nls.FEMModel(mesh)
    U = nls.smooth_field(mesh.coords, amp=0.2 / (shape[0] - 1)) + 0.01 * rng.uniform(-1, 1, mesh.nn * 3) / (shape[0] - 1)
    H = fem.handle(mat)
    out = {}
    for opt in (1, 0):
        H.call("nls_set_option", b"assembly", opt)
        R, K = np.full(fem.nrows, np.nan), nls.CSRMatrix(fem)
        K.vals[:] = np.nan
        nls.residual_jacobian(mat, fem, U.copy(), R, K)
        out[opt] = (R, K.vals.copy(), K.rowptr, K.colind)
    R1, K1, rp, ci = out[1]; R0, K0 = out[0][:2]
    nx, ny, nz = shape
    xyz = lambda n: (n % nx, (n // nx) % ny, n // (nx * ny))
    eR = np.abs(R1 - R0) / np.abs(R0).max(); eK = np.abs(K1 - K0) / np.abs(K0).max()
    eR[np.isnan(eR)] = 9; eK[np.isnan(eK)] = 9
    print(arg, "scheme", nls.lib().nls_assembly_scheme(H.h), "R err", eR.max(), "K err", eK.max(), "nan R", np.isnan(R1).sum(), "nan K", np.isnan(K1).sum())
    bad = np.nonzero(eR > 1e-10)[0]
    print("  bad R nodes:", sorted({tuple(int(v) for v in xyz(b // 3)) for b in bad})[:60], len(bad))
    badk = np.nonzero(eK > 1e-10)[0]
    rows = np.searchsorted(rp, badk, side="right") - 1
    pairs = sorted({(xyz(r // 3), tuple(np.subtract(xyz(ci[k] // 3), xyz(r // 3)))) for r, k in zip(rows[:4000], badk[:4000])})
    print("  bad K rows:", sorted({tuple(int(v) for v in p[0]) for p in pairs})[:40], len(badk))
