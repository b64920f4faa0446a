"""Times the assembly schemes (device-resident, L2 flushed between runs, CUDA events on the library stream).
Usage: python tools/asm_probe.py "dim:n:opt[:key=val,...]" ...   e.g.  2:708:1 2:708:2 3:150:2 3:150:0
Prints one line per case: fused R+K, Jacobian-only and residual-only times and the algorithmic GB/s."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
cases = sys.argv[1:] or ["2:708:1", "2:708:2", "3:70:2", "3:70:0"]
models = {}
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
for case in cases:
    f = case.split(":")
    dim, n, opt = int(f[0]), int(f[1]), int(f[2])
    extra = dict(kv.split("=") for kv in f[3].split(",")) if len(f) > 3 else {}
    if (dim, n) not in models:
        models.clear()                      # one model resident at a time
        mesh = nls.grid_mesh(n, dim)
        fem = nls.FEMModel(mesh)
        b = nls.face_nodes(n, dim, dim - 1, 0)
        d = (b[:, None] * dim + np.arange(dim)[None, :]).ravel()
        fem.addboundary(nls.Dirichlet(d, np.zeros(len(d))))
        mat = nls.NeoHookean(E=1.0, nu=0.3)
        H = fem.handle(mat)
        H.call("nls_dev_upload_u", dp(nls.smooth_field(mesh.coords, 0.2 / (n - 1))))
        models[(dim, n)] = (mesh, fem, H)
    mesh, fem, H = models[(dim, n)]
    for k, v in extra.items():
        H.call("nls_set_option", k.encode(), int(v))
    H.call("nls_set_option", b"assembly", opt)
    nrows, nnz = C.c_int64(0), C.c_int64(0)
    H.call("nls_pattern_size", C.byref(nrows), C.byref(nnz))
    abytes = 4 * mesh.nen * mesh.nel + 8 * dim * mesh.nn + 16 * nrows.value + 8 * nnz.value
    ms = C.c_float(0)
    res = []
    for (wr, wk) in [(1, 1), (0, 1), (1, 0)]:
        ts = []
        for rep in range(6):
            H.call("nls_flush_l2")
            H.call("nls_timer_start"); H.call("nls_dev_assemble", wr, wk); H.call("nls_timer_stop", C.byref(ms))
            ts.append(ms.value)
        res.append(float(np.median(ts[2:])))
    print("dim %d n %d opt %d %s ndof %d: R+K %.3f ms (%.0f GB/s alg, %.3g DoF/s, %.1f%% of 6551 GB/s) | K %.3f ms | R %.3f ms" % (
        dim, n, opt, extra, fem.ndof, res[0], abytes / res[0] / 1e6, fem.ndof / res[0] * 1e3,
        abytes / res[0] / 1e6 / 65.514, res[1], res[2]), flush=True)
