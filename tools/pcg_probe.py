"""Times a fixed number of PCG iterations on the C2 (or a 3-D) Jacobian, device-resident.
Usage: python tools/pcg_probe.py [dim n iters]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 708
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
mesh = nls.grid_mesh(n, dim)
fem = nls.FEMModel(mesh)
b = nls.face_nodes(n, dim, dim - 1, 0)
d = (b[:, None] * dim + np.arange(dim)[None, :]).ravel()
fem.addboundary(nls.Dirichlet(d, np.zeros(len(d))))
mat = nls.NeoHookean(E=1.0, nu=0.3)
H = fem.handle(mat)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
U = nls.smooth_field(mesh.coords, 0.2 / (n - 1))
H.call("nls_dev_upload_u", dp(U))
ms = C.c_float(0)
t0 = time.time(); H.call("nls_dev_assemble", 1, 1); H.call("nls_sync")
for rep in range(3):
    H.call("nls_timer_start"); H.call("nls_dev_assemble", 1, 1); H.call("nls_timer_stop", C.byref(ms))
nrows, nnz = C.c_int64(0), C.c_int64(0)
H.call("nls_pattern_size", C.byref(nrows), C.byref(nnz))
abytes = 4 * mesh.nen * mesh.nel + 8 * dim * mesh.nn + 16 * nrows.value + 8 * nnz.value
print("dim %d n %d ndof %d nnz %d  assembly %.3f ms  (%.1f GB/s algorithmic, %.2e DoF/s)" % (
    dim, n, fem.ndof, nnz.value, ms.value, abytes / ms.value / 1e6, fem.ndof / ms.value * 1e3))
if dim == 2:
    H.call("nls_set_option", b"profile_phases", 1)
    for _ in range(5): H.call("nls_dev_assemble", 1, 1)
    cnt = np.zeros(8, dtype=np.int64)
    H.call("nls_debug_counters", cnt.ctypes.data_as(C.POINTER(C.c_int64)))
    ncl = max(1, cnt[4])
    print("gather phases, cycles per cluster: staging %.0f, elements %.0f, block tasks %.0f (clusters %d)" % (
        cnt[0] / ncl, cnt[1] / ncl, cnt[3] / ncl, ncl))
    print("  staging split: cp.async wait %.0f, barrier %.0f, issue_B %.0f" % (cnt[5] / ncl, cnt[6] / ncl, cnt[7] / ncl))
    H.call("nls_set_option", b"profile_phases", 0)
H.call("nls_dev_pcg_iterations", 20)
H.call("nls_timer_start"); H.call("nls_dev_pcg_iterations", iters); H.call("nls_timer_stop", C.byref(ms))
it_bytes = 12 * nnz.value + 24 * nrows.value + 128 * nrows.value
print("pcg: %d iterations %.3f ms -> %.2f us/iter, %.1f GB/s algorithmic (SpMV 12nnz+24N, vectors 128N)" % (
    iters, ms.value, ms.value / iters * 1e3, it_bytes * iters / ms.value / 1e6))
for spmv, lanes in [(0, 0), (1, 2), (1, 4), (1, 8), (1, 16)]:
    H.call("nls_set_option", b"spmv", spmv); H.call("nls_set_option", b"spmv_lanes", lanes)
    for rep in range(2):
        H.call("nls_timer_start")
        for _ in range(50): H.call("nls_dev_spmv")
        H.call("nls_timer_stop", C.byref(ms))
    print("spmv kind %d lanes %2d: %.2f us, %.1f GB/s (of 12nnz+24N)" % (spmv, lanes, ms.value / 50 * 1e3, (12 * nnz.value + 24 * nrows.value) * 50 / ms.value / 1e6))
