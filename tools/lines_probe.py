"""Per-phase cycle counters of the 3-D line-pair assembly kernel. Usage: python tools/lines_probe.py n [n ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
for n in [int(v) for v in sys.argv[1:]] or [70]:
    mesh = nls.grid_mesh(n, 3)
    fem = nls.FEMModel(mesh)
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    H = fem.handle(mat)
    H.call("nls_dev_upload_u", dp(nls.smooth_field(mesh.coords, 0.2 / (n - 1))))
    ms = C.c_float(0)
    for rep in range(3):
        H.call("nls_dev_assemble", 1, 1)
    for mode in (1, 2):
        H.call("nls_set_option", b"profile_phases", mode)
        H.call("nls_flush_l2"); H.call("nls_timer_start"); H.call("nls_dev_assemble", 1, 1); H.call("nls_timer_stop", C.byref(ms))
        out = np.zeros(24, dtype=np.int64)
        H.call("nls_debug_counters_ext", out.ctypes.data_as(C.POINTER(C.c_int64)))
        ni = max(1, out[2])
        print("n %d mode %d (2 = no global stores): %.3f ms; intervals %d; per interval (warp 0): phase A to barrier %.0f clk, rest %.0f clk [cp.async wait %.0f, finalize %.0f, last barrier %.0f, row hand-off %.0f]; per-warp own phase-B clk: %s" % (
            n, mode, ms.value, ni, out[0] / ni, out[1] / ni, out[3] / ni, out[4] / ni, out[5] / ni, out[6] / ni, " ".join("%.0f" % (v / ni) for v in out[8:20])), flush=True)
        print("   warp 0: own phase-A work %.0f, phase A + barrier + plane/carry fetch issue %.0f (fetch issue = this - 'phase A to barrier'), first phase B to its barrier %.0f" % (out[20] / ni, out[21] / ni, out[22] / ni), flush=True)
