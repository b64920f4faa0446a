"""Runs the 3-D line-pair assembly a few times (for ncu). Usage: python tools/lines_run.py n [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mesh = nls.grid_mesh(n, 3)
fem = nls.FEMModel(mesh)
mat = nls.NeoHookean(E=1.0, nu=0.3)
H = fem.handle(mat)
H.call("nls_dev_upload_u", nls.smooth_field(mesh.coords, 0.2 / (n - 1)).ctypes.data_as(C.POINTER(C.c_double)))
ms = C.c_float(0)
for rep in range(reps):
    H.call("nls_flush_l2"); H.call("nls_timer_start"); H.call("nls_dev_assemble", 1, 1); H.call("nls_timer_stop", C.byref(ms))
    print("n", n, "rep", rep, "ms", ms.value, flush=True)
