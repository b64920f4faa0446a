"""Krylov preconditioners on the SURVEY 8d load step of a 3-D cube: iterations and time of the Newton solve per
preconditioner. Usage: python tools/mg_probe.py n [precond ...] [key=value ...]   (precond: 0 Jacobi, 1 block-Jacobi, 2 multigrid)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
nls = g.build()
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
shape = None
if "," in sys.argv[1]:      # nx,ny,nz: a box (the top displacement stays 1 % of one nx-sided block)
    shape = tuple(int(v) for v in sys.argv[1].split(","))
    n, dim = shape[0], len(shape)
else:
    n = int(sys.argv[1])
    dim = 3
    if n < 0:           # negative: 2-D, -n nodes per side
        n, dim = -n, 2
pre = [int(a) for a in sys.argv[2:] if "=" not in a] or [0, 1, 2]
kv = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
lin_rtol = float(kv.pop("lin_rtol", 1e-8))
mesh = nls.grid_mesh(shape if shape else n, dim)
fem = nls.FEMModel(mesh)
plane = int(np.prod(shape[:-1])) if shape else n ** (dim - 1)
nslow = shape[-1] if shape else n
height = (nslow - 1) / (n - 1)
d = (np.arange(plane)[:, None] * dim + np.arange(dim)[None, :]).ravel()
fem.addboundary(nls.Dirichlet(d, np.zeros(len(d))))
t_nodes = (nslow - 1) * plane + np.arange(plane)
d = t_nodes * dim + (dim - 1)
fem.addboundary(nls.Dirichlet(d, np.full(len(d), -0.01)))
mat = nls.NeoHookean(E=1.0, nu=0.3)
H = fem.handle(mat)
Uaff = np.zeros((mesh.nn, dim)); Uaff[:, dim - 1] = -0.01 * mesh.coords[:, dim - 1] / height
Uaff = Uaff.ravel()
ref = None
for pc in pre:
    H.call("nls_set_option", b"precond", pc)
    for k, v in kv.items():
        H.call("nls_set_option", k.encode(), int(v))
    H.call("nls_dev_upload_u", dp(Uaff))
    H.call("nls_solver_prepare")
    opts = nls._lib.NewtonOpts(rtol=1e-10, atol=1e-14, dtol=1e6, maxits=25, type_modified=0, lin_rtol=lin_rtol, lin_maxits=100000)
    res = nls._lib.NewtonResult()
    hist = np.zeros(64)
    H.call("nls_sync")
    t0 = time.perf_counter()
    H.call("nls_dev_newton_solve", C.byref(opts), dp(hist), C.byref(res))
    wall = time.perf_counter() - t0
    U = np.zeros(fem.ndof); H.call("nls_dev_download_u", dp(U))
    if ref is None:
        ref = U
    print("n %d precond %d %s: Newton its %d reason %d, Krylov its %d, total %.1f ms (assembly %.1f, linear %.1f), wall %.2f s, |U - U_first|/|U| %.2e, norms %s"
          % (n, pc, kv, res.iterations, res.reason, res.lin_its_total, res.ms_total, res.ms_assembly, res.ms_linear, wall,
             np.abs(U - ref).max() / np.abs(ref).max(), " ".join("%.2e" % v for v in hist[:res.nhist])), flush=True)
