"""2-rank probe of the distributed multigrid on a small slab model (debugging aid): torchrun --nproc-per-node 2 tools/mg_dist_probe.py dim nx ny [nz]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch, torch.distributed as dist
import __graft_entry__ as g
nls = g.load_package()
from helpers import node_dofs
rank, world, lr = nls.distributed.env_rank()
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dim = int(sys.argv[1]); shape = tuple(int(v) for v in sys.argv[2:2 + dim])
uid = nls.distributed.broadcast_unique_id(dist, rank, device="cuda")
mat = nls.NeoHookean(E=1.0, nu=0.3)
fem, part = nls.distributed.slab_model(nls.FEMModel, shape, dim, world, rank, lr, uid)
H = fem.handle(mat)
H.call("nls_set_option", b"precond", 2)
l2g, no = part["l2g"], part["n_owned"]
g2l = {int(gid): i for i, gid in enumerate(l2g)}
db = node_dofs(nls.face_nodes(shape, dim, dim - 1, 0), dim)
dt = node_dofs(nls.face_nodes(shape, dim, dim - 1, 1), dim, [dim - 1])
gd = np.concatenate([db, dt]); gv = np.concatenate([np.zeros(len(db)), np.full(len(dt), -0.005)])
keep = [k for k in range(len(gd)) if int(gd[k] // dim) in g2l and g2l[int(gd[k] // dim)] < no]
ld = np.array([g2l[int(gd[k] // dim)] * dim + int(gd[k] % dim) for k in keep], dtype=np.int64)
fem.addboundary(nls.Dirichlet(ld, gv[keep]))
print("rank", rank, "model built, preconditioner", nls.lib().nls_krylov_preconditioner(fem.handle(mat).h), flush=True)
s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=25, lin_rtol=1e-10)
nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
s.solve(np.zeros(fem.ndof))
print("rank", rank, "newton", s.iterations, repr(s.converged_reason), "krylov", s.lin_iterations, flush=True)
dist.barrier(); dist.destroy_process_group()
