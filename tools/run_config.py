#!/usr/bin/env python
"""Runs one BASELINE.json configuration other than the bench's C2 and prints one JSON line (rank 0).

    python tools/run_config.py --config C1                       # 1-D KAT-3 bar, 100 p=3 elements, Newton to convergence
    python tools/run_config.py --config C3 [--lin-rtol 1e-8]     # 3-D n=150 (10.1 M DoF): assembly, PCG iteration, full Newton
    torchrun --nproc-per-node 8 ... tools/run_config.py --config C4 --steps 3   # 3-D n=216 (30.2 M DoF), load stepping
    python tools/run_config.py --dim 3 --side 70 --mode assembly    # one point of the C5 assembly sweep (also --strong under torchrun)

Multi-GPU runs partition the structured mesh in slabs along the slowest axis; timings are CUDA events on
the library stream, max over ranks. --weak stacks one n^dim block per rank instead of splitting one."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None, choices=[None, "C1", "C3", "C4"])
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--side", dest="n", type=int, default=70, help="nodes per side")
    ap.add_argument("--mode", default="newton", choices=["assembly", "newton", "steps"])
    ap.add_argument("--weak", action="store_true")
    ap.add_argument("--steps", type=int, default=2, help="load steps (mode steps)")
    ap.add_argument("--dstep", type=float, default=0.0, help="compression per load step (default: 0.4 element heights)")
    ap.add_argument("--lin-rtol", dest="lin_rtol", type=float, default=1e-8)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--maxits", type=int, default=25)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--assembly", type=int, default=None)
    a = ap.parse_args()
    # prescribed compression -0.1 t of the top face, PseudoTimeStepper steps of dt = 0.1 (1 % of the height per
    # step), Newton rtol 1e-10. C3: the first step alone; C4: all ten (t = 0.1 ... 1.0, 10 % compression at the end).
    if a.config == "C3":
        a.dim, a.n, a.mode = 3, 150, "newton"
        a.rtol = min(a.rtol, 1e-10)
        if a.dstep <= 0: a.dstep = 0.01
    if a.config == "C4":
        a.dim, a.n, a.mode = 3, 216, "steps"
        a.rtol = min(a.rtol, 1e-10)
        if a.dstep <= 0: a.dstep = 0.01
        if a.steps == 2: a.steps = 10

    real = os.dup(1); os.dup2(2, 1)         # NCCL prints to stdout
    import __graft_entry__ as g
    nls = g.build()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.config == "C1":
        out = run_c1(nls)
        os.write(real, (json.dumps(out) + "\n").encode())
        return

    dist, uid = None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert nls.lib().nls_comm_unique_id(nls._lib.nccl_lib_path().encode(), buf) == 0
            idt.copy_(torch.tensor(list(bytes(buf)), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        uid = bytes(idt.cpu().numpy().tolist())

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dim, n = a.dim, a.n
    nslow = n * world if a.weak else n
    shape = (n,) * (dim - 1) + (nslow,)
    t_setup = time.perf_counter()
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    plane = n ** (dim - 1)
    if world == 1:
        mesh = nls.grid_mesh(shape, dim)
        fem = nls.FEMModel(mesh)
        n_owned, p0, p1 = mesh.nn, 0, nslow
    else:
        sl = nls.slab_local_mesh(shape, dim, world, rank)
        mesh = sl["mesh"]
        fem = nls.FEMModel(mesh, device=local_rank)
        fem.set_partition(sl["n_owned"], sl["neigh"], sl["send_ptr"], sl["send_nodes"], sl["recv_ptr"], sl["recv_nodes"])
        fem.set_comm(rank, world, uid)
        n_owned = sl["n_owned"]
        cuts = nls.slab_planes(nslow, world)
        p0, p1 = cuts[rank], cuts[rank + 1]
    height = (nslow - 1) / (n - 1)
    top_bc = None
    if p0 == 0:                                   # face (slowest axis) = 0 clamped
        b = np.arange(plane)
        d = (b[:, None] * dim + np.arange(dim)[None, :]).ravel()
        fem.addboundary(nls.Dirichlet(d, np.zeros(len(d))))
    if p1 == nslow:                               # opposite face: prescribed compression, time dependent
        t_nodes = (nslow - 1 - p0) * plane + np.arange(plane)
        d = t_nodes * dim + (dim - 1)
        top_bc = nls.TimeDependent(nls.Dirichlet(d, np.zeros(len(d))), lambda t, m=len(d): np.full(m, -t))
        fem.addboundary(top_bc)
    H = fem.handle(mat)
    if a.assembly is not None:
        H.call("nls_set_option", b"assembly", a.assembly)
    L = nls.lib()
    nrows, nnz = C.c_int64(0), C.c_int64(0)
    H.call("nls_pattern_size", C.byref(nrows), C.byref(nnz))
    nrows, nnz = nrows.value, nnz.value
    t_setup = time.perf_counter() - t_setup
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    ms = C.c_float(0)
    total_dof = int(np.prod(shape)) * dim
    h_el = 1.0 / (n - 1)
    out = {"config": a.config or "%d-D n=%d" % (dim, n), "dim": dim, "nodes_per_side": n, "n_gpus": world,
           "partition": "weak (one block per rank)" if a.weak else ("slabs of one mesh" if world > 1 else "single GPU"),
           "total_dof": total_dof, "per_gpu_dof": nrows, "per_gpu_nnz": nnz, "setup_s": t_setup,
           "assembly_scheme": int(L.nls_assembly_scheme(H.h))}

    # ---- assembly (fused R+K), smooth field ----
    U0 = nls.smooth_field(mesh.coords, 0.2 * h_el)
    H.call("nls_dev_upload_u", dp(U0))
    ts = []
    for rep in range(a.reps + 2):
        H.call("nls_flush_l2"); H.call("nls_dev_barrier")
        H.call("nls_timer_start"); H.call("nls_dev_assemble", 1, 1); H.call("nls_timer_stop", C.byref(ms))
        ts.append(ms.value)
    asm_ms = allmax(float(np.median(ts[2:])))
    abytes = 4 * mesh.nen * mesh.nel + 8 * dim * mesh.nn + 16 * nrows + 8 * nnz
    hbm = 6551.4
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out["assembly"] = {"ms": asm_ms, "dof_per_s": total_dof / asm_ms * 1e3, "algorithmic_GB_per_gpu": abytes / 1e9,
                       "frac_of_hbm_per_gpu": abytes / asm_ms / 1e6 / hbm}
    if a.mode != "assembly":
        # ---- PCG iteration time on that Jacobian ----
        iters = 50
        H.call("nls_dev_pcg_iterations", 5)
        H.call("nls_timer_start"); H.call("nls_dev_pcg_iterations", iters); H.call("nls_timer_stop", C.byref(ms))
        it_ms = allmax(ms.value) / iters
        it_bytes = 12 * nnz + 24 * nrows + 128 * nrows
        out["pcg_iteration"] = {"ms": it_ms, "algorithmic_GB_per_gpu": it_bytes / 1e9, "frac_of_hbm_per_gpu": it_bytes / it_ms / 1e6 / hbm}
        # ---- Newton: load steps of `dstep` compression each; the initial guess of a step is the previous solution plus
        #      the affine field of the increment (Dirichlet by overwrite alone would invert the top element layer) ----
        dstep = a.dstep if a.dstep > 0 else 0.4 * h_el
        nsteps = a.steps if a.mode == "steps" else 1
        opts = nls._lib.NewtonOpts(0, 1e-14, a.rtol, 1e6, a.maxits, a.lin_rtol, 200000)
        U = np.zeros(fem.ndof)
        H.call("nls_solver_prepare")
        out["preconditioner"] = int(L.nls_krylov_preconditioner(H.h))
        steps = []
        for s in range(1, nsteps + 1):
            fem.updateboundaries(s * dstep)
            H = fem.handle(mat)
            Ug = U.reshape(-1, dim).copy()
            Ug[:, dim - 1] += -dstep * mesh.coords[:, dim - 1] / height
            H.call("nls_dev_upload_u", dp(np.ascontiguousarray(Ug.ravel())))
            res = nls._lib.NewtonResult()
            hist = np.zeros(a.maxits + 2)
            H.call("nls_dev_newton_solve", C.byref(opts), dp(hist), C.byref(res))
            H.call("nls_dev_halo_exchange_u")
            H.call("nls_dev_download_u", dp(U))
            steps.append({"t": s * dstep, "reason": int(res.reason), "newton_its": int(res.iterations), "pcg_its": int(res.lin_its_total),
                          "ms_total": allmax(res.ms_total), "ms_assembly": allmax(res.ms_assembly), "ms_linear": allmax(res.ms_linear),
                          "norm_r0": res.norm_r0, "norm_r": res.norm_r})
        out["newton"] = {"rtol": a.rtol, "lin_rtol": a.lin_rtol, "compression_per_step": dstep, "steps": steps}
        tot_pcg = sum(x["pcg_its"] for x in steps); tot_ms = sum(x["ms_total"] for x in steps)
        tot_newton = sum(x["newton_its"] for x in steps)
        alg = tot_pcg * it_bytes + (2 * tot_newton + nsteps) * abytes
        out["newton"]["total_ms"] = tot_ms
        out["newton"]["frac_of_hbm_per_gpu"] = alg / tot_ms / 1e6 / hbm if tot_ms > 0 else None
        out["newton"]["max_abs_u"] = float(np.abs(U).max())
    if rank == 0:
        os.write(real, (json.dumps(out) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_c1(nls):
    """C1: KAT-3 bar, 100 p=3 elements (301 DoF), Exponential material, tip force 0.5, Newton to convergence,
    checked against the closed form u(x) = x ln2 / 2."""
    import ctypes as C
    mat = nls.ExponentialMaterial(1.0, 2.0)
    mesh = nls.bar_mesh(100, 3)
    fem = nls.FEMModel(mesh, 4, 3, data={"A": 1.0})
    fem.addboundary(nls.Dirichlet([0], [0.0]))
    fem.addboundary(nls.Neumann([mesh.nn - 1], [0.5]))
    s = nls.NewtonSolver(fem, rtol=1e-10, atol=1e-14, maxits=25)
    nls.set_residual_fn(s, nls.Residual(mat)); nls.set_jacobian_fn(s, nls.Jacobian(mat))
    s.solve()                       # warm-up (module load, first launches)
    t0 = time.perf_counter(); s.solve(); wall = time.perf_counter() - t0
    return {"config": "C1", "ndof": fem.ndof, "newton_its": s.iterations, "reason": repr(s.converged_reason),
            "tip": float(s.U[-1]), "tip_exact": float(np.log(2.0) / 2), "history": [float(v) for v in s.norm_history],
            "ms_device_total": s.timing_ms.get("total"), "ms_wall_host_call": wall * 1e3, "pcg_its": s.lin_iterations}


if __name__ == "__main__":
    main()
