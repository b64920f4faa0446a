#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Newton hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): assembled DoF/s of one fused residual+Jacobian evaluation, plus the
Newton-step time, on the C2 workload (2-D plane-strain Neo-Hookean, structured Q1 mesh,
708 nodes per side = 1 002 528 DoF per GPU; weak scaling stacks one such slab per rank along y).

A "step" is one residual+Jacobian assembly pass over the whole (local) mesh:
  value : device-resident (U, R, CSR values stay in HBM), CUDA-event timed on the library's stream,
          L2 flushed between iterations, max over ranks.
  e2e   : the same pass through the reference-facing call nls_residual_jacobian with HOST buffers
          (pinned), H2D of U and D2H of U, R and the CSR values inside the timed region.
  newton_step : one Newton step (assembly + Jacobi-PCG solve + update + residual norm), device-resident.
  cpu_baseline: the oracle's C restatement of the reference algorithm on the box's host cores.
--impl reference times that CPU restatement alone (Julia is not installed, so the reference
itself cannot run).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE = 708          # C2: 708^2 nodes = 1 002 528 DoF
NCU_TRAFFIC_BYTES = 176.1e6   # measured DRAM traffic of the assembly kernel at C2 (see roofline.traffic_source)
DIM = 2
E_MOD, NU = 1.0, 0.3


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(dim, nen, nel, nn, ndof_rows, nnz):
    """conn + coords + U read + R write + vals write (colind / slot maps not counted)."""
    return 4 * nen * nel + 8 * dim * nn + 8 * ndof_rows + 8 * ndof_rows + 8 * nnz


def build_local_model(nls, nranks, rank, device, uid=None):
    shape = (N_SIDE, N_SIDE * nranks)
    mat = nls.NeoHookean(E=E_MOD, nu=NU)
    if nranks == 1:
        mesh = nls.grid_mesh(shape, DIM)
        fem = nls.FEMModel(mesh, device=device)
        n_owned, p0 = mesh.nn, 0
    else:
        sl = nls.slab_local_mesh(shape, DIM, nranks, rank)
        mesh = sl["mesh"]
        fem = nls.FEMModel(mesh, device=device)
        fem.set_partition(sl["n_owned"], sl["neigh"], sl["send_ptr"], sl["send_nodes"], sl["recv_ptr"], sl["recv_nodes"])
        fem.set_comm(rank, nranks, uid)
        n_owned, p0 = sl["n_owned"], sl["node_off"][rank]
    plane = N_SIDE
    # face y=0 clamped, opposite face pushed by -0.1*t (t = 0.1 -> -0.01 ... here one step, t = 0.1)
    if rank == 0:
        b = np.arange(plane)
        d = (b[:, None] * DIM + np.arange(DIM)[None, :]).ravel()
        fem.addboundary(nls.Dirichlet(d, np.zeros(len(d))))
    if rank == nranks - 1:
        ytop = N_SIDE * nranks - 1
        t_nodes = ytop * plane + np.arange(plane) - p0
        d = t_nodes * DIM + (DIM - 1)
        fem.addboundary(nls.Dirichlet(d, np.full(len(d), -0.1 * 0.1 * (N_SIDE * nranks - 1) / (N_SIDE - 1))))
    return fem, mat, mesh, n_owned


def _quiet_stdout():
    """Route fd 1 to stderr while the benchmark runs (NCCL prints its version banner to stdout) and
    return a function that writes the single JSON result line to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real, (json.dumps(obj) + "\n").encode())
    return emit


def run_ours(args):
    emit = _quiet_stdout()
    import __graft_entry__ as g
    nls = g.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    uid = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            rc = nls.lib().nls_comm_unique_id(nls._lib.nccl_lib_path().encode(), buf)
            assert rc == 0, nls.lib().nls_last_error(None)
            idt.copy_(torch.tensor(list(bytes(buf)), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        uid = bytes(idt.cpu().numpy().tolist())

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fem, mat, mesh, n_owned = build_local_model(nls, world, rank, local_rank, uid)
    H = fem.handle(mat)
    L = nls.lib()
    if args.assembly is not None:
        H.call("nls_set_option", b"assembly", args.assembly)
    if args.gather_ctas:
        H.call("nls_set_option", b"gather_ctas", args.gather_ctas)
    scheme = L.nls_assembly_scheme(H.h)
    comm_path = L.nls_comm_path(H.h)
    comm_us = None
    if world > 1:
        uh, ua = C.c_float(0), C.c_float(0)
        H.call("nls_dev_comm_probe", 200, C.byref(uh), C.byref(ua))
        comm_us = {"halo_exchange_us": uh.value, "allreduce_us": ua.value}
    sys.stderr.write("rank %d: assembly scheme %d (0 atomic, 1 gather, 2 element records), comm path %d (1 NCCL, 2 peer memory) %s\n"
                     % (rank, scheme, comm_path, comm_us))
    nrows, nnz = C.c_int64(0), C.c_int64(0)
    H.call("nls_pattern_size", C.byref(nrows), C.byref(nnz))
    nrows, nnz = nrows.value, nnz.value
    U0 = nls.smooth_field(mesh.coords)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    H.call("nls_dev_upload_u", dp(U0))
    ms = C.c_float(0)

    # ---- value: device-resident fused R+K assembly ----
    def step_dev():
        H.call("nls_flush_l2")
        H.call("nls_dev_barrier")        # N > 1: line the ranks' streams up, so the halo wait inside the step is not host jitter
        H.call("nls_timer_start")
        H.call("nls_dev_assemble", 1, 1)
        H.call("nls_timer_stop", C.byref(ms))
        return ms.value

    for _ in range(args.warmup):
        step_dev()
    clocks = ClockSampler(local_rank)
    H.call("nls_sync"); barrier()
    clocks.start()
    l0 = L.nls_kernel_launches(H.h)
    t_wall0 = time.perf_counter()
    times = [step_dev() for _ in range(args.steps)]
    H.call("nls_sync"); barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.nls_kernel_launches(H.h) - l0
    ms_step = allmax(float(np.sum(times)) / args.steps)
    # extend the clock sample over a longer loaded window (the K steps alone are a few ms)
    for _ in range(2500):          # fixed count on every rank: each assembly holds a halo exchange
        H.call("nls_dev_assemble", 1, 1)
    H.call("nls_sync")
    clk = clocks.stop()
    total_dof = N_SIDE * N_SIDE * world * DIM

    # ---- e2e: host buffers through the reference-facing call ----
    try:
        import torch
        pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    except Exception:
        pin = lambda n: np.empty(n)
    Uh, Rh, Vh = pin(fem.ndof), pin(nrows), pin(nnz)
    Uh[:] = U0
    e2e_steps = 0 if args.skip_e2e else args.steps
    for _ in range(0 if args.skip_e2e else max(1, args.warmup // 2)):
        H.call("nls_residual_jacobian", dp(Uh), dp(Rh), dp(Vh))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        H.call("nls_residual_jacobian", dp(Uh), dp(Rh), dp(Vh))
    e2e_s = allmax((time.perf_counter() - t0) / max(1, e2e_steps)) or 1e-30
    barrier()

    # ---- Newton step (assembly + PCG + update + norm), device-resident ----
    # initial guess: the affine field that interpolates the two Dirichlet faces (a zero guess would
    # invert the top element layer: the prescribed 1% compression exceeds the element height 1/707)
    height = (N_SIDE * world - 1) / (N_SIDE - 1)
    Uaff = np.zeros((mesh.nn, DIM)); Uaff[:, DIM - 1] = -0.1 * 0.1 * mesh.coords[:, DIM - 1]
    Uaff = Uaff.ravel()
    H.call("nls_dev_upload_u", dp(Uaff))
    li, nr = C.c_int32(0), C.c_double(0)
    newton = None
    if not args.skip_newton:
        H.call("nls_dev_newton_step", args.lin_rtol, 200, C.byref(li), C.byref(nr))      # warm-up (short)
        H.call("nls_dev_upload_u", dp(Uaff))
        H.call("nls_sync"); barrier()
        H.call("nls_timer_start")
        H.call("nls_dev_newton_step", args.lin_rtol, 50000, C.byref(li), C.byref(nr))
        H.call("nls_timer_stop", C.byref(ms))
        nstep_ms = allmax(ms.value)
        spmv_bytes = 12 * nnz + 24 * nrows
        pcg_iter_bytes = spmv_bytes + 128 * nrows
        abytes = algorithmic_bytes(DIM, mesh.nen, mesh.nel, mesh.nn, nrows, nnz)
        alg = 2 * abytes - 8 * nnz + li.value * pcg_iter_bytes + 32 * nrows
        hbm, _ = measured_peaks()
        newton = {"ms": nstep_ms, "pcg_iterations": li.value, "lin_rtol": args.lin_rtol,
                  "residual_norm_after": nr.value, "algorithmic_GB": alg / 1e9,
                  "achieved_GBs": alg / nstep_ms / 1e6, "frac_of_hbm": alg / nstep_ms / 1e6 / hbm}

    abytes = algorithmic_bytes(DIM, mesh.nen, mesh.nel, mesh.nn, nrows, nnz)
    hbm, which = measured_peaks()
    achieved = abytes / (float(np.mean(times)) * 1e-3) / 1e9
    out = None
    if rank == 0:
        out = {
            "metric": "assembled DoF/s (residual+Jacobian)", "value": total_dof / (ms_step * 1e-3), "unit": "DoF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 2-D plane-strain Neo-Hookean, structured Q1, %d x %d nodes (%d DoF), fused residual+Jacobian assembly"
                       % (N_SIDE, N_SIDE * world, total_dof), "per_gpu_dof": N_SIDE * N_SIDE * DIM,
                       "nnz_per_gpu": nnz, "partition": "y-slabs, one per GPU" if world > 1 else "single GPU",
                       "cache": "L2 flushed (256 MiB memset) between timed iterations"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": NCU_TRAFFIC_BYTES if args.assembly in (None, 1) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                           "(profiles/r1k_ncu_asm2d_gather_v11.txt): 79.4 MB read + 96.7 MB written; the reads include ~45 MB of staged "
                                           "cluster plan (node lists, per-block source lists) on top of U and the coordinates, while part of the "
                                           "freshly written CSR values is still in the 126 MB L2 when the kernel ends",
                         "kernel": "k_asm_q1_2d_gather<R,K> (one launch = one fused assembly pass; no zero-fill needed)"
                                   if args.assembly in (None, 1) else "k_asm_q1_2d<R,K> + zero-fill of R and CSR values",
                         "algorithmic_bytes": abytes, "peak_source": which},
            "e2e": {"value": total_dof / e2e_s, "unit": "DoF/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": 8 * fem.ndof, "d2h_bytes_per_step": 8 * (fem.ndof + nrows + nnz)},
            "gpu_launches": int(launches), "clocks": clk, "newton_step": newton,
            "comm": {"path": {0: "single GPU", 1: "NCCL send/recv + allreduce", 2: "peer-memory windows over NVLink (CUDA IPC)"}[comm_path],
                     "latency": comm_us},
            "wall_s_timed_region": t_wall,
        }
    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.skip_cpu:
        out["cpu_baseline"] = cpu_baseline(nls, sample_rows=args.cpu_rows)
    if rank == 0:
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(nls, sample_rows=N_SIDE, steps=1, threads=None):
    """Oracle (C restatement of the reference algorithm: separate residual and Jacobian passes,
    CSR storage) on the host cores; a bounded sample = the first `sample_rows` node rows of C2."""
    from oracle import oracle as orc
    threads = threads or os.cpu_count() or 1
    shape = (N_SIDE, sample_rows)
    mesh = nls.grid_mesh(shape, DIM)
    mat = nls.NeoHookean(E=E_MOD, nu=NU)
    om = orc.OracleModel(DIM, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam), nthreads=threads)
    U = nls.smooth_field(mesh.coords)
    om.pattern()
    om.residual(U); om.jacobian_csr(U)          # warm-up (page-in)
    best = 1e30
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        om.residual(U)
        om.jacobian_csr(U)
        best = min(best, time.perf_counter() - t0)
    ndof = mesh.nn * DIM
    return {"value": ndof / best, "unit": "DoF/s", "cores": threads, "kind": "port", "seconds": best,
            "sample": "%d x %d-node slab of the C2 mesh (%d DoF), residual pass + Jacobian pass, CSR" % (N_SIDE, sample_rows, ndof)}


def run_reference(args):
    """The reference's CPU path, restated in C (oracle/), timed on the host cores with all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    nls = g.load_package()   # host-side mesh generators only; no GPU, no libnls_b200 compute
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    for _ in range(max(0, args.warmup - 1)):
        cpu_baseline(nls, sample_rows=args.cpu_rows, threads=threads)
    t0 = time.perf_counter()
    vals = [cpu_baseline(nls, sample_rows=args.cpu_rows, threads=threads) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = float(np.median([x["value"] for x in vals]))
    cb = dict(vals[-1]); cb["value"] = v
    total_dof = N_SIDE * N_SIDE * world * DIM
    print(json.dumps({
        "impl": "reference", "metric": "assembled DoF/s (residual+Jacobian)", "value": v, "unit": "DoF/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.median([x["seconds"] for x in vals])),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 2-D plane-strain Neo-Hookean, structured Q1, %d x %d nodes (%d DoF), residual+Jacobian assembly"
                   % (N_SIDE, N_SIDE * world, total_dof),
                   "note": "CPU restatement (oracle/) of the reference algorithm on %d host threads; Julia is not installed. "
                           "Each step is a bounded sample: %s" % (threads, cb["sample"])},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lin-rtol", dest="lin_rtol", type=float, default=1e-8)
    ap.add_argument("--skip-newton", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--gather-ctas", dest="gather_ctas", type=int, default=0)
    ap.add_argument("--assembly", type=int, default=None, help="0 atomic scatter, 1 gather/row-owner (default)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-rows", dest="cpu_rows", type=int, default=N_SIDE)
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
