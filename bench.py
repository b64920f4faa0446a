#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Newton hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong] [--config C3|C2]

Metric (BASELINE.json): assembled DoF/s of one fused residual+Jacobian evaluation, plus the Newton-step time, on the
north-star workload C3: 3-D Q1 hexahedral compressible Neo-Hookean cube, structured, 150 nodes per side = 10 125 000 DoF
per GPU (weak scaling stacks one such block per rank along z; --scaling strong splits the one cube in z-slabs).

A "step" is one residual+Jacobian assembly pass over the whole (local) mesh:
  value        device-resident (U, R, CSR values stay in HBM), CUDA-event timed on the library's stream, L2 flushed between
               iterations, max over ranks.
  e2e          the same pass through the reference-facing call nls_residual_jacobian with HOST buffers (pinned): H2D of U
               and D2H of U, R and the CSR values inside the timed region.
  parity       outside the timed region, on every rank: the owned rows of two node planes in the middle of the rank's
               slab (R and CSR rows) against the CPU oracle run on the five-plane sub-mesh around them.
  newton_step  full Newton solve of the load step of SURVEY 8d (face z=0 clamped, opposite face pushed by -0.1*t at
               t = 0.1, affine initial guess), rtol 1e-10, device-resident; per-iteration time and PCG counts.
  c2           the round-1 headline (2-D plane strain, 708^2 nodes) for continuity.
  cpu_baseline the oracle's C restatement of the reference algorithm on the box's host cores (bounded sample).
--impl reference times that CPU restatement alone (Julia is not installed, so the reference itself cannot run).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {"C3": dict(dim=3, n=150, name="C3: 3-D Q1 hexahedral compressible Neo-Hookean cube, structured"),
           "C2": dict(dim=2, n=708, name="C2: 2-D plane-strain Q1 compressible Neo-Hookean, structured")}
E_MOD, NU = 1.0, 0.3
LOAD_T = 0.1                      # prescribed compression -0.1 * t, single step t = 0.1


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg_key, kernel_hint):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu summary (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ent = json.load(f).get(cfg_key)
        if ent and (kernel_hint is None or kernel_hint in ent.get("kernel", "")):
            return float(ent["dram_bytes"]), ent.get("source")
    except Exception:
        pass
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the GPU is loaded."""

    def __init__(self, gpu):
        self.gpu, self.stop_flag, self.thread = gpu, False, None
        self.sm, self.smax, self.reasons, self.err = [], None, set(), None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:      # noqa: BLE001
            self.err = "nvml unavailable: %s" % e
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:   # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nm, bit in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
            time.sleep(0.05)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.smax, "samples": 0, "reasons": [self.err or "no samples"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.smax, "samples": len(self.sm),
                "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(gpu):
    """Pin this process (and so the pinned host buffers it first-touches) to the CPUs of the GPU's NUMA node: at N > 1 every
    rank moves gigabytes of CSR values to the host at once, and remote-socket buffers halve that rate. What `numactl
    --cpunodebind --membind` would do for a user of the C ABI. Returns a short description for the JSON line."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(gpu)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:            # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "numa node %d (%d cpus)" % (node, len(allowed))
        return "numa node %d (no allowed cpu)" % node
    except Exception as e:      # noqa: BLE001
        return "not bound: %s" % e


def _set_preferred_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread (node < 0: the default policy); errno, 0 = ok."""
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    if node < 0:
        r = libc.syscall(238, 0, None, 0)                      # SYS_set_mempolicy (x86_64), MPOL_DEFAULT
    else:
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        r = libc.syscall(238, 1, mask, 16 * 64 + 1)            # MPOL_PREFERRED: soft, falls back to other nodes
    return 0 if r == 0 else ctypes.get_errno()


def choose_host_node(numa):
    """Where should this rank's pinned host buffers live? If sysfs could not say (`numa` = "numa node unknown": the boxes
    report numa_node = -1) and the machine has several NUMA nodes, measure: a 256 MB pinned buffer per node under a
    preferred-node memory policy, device -> host copies into it, the fastest node wins (tools/numa_probe.py is the same
    probe stand-alone). Returns (node or None, description). Never raises."""
    try:
        import glob
        import torch
        nodes = sorted(int(q.rsplit("node", 1)[1]) for q in glob.glob("/sys/devices/system/node/node[0-9]*"))
        if numa != "numa node unknown" or (len(nodes) < 2 and not os.environ.get("NLS_BENCH_FORCE_NUMA_MEASURE")):
            return None, numa if numa != "numa node unknown" else "single NUMA node"
        n = 32 << 20
        dev = torch.ones(n, dtype=torch.float64, device="cuda")
        rate = {}
        for node in nodes:
            if _set_preferred_node(node):
                continue
            try:
                host = torch.empty(n, dtype=torch.float64).pin_memory()
            finally:
                _set_preferred_node(-1)
            best = 0.0
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
                best = max(best, 8.0 * n / (time.perf_counter() - t0) / 1e9)
            rate[node] = best
            del host
        del dev
        if not rate:
            return None, "numa node unknown (set_mempolicy refused)"
        node = max(rate, key=rate.get)
        return node, "numa node %d by measured D2H rate (%s GB/s)" % (node, ", ".join("%d: %.1f" % kv for kv in sorted(rate.items())))
    except Exception as e:      # noqa: BLE001
        return None, "numa node unknown (%s)" % e


def algorithmic_bytes(dim, nen, nel, nn, ndof_rows, nnz):
    """conn + coords + U read + R write + vals write (colind / slot maps not counted)."""
    return 4 * nen * nel + 8 * dim * nn + 8 * ndof_rows + 8 * ndof_rows + 8 * nnz


def global_shape(cfg, nranks, scaling):
    n, dim = cfg["n"], cfg["dim"]
    return (n,) * (dim - 1) + ((n * nranks) if scaling == "weak" else n,)


def workload_string(cfg, nranks, scaling):
    sh = global_shape(cfg, nranks, scaling)
    return "%s, %s nodes (%d DoF), fused residual+Jacobian assembly" % (
        cfg["name"], " x ".join(str(s) for s in sh), int(np.prod(sh)) * cfg["dim"])


def build_local_model(nls, cfg, nranks, rank, device, uid, scaling):
    """The rank's slab of the structured mesh with the load step of SURVEY 8d: slowest-axis face 0 clamped, opposite face
    pushed by -0.1 t: the SAME displacement at every N. Weak scaling stacks the blocks into a column of aspect ratio N; holding
    the strain at 1 % instead would put it past its Euler buckling strain (0.2056 / N^2 for this clamped / laterally free
    column: 0.32 % at N = 8), where the tangent is indefinite and any CG breaks down -- measured at N = 8 in round 2."""
    dim, n = cfg["dim"], cfg["n"]
    shape = global_shape(cfg, nranks, scaling)
    mat = nls.NeoHookean(E=E_MOD, nu=NU)
    plane = int(np.prod(shape[:-1]))
    if nranks == 1:
        mesh = nls.grid_mesh(shape, dim)
        fem = nls.FEMModel(mesh, device=device)
        n_owned, p0, p1 = mesh.nn, 0, shape[-1]
    else:
        sl = nls.slab_local_mesh(shape, dim, nranks, rank)
        mesh = sl["mesh"]
        fem = nls.FEMModel(mesh, device=device)
        fem.set_partition(sl["n_owned"], sl["neigh"], sl["send_ptr"], sl["send_nodes"], sl["recv_ptr"], sl["recv_nodes"])
        fem.set_comm(rank, nranks, uid)
        n_owned = sl["n_owned"]
        cuts = nls.slab_planes(shape[-1], nranks)
        p0, p1 = cuts[rank], cuts[rank + 1]
    height = (shape[-1] - 1) / (n - 1)
    if p0 == 0:
        d = (np.arange(plane)[:, None] * dim + np.arange(dim)[None, :]).ravel()
        fem.addboundary(nls.Dirichlet(d, np.zeros(len(d))))
    if p1 == shape[-1] and p1 > p0:
        t_nodes = (shape[-1] - 1 - p0) * plane + np.arange(plane)
        d = t_nodes * dim + (dim - 1)
        fem.addboundary(nls.Dirichlet(d, np.full(len(d), -0.1 * LOAD_T)))
    info = dict(shape=shape, plane=plane, p0=p0, p1=p1, n_owned=n_owned, height=height)
    return fem, mat, mesh, info


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


dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))      # noqa: E731


def parity_check(nls, cfg, info, mesh, H, U_local, threads):
    """Owned rows of two node planes in the middle of this rank's slab -- R and the CSR rows, pattern included -- against
    the CPU oracle on the five-plane sub-mesh around them (every element that touches the two planes is inside it)."""
    from oracle import oracle as orc
    dim, plane, p0, p1 = cfg["dim"], info["plane"], info["p0"], info["p1"]
    if p1 - p0 < 5:
        return None
    zc = (p0 + p1) // 2
    za, zb = zc - 2, zc + 3                      # sub-mesh node planes [za, zb)
    sub_shape = info["shape"][:-1] + (zb - za,)
    sub = nls.grid_mesh(sub_shape, dim)
    lo_node = (za - p0) * plane                  # local id of the sub-mesh's first node (owned planes are numbered in order)
    sub.coords[:] = mesh.coords[lo_node:lo_node + sub.nn]
    mat = nls.NeoHookean(E=E_MOD, nu=NU)
    om = orc.OracleModel(dim, sub.conn, sub.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam), nthreads=threads)
    Usub = U_local[lo_node * dim:(lo_node + sub.nn) * dim].copy()
    _, Ro = om.residual(Usub)
    Ko = om.jacobian_csr(Usub)
    Rabs, Kabs = om.contribution_scale(Usub)
    rpo, cio = om.pattern()
    r0s, r1s = 1 * plane * dim, 3 * plane * dim                  # sub-mesh rows of planes zc-1, zc
    row0, row1 = (zc - 1 - p0) * plane * dim, (zc + 1 - p0) * plane * dim
    e0, e1 = rpo[r0s], rpo[r1s]
    R = np.zeros(row1 - row0); rp = np.zeros(row1 - row0 + 1, dtype=np.int64)
    ci = np.zeros(e1 - e0, dtype=np.int32); vals = np.zeros(e1 - e0)
    H.call("nls_dev_download_rows", row0, row1, dp(R), rp.ctypes.data_as(C.POINTER(C.c_int64)),
           ci.ctypes.data_as(C.POINTER(C.c_int32)), dp(vals))
    pattern_ok = bool(np.array_equal(rp - rp[0], rpo[r0s:r1s + 1] - e0) and
                      np.array_equal(ci.astype(np.int64) - lo_node * dim, cio[e0:e1]))
    Ros, Kos = Ro[r0s:r1s], Ko[e0:e1]
    out = {"rows": int(row1 - row0), "entries": int(e1 - e0), "planes": [int(zc - 1), int(zc)], "pattern_bit_exact": pattern_ok,
           "R_rel": float(np.abs(R - Ros).max() / np.abs(Ros).max()),
           "K_rel": float(np.abs(vals - Kos).max() / np.abs(Kos).max()),
           "R_entrywise_rel": float((np.abs(R - Ros) / np.where(Rabs[r0s:r1s] > 0, Rabs[r0s:r1s], 1.0)).max()),
           "K_entrywise_rel": float((np.abs(vals - Kos) / np.where(Kabs[e0:e1] > 0, Kabs[e0:e1], 1.0)).max()),
           "oracle": "oracle/nls_oracle.c on the %s-node sub-mesh, %d threads" % (" x ".join(str(s) for s in sub_shape), threads)}
    return out


def time_assembly(H, steps, warmup, barrier):
    ms = C.c_float(0)

    def step_dev():
        H.call("nls_flush_l2")
        H.call("nls_dev_barrier")        # N > 1: line the ranks' streams up, so the halo wait inside the step is not host jitter
        H.call("nls_timer_start")
        H.call("nls_dev_assemble", 1, 1)
        H.call("nls_timer_stop", C.byref(ms))
        return ms.value
    for _ in range(warmup):
        step_dev()
    H.call("nls_sync"); barrier()
    return [step_dev() for _ in range(steps)]


def run_ours(args):
    emit = _quiet_stdout()
    import __graft_entry__ as g
    nls = g.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    dim = cfg["dim"]
    numa = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_uid():
        if dist is None:
            return None
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            rc = nls.lib().nls_comm_unique_id(nls._lib.nccl_lib_path().encode(), buf)
            assert rc == 0, nls.lib().nls_last_error(None)
            idt.copy_(torch.tensor(list(bytes(buf)), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tolist())

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

    def gather_obj(o):
        if dist is None:
            return [o]
        lst = [None] * world
        dist.all_gather_object(lst, o)
        return lst

    t_setup = time.perf_counter()
    fem, mat, mesh, info = build_local_model(nls, cfg, world, rank, local_rank, new_uid(), args.scaling)
    H = fem.handle(mat)
    L = nls.lib()
    setup_s = time.perf_counter() - t_setup
    if args.assembly is not None:
        H.call("nls_set_option", b"assembly", args.assembly)
    scheme = L.nls_assembly_scheme(H.h)
    comm_path = L.nls_comm_path(H.h)
    comm_us = None
    if world > 1:
        uh, ua = C.c_float(0), C.c_float(0)
        H.call("nls_dev_comm_probe", 200, C.byref(uh), C.byref(ua))
        comm_us = {"halo_exchange_us": uh.value, "allreduce_us": ua.value}
    sys.stderr.write("rank %d: setup %.1f s, assembly scheme %d (0 atomic, 1 gather, 2 element records, 3 line pairs), comm path %d %s\n"
                     % (rank, setup_s, scheme, comm_path, comm_us))
    nrows, nnz = C.c_int64(0), C.c_int64(0)
    H.call("nls_pattern_size", C.byref(nrows), C.byref(nnz))
    nrows, nnz = nrows.value, nnz.value
    U0 = nls.smooth_field(mesh.coords)
    H.call("nls_dev_upload_u", dp(U0))

    # ---- value: device-resident fused R+K assembly ----
    times = time_assembly(H, 0, args.warmup, barrier)
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = L.nls_kernel_launches(H.h)
    t_wall0 = time.perf_counter()
    times = time_assembly(H, args.steps, 0, barrier)
    H.call("nls_sync"); barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.nls_kernel_launches(H.h) - l0
    ms_local = float(np.sum(times)) / args.steps
    ms_step = allmax(ms_local)
    # keep the GPU loaded for ~1.5 s more so the clock sample covers a loaded window (same count on every rank)
    for _ in range(int(min(4000, max(20, 1500.0 / max(ms_step, 1e-3))))):
        H.call("nls_dev_assemble", 1, 1)
    H.call("nls_sync")
    clk = clocks.stop()
    total_dof = int(np.prod(info["shape"])) * dim
    kernel = {0: "atomic scatter: k_asm_q1_%dd<R,K> + zero-fill of R and the CSR values" % dim,
              1: "k_asm_q1_2d_gather<R,K> (one launch = one fused pass, no zero-fill)",
              2: "k_elem_q1_store + k_rows_from_records", 3: "k_asm_q1_3d_lines<R,K> (one launch = one fused pass, no zero-fill, no atomics)"}[scheme]

    # ---- parity against the CPU oracle (outside the timed region), every rank ----
    parity = None
    if not args.skip_parity:
        Udev = np.zeros(fem.ndof)
        H.call("nls_dev_download_u", dp(Udev))       # as assembled: Dirichlet values written, ghosts exchanged
        try:
            par = parity_check(nls, cfg, info, mesh, H, Udev, max(1, (os.cpu_count() or 1) // world))
        except Exception as e:      # noqa: BLE001
            par = {"error": repr(e)}
        allp = gather_obj(par)
        if rank == 0:
            good = [p for p in allp if p and "error" not in p]
            parity = {k: max(p[k] for p in good) for k in ("R_rel", "K_rel", "R_entrywise_rel", "K_entrywise_rel")} if good else {}
            parity.update({"pattern_bit_exact": all(p["pattern_bit_exact"] for p in good) if good else None,
                           "ranks_checked": len(good), "rows_per_rank": good[0]["rows"] if good else 0,
                           "entries_per_rank": good[0]["entries"] if good else 0, "oracle": good[0]["oracle"] if good else None,
                           "tolerance": 1e-12, "errors": [p["error"] for p in allp if p and "error" in p]})

    # ---- e2e: host buffers through the reference-facing call ----
    e2e = None
    if not args.skip_e2e:
        try:
            import torch
            pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory().numpy()     # noqa: E731
        except Exception:       # noqa: BLE001
            pin = lambda n: np.empty(n)     # noqa: E731
        node, numa = choose_host_node(numa)
        if node is not None:
            _set_preferred_node(node)
        try:
            Uh, Rh, Vh = pin(fem.ndof), pin(nrows), pin(nnz)
        finally:
            if node is not None:
                _set_preferred_node(-1)
        Uh[:] = U0
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        H.call("nls_residual_jacobian", dp(Uh), dp(Rh), dp(Vh))          # warm-up (page-in of the pinned buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            H.call("nls_residual_jacobian", dp(Uh), dp(Rh), dp(Vh))
        e2e_s = allmax((time.perf_counter() - t0) / e2e_steps)
        barrier()
        e2e = {"value": total_dof / e2e_s, "unit": "DoF/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "host_buffers": "pinned, " + numa,
               "h2d_bytes_per_step": 8 * fem.ndof, "d2h_bytes_per_step": 8 * (fem.ndof + nrows + nnz)}
        del Uh, Rh, Vh

    # ---- Newton: the load step of SURVEY 8d, full solve, device-resident ----
    newton = None
    if not args.skip_newton:
        # initial guess: the affine field that interpolates the two Dirichlet faces (a zero guess would invert the top
        # element layer: the prescribed compression exceeds the element height)
        Uaff = np.zeros((mesh.nn, dim)); Uaff[:, dim - 1] = -0.1 * LOAD_T * mesh.coords[:, dim - 1] / info["height"]
        Uaff = Uaff.ravel()
        H.call("nls_dev_upload_u", dp(Uaff))
        H.call("nls_solver_prepare")          # work space of the Krylov preconditioner: allocated once per model, not per solve
        opts = nls._lib.NewtonOpts(rtol=1e-10, atol=1e-14, dtol=1e6, maxits=25, type_modified=0,
                                   lin_rtol=args.lin_rtol, lin_maxits=100000)
        res = nls._lib.NewtonResult()
        hist = np.zeros(64)
        H.call("nls_sync"); barrier()
        H.call("nls_dev_newton_solve", C.byref(opts), dp(hist), C.byref(res))
        tot_ms = allmax(res.ms_total)
        its = max(1, res.iterations)
        spmv_bytes = 12 * nnz + 24 * nrows
        precond = int(L.nls_krylov_preconditioner(H.h))
        if precond == 2:
            # multigrid-preconditioned CG: one FP64 product, 2 x degree products on the single-precision copy of the values
            # (4 B value + 4/9 B block column per nonzero), ~60 vector passes on level 0; coarser levels add about 1/7
            pcg_iter_bytes = spmv_bytes + (8.0 / 7.0) * (4 * ((4 + 4.0 / 9.0) * nnz + 24 * nrows) + 60 * 8 * nrows)
        else:
            pcg_iter_bytes = spmv_bytes + 128 * nrows
        abytes = algorithmic_bytes(dim, mesh.nen, mesh.nel, mesh.nn, nrows, nnz)
        alg = (res.iterations + 1) * abytes + res.lin_its_total * pcg_iter_bytes + res.iterations * 32 * nrows
        hbm, _ = measured_peaks()
        newton = {"ms": tot_ms / its, "total_ms": tot_ms, "iterations": int(res.iterations), "reason": int(res.reason),
                  "converged": int(res.reason) in (1, 2), "pcg_iterations": int(res.lin_its_total),
                  "preconditioner": {0: "Jacobi", 1: "block-Jacobi", 2: "geometric multigrid V-cycle (Chebyshev / block-Jacobi smoothing, Galerkin coarse operators)"}.get(precond, str(precond)),
                  "ms_assembly": allmax(res.ms_assembly), "ms_linear": allmax(res.ms_linear),
                  "rtol": 1e-10, "atol": 1e-14, "lin_rtol": args.lin_rtol,
                  "load": "face 0 of the slowest axis clamped, opposite face displaced by -0.1*t at t=0.1 (1 %% of one block's height: the same displacement at every N), affine initial guess",
                  "residual_norms": [float(v) for v in hist[:res.nhist]],
                  "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / tot_ms / 1e6, "frac_of_hbm": alg / tot_ms / 1e6 / hbm}

    # ---- C2 (round-1 headline) for continuity: assembly only ----
    c2 = None
    if args.config == "C3" and not args.skip_c2:
        try:
            cfg2 = CONFIGS["C2"]
            fem2, mat2, mesh2, info2 = build_local_model(nls, cfg2, world, rank, local_rank, new_uid(), "weak")
            H2 = fem2.handle(mat2)
            H2.call("nls_dev_upload_u", dp(nls.smooth_field(mesh2.coords)))
            t2 = time_assembly(H2, max(10, args.steps), 5, barrier)
            ms2 = allmax(float(np.mean(t2)))
            nr2, nz2 = C.c_int64(0), C.c_int64(0)
            H2.call("nls_pattern_size", C.byref(nr2), C.byref(nz2))
            ab2 = algorithmic_bytes(2, mesh2.nen, mesh2.nel, mesh2.nn, nr2.value, nz2.value)
            hbm, _ = measured_peaks()
            dof2 = int(np.prod(info2["shape"])) * 2
            tr2, _src = ncu_traffic("C2", None)
            c2 = {"workload": workload_string(cfg2, world, "weak"), "value": dof2 / (ms2 * 1e-3), "unit": "DoF/s", "ms_per_step": ms2,
                  "roofline_frac": ab2 / (float(np.mean(t2)) * 1e-3) / 1e9 / hbm, "algorithmic_bytes": ab2, "traffic": tr2}
            if not args.skip_newton:     # the same load step on C2: Newton-step time of BASELINE config 2
                U2 = np.zeros((mesh2.nn, 2)); U2[:, 1] = -0.1 * LOAD_T * mesh2.coords[:, 1] / info2["height"]
                H2.call("nls_dev_upload_u", dp(np.ascontiguousarray(U2.ravel())))
                H2.call("nls_solver_prepare")
                o2 = nls._lib.NewtonOpts(rtol=1e-10, atol=1e-14, dtol=1e6, maxits=25, type_modified=0, lin_rtol=args.lin_rtol, lin_maxits=100000)
                r2 = nls._lib.NewtonResult()
                h2 = np.zeros(64)
                H2.call("nls_sync"); barrier()
                H2.call("nls_dev_newton_solve", C.byref(o2), dp(h2), C.byref(r2))
                tot2 = allmax(r2.ms_total)
                c2["newton_step"] = {"ms": tot2 / max(1, r2.iterations), "total_ms": tot2, "iterations": int(r2.iterations),
                                     "converged": int(r2.reason) in (1, 2), "krylov_iterations": int(r2.lin_its_total),
                                     "preconditioner": int(L.nls_krylov_preconditioner(H2.h)), "rtol": 1e-10, "lin_rtol": args.lin_rtol}
            del H2, fem2
        except Exception as e:      # noqa: BLE001
            c2 = {"error": repr(e)}

    abytes = algorithmic_bytes(dim, mesh.nen, mesh.nel, mesh.nn, nrows, nnz)
    hbm, which = measured_peaks()
    achieved = abytes / (ms_local * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.config, "lines" if scheme == 3 else ("gather" if scheme == 1 else None))
    out = None
    if rank == 0:
        out = {
            "metric": "assembled DoF/s (residual+Jacobian)", "value": total_dof / (ms_step * 1e-3), "unit": "DoF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(cfg, world, args.scaling), "per_gpu_dof": int(nrows),
                       "nnz_per_gpu": int(nnz), "partition": "z-slabs, one per GPU" if world > 1 else "single GPU",
                       "cache": "L2 flushed (256 MiB memset) between timed iterations", "setup_s": setup_s},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel,
                         "algorithmic_bytes": abytes, "peak_source": which},
            "e2e": e2e, "parity": parity,
            "gpu_launches": int(launches), "clocks": clk, "newton_step": newton, "c2": c2,
            "comm": {"path": {0: "single GPU", 1: "NCCL send/recv + allreduce", 2: "peer-memory windows over NVLink (CUDA IPC)"}[comm_path],
                     "latency": comm_us},
            "wall_s_timed_region": t_wall,
        }
    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.skip_cpu:
        out["cpu_baseline"] = cpu_baseline(nls, cfg, planes=args.cpu_planes)
    if rank == 0:
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(nls, cfg, planes, steps=1, threads=None):
    """Oracle (C restatement of the reference algorithm: separate residual and Jacobian passes, CSR storage) on the host
    cores; a bounded sample = the first `planes` node planes of the workload's mesh."""
    from oracle import oracle as orc
    threads = threads or os.cpu_count() or 1
    dim, n = cfg["dim"], cfg["n"]
    shape = (n,) * (dim - 1) + (planes,)
    mesh = nls.grid_mesh(shape, dim)
    mat = nls.NeoHookean(E=E_MOD, nu=NU)
    om = orc.OracleModel(dim, mesh.conn, mesh.coords, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mat.mu, mat.lam), nthreads=threads)
    U = nls.smooth_field(mesh.coords)
    om.pattern()
    om.residual(U); om.jacobian_csr(U)          # warm-up (page-in)
    best = 1e30
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        om.residual(U)
        om.jacobian_csr(U)
        best = min(best, time.perf_counter() - t0)
    ndof = mesh.nn * dim
    return {"value": ndof / best, "unit": "DoF/s", "cores": threads, "kind": "port", "seconds": best,
            "sample": "the first %d node planes of the mesh (%s nodes, %d DoF), residual pass + Jacobian pass, CSR"
                      % (planes, " x ".join(str(s) for s in shape), ndof)}


def run_reference(args):
    """The reference's CPU path, restated in C (oracle/), timed on the host cores with all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    nls = g.load_package()   # host-side mesh generators only; no GPU, no libnls_b200 compute
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    for _ in range(max(0, args.warmup - 1)):
        cpu_baseline(nls, cfg, planes=args.cpu_planes, threads=threads)
    t0 = time.perf_counter()
    vals = [cpu_baseline(nls, cfg, planes=args.cpu_planes, threads=threads) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = float(np.median([x["value"] for x in vals]))
    cb = dict(vals[-1]); cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "assembled DoF/s (residual+Jacobian)", "value": v, "unit": "DoF/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.median([x["seconds"] for x in vals])),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(cfg, world, args.scaling),
                   "note": "CPU restatement (oracle/) of the reference algorithm on %d host threads; Julia is not installed. The rate "
                           "is measured on a bounded sample of the workload at every N (the CPU path does not scale with the GPU "
                           "count): %s" % (threads, cb["sample"])},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--lin-rtol", dest="lin_rtol", type=float, default=1e-8)
    ap.add_argument("--skip-newton", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-c2", action="store_true")
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=5)
    ap.add_argument("--assembly", type=int, default=None, help="0 atomic scatter, 1 best atomic-free scheme (default)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-planes", dest="cpu_planes", type=int, default=0, help="node planes of the CPU sample (0: 33 in 3-D, 708 in 2-D)")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.cpu_planes <= 0:
        a.cpu_planes = 33 if CONFIGS[a.config]["dim"] == 3 else CONFIGS[a.config]["n"]
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
