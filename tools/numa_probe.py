"""Which NUMA node should the pinned host buffers of a GPU live on? sysfs has no answer inside the GPU boxes (numa_node = -1),
so measure: allocate a pinned buffer under a preferred-node memory policy and time device -> host copies into it.
Usage: python tools/numa_probe.py [gpu]"""
import ctypes, glob, os, sys, time
import torch

MPOL_DEFAULT, MPOL_PREFERRED = 0, 1
SYS_set_mempolicy = 238      # x86_64
libc = ctypes.CDLL(None, use_errno=True)


def set_preferred(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread; node < 0: default policy. Returns errno (0 = ok)."""
    if node < 0:
        r = libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    else:
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        r = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, 16 * 64 + 1)
    return 0 if r == 0 else ctypes.get_errno()


def main():
    gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    torch.cuda.set_device(gpu)
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("nodes", nodes, "affinity", len(os.sched_getaffinity(0)), "cpus", flush=True)
    n = 64 << 20      # 512 MB of doubles
    dev = torch.empty(n, dtype=torch.float64, device="cuda").fill_(1.0)
    for node in [-1] + nodes:
        e = set_preferred(node)
        if e:
            print("node", node, "set_mempolicy errno", e, os.strerror(e), flush=True)
            continue
        host = torch.empty(n, dtype=torch.float64).pin_memory()
        set_preferred(-1)
        best = 0.0
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
            best = max(best, 8.0 * n / (time.perf_counter() - t0) / 1e9)
        h2d = 0.0
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
            h2d = max(h2d, 8.0 * n / (time.perf_counter() - t0) / 1e9)
        print("node %d: D2H %.1f GB/s, H2D %.1f GB/s" % (node, best, h2d), flush=True)
        del host


if __name__ == "__main__":
    main()
