"""CPU tests (no GPU) of the product's host side: the C ABI loads and exports every declared
symbol, fails loudly without a device, and its host-only integer work (pattern, partition, slab
maps) is bit-exact against the oracle. Includes a world_size-2 gloo test of the multi-GPU plumbing."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(nls):
    hdr = open(os.path.join(ROOT, "include", "nls_b200.h")).read()
    declared = set(re.findall(r"^(?:int32_t|int64_t|const char \*)\s*(nls_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 45
    L = nls.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libnls_b200.so lacks " + name
    assert declared == set(nls._lib.PROTOTYPES), declared ^ set(nls._lib.PROTOTYPES)
    assert b"sm_100a" in L.nls_version()


def test_library_is_sm100a_cuda(nls):
    """The shipped binary holds sm_100a SASS for the hot-path kernels (no PTX-JIT, no CPU path)."""
    out = subprocess.run(["cuobjdump", "-lelf", nls._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", nls._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "k_asm_q1_2d" in sass and "k_asm_q1_3d" in sass and "k_spmv" in sass
    assert "DFMA" in sass and "REDG.E.ADD.F64" in sass


def test_no_gpu_fails_loudly(nls):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nls.NLSError) as e:
        nls.Handle(0)
    assert "NOGPU" in str(e.value) and "no CPU fallback" in str(e.value)
    fem = nls.FEMModel(nls.grid_mesh(3, 2))
    with pytest.raises(nls.NLSError):
        nls.Residual(nls.NeoHookean(1.0, 1.0))(fem, np.zeros(fem.ndof), np.zeros(fem.ndof))


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "nonlinearsolids.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, os.path.join(dp, f)


def _host_pattern(nls, mesh, n_owned=None):
    n_owned = mesh.nn if n_owned is None else n_owned
    L = nls.lib()
    rowptr = np.zeros(n_owned * mesh.dim + 1, dtype=np.int64)
    nnz = C.c_int64(0)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    assert L.nls_pattern_host(mesh.dim, mesh.nen, mesh.nel, ip(mesh.conn), mesh.nn, n_owned, lp(rowptr), None, C.byref(nnz)) == 0
    colind = np.zeros(nnz.value, dtype=np.int32)
    assert L.nls_pattern_host(mesh.dim, mesh.nen, mesh.nel, ip(mesh.conn), mesh.nn, n_owned, lp(rowptr), ip(colind), C.byref(nnz)) == 0
    return rowptr, colind


@pytest.mark.parametrize("kind", ["bar_p1", "bar_p3", "grid2", "grid3", "grid2_rect", "shuffled"])
def test_pattern_bit_exact(nls, orc, kind):
    if kind.startswith("bar"):
        p = int(kind[-1]); mesh = nls.bar_mesh(11, p)
        om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=2, mat_kind=orc.MAT_LINEAR, mat=(1.0,))
    else:
        mesh = {"grid2": nls.grid_mesh(9, 2), "grid3": nls.grid_mesh(5, 3), "grid2_rect": nls.grid_mesh((4, 7), 2),
                "shuffled": nls.grid_mesh(6, 3)}[kind]
        if kind == "shuffled":      # unstructured numbering: random node relabelling + element order
            rng = np.random.default_rng(1)
            perm = rng.permutation(mesh.nn)
            conn = perm[mesh.conn][rng.permutation(mesh.nel)].astype(np.int32)
            coords = np.zeros_like(mesh.coords); coords[perm] = mesh.coords
            mesh = nls.Mesh(3, conn, coords=coords)
        om = orc.OracleModel(mesh.dim, mesh.conn, mesh.coords)
    rp, ci = _host_pattern(nls, mesh)
    rpo, cio = om.pattern()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    assert rp.dtype == np.int64 and ci.dtype == np.int32
    n = mesh.nn
    if kind == "grid2":
        assert rp[-1] == 4 * (3 * 9 - 2) ** 2           # nnz = 4(3n-2)^2
    if kind == "grid3":
        assert rp[-1] == 9 * (3 * 5 - 2) ** 3           # nnz = 9(3n-2)^3
    for r in range(0, len(rp) - 1, max(1, n // 7)):     # sorted, unique columns
        row = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(row) > 0)


@pytest.mark.parametrize("dim,shape,P", [(2, (5, 9), 3), (3, (3, 4, 7), 4), (2, (4, 4), 2), (3, (4, 4, 8), 8), (2, (6, 3), 4)])
def test_partition_bit_exact(nls, orc, dim, shape, P):
    gm = nls.grid_mesh(shape, dim)
    om = orc.OracleModel(dim, gm.conn, gm.coords)
    L = nls.lib()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    owned_total = 0
    for r in range(P):
        sl = nls.slab_local_mesh(shape, dim, P, r)
        off = np.array(sl["node_off"], dtype=np.int64)
        ne, ng, sp = C.c_int64(), C.c_int64(), np.zeros(P + 1, dtype=np.int64)
        assert L.nls_partition(gm.nen, gm.nel, ip(gm.conn), P, lp(off), r, C.byref(ne), None, C.byref(ng), None, lp(sp), None) == 0
        le, gn, sn = np.zeros(ne.value, dtype=np.int64), np.zeros(ng.value, dtype=np.int64), np.zeros(sp[-1], dtype=np.int64)
        L.nls_partition(gm.nen, gm.nel, ip(gm.conn), P, lp(off), r, C.byref(ne), lp(le), C.byref(ng), lp(gn), lp(sp), lp(sn))
        want = om.partition(P, off, r)                                       # oracle: bit-exact maps
        assert np.array_equal(le, want["loc_elems"]) and np.array_equal(gn, want["ghost_nodes"])
        assert np.array_equal(sp, want["send_ptr"]) and np.array_equal(sn, want["send_nodes"])
        # the closed-form slab (what bench.py uses at scale) is the same partition
        l2g = sl["l2g"]
        assert np.array_equal(gn, l2g[sl["n_owned"]:])
        assert np.array_equal(l2g[sl["mesh"].conn], gm.conn[le])
        assert np.array_equal(np.sort(l2g[sl["send_nodes"]]), np.sort(sn))
        assert np.array_equal(sl["mesh"].coords, gm.coords[l2g])
        owned_total += sl["n_owned"]
        # owned-row pattern of the local mesh == the global pattern's rows of the owned nodes
        rp, ci = _host_pattern(nls, sl["mesh"], sl["n_owned"])
        grp, gci = om.pattern()
        lo = off[r] * dim
        for lr in range(0, sl["n_owned"] * dim, 3):
            gcols = gci[grp[lo + lr]:grp[lo + lr + 1]]
            lcols = ci[rp[lr]:rp[lr + 1]]
            assert np.array_equal(np.sort((l2g[lcols // dim] * dim + lcols % dim)), gcols)
    assert owned_total == gm.nn


def test_mesh_generators(nls):
    m = nls.grid_mesh(708, 2)
    assert m.nn == 501264 and m.nel == 499849 and m.nn * 2 == 1002528             # C2
    assert m.coords[707, 0] == 1.0 and m.coords[-1, 1] == 1.0 and m.coords[1, 0] == 1 / 707
    b = nls.bar_mesh(100, 3)
    assert b.nn == 301 and b.conn[-1].tolist() == [297, 298, 299, 300]            # C1
    assert nls.face_nodes(4, 3, 2, 1).tolist() == list(range(48, 64))
    assert nls.face_nodes((3, 5), 2, 0, 0).tolist() == [0, 3, 6, 9, 12]
    U = nls.smooth_field(nls.grid_mesh(3, 2).coords)
    assert U.shape == (18,) and abs(U[2 * 1 + 0] - 0.01 * np.sin(np.pi) * 1.0) < 1e-17
    ts_err = None
    try:
        nls.PseudoTimeStepper(nls.NewtonSolver(nls.FEMModel(b, 2, 3)), dt=0.3, maxtime=1.0)
    except ValueError as e:
        ts_err = e
    assert ts_err is not None and "InexactError" in str(ts_err)                  # pseudotime.jl:16


_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as g
nls = g.load_package()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
uid = nls.distributed.broadcast_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
assert uid == bytes(range(128))
for dim, shape in [(2, (5, 8)), (3, (3, 4, 6))]:
    part = nls.slab_local_mesh(shape, dim, world, rank)
    gm = nls.grid_mesh(shape, dim)
    gvals = (np.arange(gm.nn * dim) * 1.5 + 0.25)            # a global field, node-major interleaved
    l2g = part["l2g"]
    loc = gvals.reshape(-1, dim)[l2g].copy()
    loc[part["n_owned"]:] = -777.0                            # ghosts unknown before the exchange
    out = nls.distributed.exchange_halo_host(dist, part, loc.ravel(), dim)
    assert np.array_equal(out.reshape(-1, dim), gvals.reshape(-1, dim)[l2g]), (rank, dim)
    # owned-row sums over ranks reproduce a global quantity (what the Krylov allreduce does)
    t = torch.tensor([float(out.reshape(-1, dim)[:part["n_owned"]].sum())], dtype=torch.float64)
    dist.all_reduce(t)
    assert abs(t.item() - gvals.sum()) < 1e-9 * abs(gvals.sum())
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_halo_plumbing(tmp_path):
    """world_size 2 over gloo: id broadcast, and the send/recv lists of the slab partition move
    exactly the ghost DoFs (the exchange the library does with ncclSend/ncclRecv on the GPUs)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_bench_host_buffer_placement_is_safe():
    """bench.py's NUMA helpers never raise: a known node is passed through, the default memory policy can always be restored,
    and without CUDA the measured choice degrades to 'unknown' instead of failing the e2e leg."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    node, desc = bench.choose_host_node("numa node 1 (8 cpus)")
    assert node is None and desc == "numa node 1 (8 cpus)"
    assert isinstance(bench._set_preferred_node(-1), int)
    node, desc = bench.choose_host_node("numa node unknown")
    assert (node is None or isinstance(node, int)) and isinstance(desc, str) and desc
