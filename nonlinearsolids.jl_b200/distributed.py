"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the bootstrap only.

The data path (halo exchange of ghost DoFs, Krylov allreduces) runs inside libnls_b200.so on its
own NCCL communicator; torch.distributed just carries the 128-byte NCCL id from rank 0 to the
others and provides barriers for the benchmark.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .meshes import slab_local_mesh


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def make_unique_id():
    """ncclGetUniqueId through the library (rank 0 only)."""
    buf = (C.c_uint8 * 128)()
    L = _lib.lib()
    rc = L.nls_comm_unique_id(_lib.nccl_lib_path().encode(), buf)
    if rc != 0:
        raise _lib.NLSError(rc, L.nls_last_error(None).decode())
    return bytes(buf)


def broadcast_unique_id(dist, rank, make_id=make_unique_id, device=None):
    """Rank 0 creates the id, every rank returns the same 128 bytes. `dist` = torch.distributed
    (any backend: gloo in the CPU tests, nccl on the GPU box)."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        t.copy_(torch.tensor(list(make_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tolist())


def exchange_halo_host(dist, part, values, dim=1):
    """Host emulation of the ghost exchange the library does with ncclSend/ncclRecv, used by the
    CPU (gloo) tests of the partition maps. `values` is a local array [nn_local*dim]; returns it
    with the ghost entries filled from their owners."""
    import torch
    v = np.array(values, dtype=np.float64).reshape(-1, dim)
    reqs, recvs = [], []
    for q, peer in enumerate(part["neigh"]):
        s = part["send_nodes"][part["send_ptr"][q]:part["send_ptr"][q + 1]]
        r = part["recv_nodes"][part["recv_ptr"][q]:part["recv_ptr"][q + 1]]
        sb = torch.from_numpy(np.ascontiguousarray(v[s]))
        rb = torch.zeros((len(r), dim), dtype=torch.float64)
        reqs.append(dist.isend(sb, int(peer)))
        reqs.append(dist.irecv(rb, int(peer)))
        recvs.append((r, rb))
    for q in reqs:
        q.wait()
    for r, rb in recvs:
        v[r] = rb.numpy()
    return v.ravel()


def slab_model(nls_FEMModel, shape, dim, nranks, rank, device, unique_id):
    """Local FEMModel of one y/z-slab of a structured grid, wired for halo exchange."""
    part = slab_local_mesh(shape, dim, nranks, rank)
    fem = nls_FEMModel(part["mesh"], device=device)
    if nranks > 1:
        fem.set_partition(part["n_owned"], part["neigh"], part["send_ptr"], part["send_nodes"],
                          part["recv_ptr"], part["recv_nodes"])
        fem.set_comm(rank, nranks, unique_id)
    return fem, part
