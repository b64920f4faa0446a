"""Structured synthetic meshes and their slab partitions (8e). Host integer work."""
import numpy as np

from .fem import Mesh


def bar_mesh(nel, p=1, length=1.0):
    """1-D bar of `nel` order-p elements on [0, length]; element nodes ascending."""
    nn = nel * p + 1
    coords = np.linspace(0.0, length, nn)
    conn = (np.arange(nel, dtype=np.int32)[:, None] * p + np.arange(p + 1, dtype=np.int32)[None, :])
    return Mesh(1, conn, coords=coords)


def _grid_conn(n, dim, k0=0, k1=None):
    """Q1 connectivity of an n^dim-node grid (x fastest), element layers k0..k1 of the slowest axis."""
    k1 = (n - 1) if k1 is None else k1
    if dim == 2:
        ei, ej = np.meshgrid(np.arange(n - 1), np.arange(k0, k1), indexing="xy")
        base = (ei + n * ej).ravel().astype(np.int64)
        offs = np.array([0, 1, n, n + 1], dtype=np.int64)
    else:
        ek, ej, ei = np.meshgrid(np.arange(k0, k1), np.arange(n - 1), np.arange(n - 1), indexing="ij")
        base = (ei + n * (ej + n * ek)).ravel().astype(np.int64)
        offs = np.array([0, 1, n, n + 1, n * n, n * n + 1, n * n + n, n * n + n + 1], dtype=np.int64)
    return base[:, None] + offs[None, :]


def grid_coords(n, dim, k0=0, k1=None):
    """Coordinates i/(n-1) of node planes k0..k1 (exclusive) of the slowest axis."""
    k1 = n if k1 is None else k1
    x = np.arange(n) / (n - 1)
    s = np.arange(k0, k1) / (n - 1)
    if dim == 2:
        X, Y = np.meshgrid(x, s, indexing="xy")
        return np.stack([X.ravel(), Y.ravel()], axis=1)
    Z, Y, X = np.meshgrid(s, x, x, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def grid_mesh(n, dim):
    """Unit square / cube, n nodes per side, Q1 elements, lexicographic numbering (x fastest)."""
    return Mesh(dim, _grid_conn(n, dim).astype(np.int32), coords=grid_coords(n, dim))


def smooth_field(coords, amp=0.01):
    """u_c(X) = amp*sin(2 pi X_c) * prod_{d != c} cos(pi X_d), interleaved DoFs."""
    X = np.asarray(coords)
    dim = X.shape[1]
    U = np.zeros_like(X)
    for c in range(dim):
        v = amp * np.sin(2 * np.pi * X[:, c])
        for d in range(dim):
            if d != c:
                v = v * np.cos(np.pi * X[:, d])
        U[:, c] = v
    return U.ravel()


def face_nodes(n, dim, axis, side):
    """Node ids on the face `axis` = 0 (side 0) or 1 (side 1) of the unit grid."""
    idx = np.arange(n ** dim).reshape((n,) * dim)   # [.., y, x]: axis d is array axis dim-1-d
    sl = [slice(None)] * dim
    sl[dim - 1 - axis] = 0 if side == 0 else n - 1
    nodes = np.sort(idx[tuple(sl)].ravel())
    return nodes


def slab_planes(n, nranks):
    """Owned node-plane ranges [floor(r n/P), floor((r+1) n/P)) of the slowest axis."""
    return [(r * n) // nranks for r in range(nranks + 1)]


def slab_local_mesh(n, dim, nranks, rank):
    """Closed-form local slab of the unit grid for `rank`: local mesh (owned nodes first, then the
    ghost planes below/above), halo lists, and the local->global node map. Matches nls_partition
    applied to the global mesh (tests check this bit-exactly on small grids)."""
    cuts = slab_planes(n, nranks)
    p0, p1 = cuts[rank], cuts[rank + 1]
    plane = n ** (dim - 1)
    has_lo, has_hi = p0 > 0 and p1 > p0, p1 < n and p1 > p0
    g0, g1 = p0 - (1 if has_lo else 0), p1 + (1 if has_hi else 0)   # node planes present locally
    n_owned = (p1 - p0) * plane
    # global node ids present, in local order: owned, then ghosts ascending (low plane, high plane)
    owned = np.arange(p0 * plane, p1 * plane, dtype=np.int64)
    ghosts = []
    if has_lo:
        ghosts.append(np.arange((p0 - 1) * plane, p0 * plane, dtype=np.int64))
    if has_hi:
        ghosts.append(np.arange(p1 * plane, (p1 + 1) * plane, dtype=np.int64))
    l2g = np.concatenate([owned] + ghosts) if ghosts else owned
    # elements touching an owned node: layers g0..g1-1
    conn_g = _grid_conn(n, dim, g0, g1 - 1)
    g2l = {}
    # vectorised global->local: owned block, low ghost block, high ghost block
    def to_local(g):
        out = np.empty_like(g)
        m_own = (g >= p0 * plane) & (g < p1 * plane)
        out[m_own] = g[m_own] - p0 * plane
        m_lo = g < p0 * plane
        out[m_lo] = n_owned + (g[m_lo] - (p0 - 1) * plane)
        m_hi = g >= p1 * plane
        out[m_hi] = n_owned + (plane if has_lo else 0) + (g[m_hi] - p1 * plane)
        return out
    conn_l = to_local(conn_g).astype(np.int32)
    coords = grid_coords(n, dim)[l2g] if n ** dim <= 2_000_000 else _coords_of(l2g, n, dim)
    neigh, send_ptr, send_nodes, recv_ptr, recv_nodes = [], [0], [], [0], []
    if has_lo:   # neighbour rank-1: send my lowest owned plane, receive its highest
        neigh.append(rank - 1)
        send_nodes.append(np.arange(0, plane, dtype=np.int32))
        recv_nodes.append(n_owned + np.arange(0, plane, dtype=np.int32))
        send_ptr.append(send_ptr[-1] + plane); recv_ptr.append(recv_ptr[-1] + plane)
    if has_hi:
        neigh.append(rank + 1)
        send_nodes.append(np.arange(n_owned - plane, n_owned, dtype=np.int32))
        recv_nodes.append(n_owned + (plane if has_lo else 0) + np.arange(0, plane, dtype=np.int32))
        send_ptr.append(send_ptr[-1] + plane); recv_ptr.append(recv_ptr[-1] + plane)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int32)
    return dict(mesh=Mesh(dim, conn_l, coords=coords), n_owned=n_owned, l2g=l2g,
                neigh=np.array(neigh, dtype=np.int32), send_ptr=np.array(send_ptr, dtype=np.int64),
                send_nodes=cat(send_nodes), recv_ptr=np.array(recv_ptr, dtype=np.int64),
                recv_nodes=cat(recv_nodes), elem_layers=(g0, g1 - 1), node_off=[c * plane for c in cuts])


def _coords_of(gids, n, dim):
    g = np.asarray(gids, dtype=np.int64)
    out = np.zeros((len(g), dim))
    for d in range(dim):
        out[:, d] = (g % n) / (n - 1)
        g = g // n
    return out
