"""Structured synthetic meshes and their slab partitions (8e). Host integer work."""
import numpy as np

from .fem import Mesh


def bar_mesh(nel, p=1, length=1.0):
    """1-D bar of `nel` order-p elements on [0, length]; element nodes ascending."""
    nn = nel * p + 1
    coords = np.linspace(0.0, length, nn)
    conn = (np.arange(nel, dtype=np.int32)[:, None] * p + np.arange(p + 1, dtype=np.int32)[None, :])
    return Mesh(1, conn, coords=coords)


def _shape(n, dim):
    """(n, dim) -> per-axis node counts (x, y[, z]); a tuple is taken as is (slowest axis last)."""
    return tuple(int(v) for v in n) if np.ndim(n) else (int(n),) * dim


def _grid_conn(n, dim, k0=0, k1=None):
    """Q1 connectivity of a structured grid (x fastest), element layers k0..k1 of the slowest axis."""
    sh = _shape(n, dim)
    k1 = (sh[-1] - 1) if k1 is None else k1
    nx = sh[0]
    if dim == 2:
        ei, ej = np.meshgrid(np.arange(nx - 1), np.arange(k0, k1), indexing="xy")
        base = (ei + nx * ej).ravel().astype(np.int64)
        offs = np.array([0, 1, nx, nx + 1], dtype=np.int64)
    else:
        ny = sh[1]
        ek, ej, ei = np.meshgrid(np.arange(k0, k1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
        base = (ei + nx * (ej + ny * ek)).ravel().astype(np.int64)
        pl = nx * ny
        offs = np.array([0, 1, nx, nx + 1, pl, pl + 1, pl + nx, pl + nx + 1], dtype=np.int64)
    return base[:, None] + offs[None, :]


def grid_coords(n, dim, k0=0, k1=None):
    """Coordinates i/(nx-1) (exactly) of node planes k0..k1 (exclusive) of the slowest axis."""
    sh = _shape(n, dim)
    k1 = sh[-1] if k1 is None else k1
    ax = [np.arange(m) / (sh[0] - 1) for m in sh[:-1]] + [np.arange(k0, k1) / (sh[0] - 1)]
    if dim == 2:
        X, Y = np.meshgrid(ax[0], ax[1], indexing="xy")
        return np.stack([X.ravel(), Y.ravel()], axis=1)
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def grid_mesh(n, dim):
    """Structured Q1 mesh, lexicographic numbering (x fastest). n: nodes per side, or a tuple of
    per-axis node counts (slowest axis last); n-per-side gives the unit square / cube."""
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
    """Node ids on the face `axis` = const (side 0: low, 1: high) of the structured grid."""
    sh = _shape(n, dim)
    idx = np.arange(int(np.prod(sh))).reshape(sh[::-1])   # [.., y, x]: axis d is array axis dim-1-d
    sl = [slice(None)] * dim
    sl[dim - 1 - axis] = 0 if side == 0 else sh[axis] - 1
    return np.sort(idx[tuple(sl)].ravel())


def slab_planes(n, nranks):
    """Owned node-plane ranges [floor(r n/P), floor((r+1) n/P)) of the slowest axis."""
    return [(r * n) // nranks for r in range(nranks + 1)]


def slab_local_mesh(n, dim, nranks, rank):
    """Closed-form local slab of a structured grid for `rank`: local mesh (owned nodes first, then
    the ghost planes below/above), halo lists, and the local->global node map. Matches
    nls_partition applied to the global mesh (tests check this bit-exactly on small grids)."""
    sh = _shape(n, dim)
    ns = sh[-1]
    cuts = slab_planes(ns, nranks)
    p0, p1 = cuts[rank], cuts[rank + 1]
    plane = int(np.prod(sh[:-1]))
    has_lo, has_hi = p0 > 0 and p1 > p0, p1 < ns and p1 > p0
    g0, g1 = p0 - (1 if has_lo else 0), p1 + (1 if has_hi else 0)   # node planes present locally
    n_owned = (p1 - p0) * plane
    owned = np.arange(p0 * plane, p1 * plane, dtype=np.int64)
    ghosts = []
    if has_lo:
        ghosts.append(np.arange((p0 - 1) * plane, p0 * plane, dtype=np.int64))
    if has_hi:
        ghosts.append(np.arange(p1 * plane, (p1 + 1) * plane, dtype=np.int64))
    l2g = np.concatenate([owned] + ghosts) if ghosts else owned
    conn_g = _grid_conn(sh, dim, g0, g1 - 1)          # elements touching an owned node

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
    coords = _coords_of(l2g, sh, dim)
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


def _coords_of(gids, sh, dim):
    """Coordinates of global node ids on a structured grid with spacing 1/(nx-1)."""
    g = np.asarray(gids, dtype=np.int64)
    out = np.zeros((len(g), dim))
    for d in range(dim):
        out[:, d] = (g % sh[d]) / (sh[0] - 1)
        g = g // sh[d]
    return out
