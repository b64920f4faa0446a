"""Host-side mirror of src/fem/ (element.jl, shapefunctions.jl, quadrature.jl, boundary.jl, fem.jl).

Same names and argument meaning as the reference's Julia API, with two Python
adaptations: indices are 0-based, and `f!` becomes `f` (no bang). Host objects
only describe the model; every operator on the hot path runs in libnls_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, dptr, iptr, lptr


# ---- shapefunctions.jl -----------------------------------------------------------------
class Lagrange:
    """Lagrange interpolating polynomials in barycentric form (shapefunctions.jl:8-11)."""

    def __init__(self, w, nodes):
        self.w = as_f64(w)
        self.nodes = as_f64(nodes)


def lagrange(order_or_points, t="chebyshev2"):
    """lagrange(points) / lagrange(order, :equidistant | :chebyshev2) (shapefunctions.jl:23-47)."""
    L = _lib.lib()
    if np.ndim(order_or_points) > 0:
        pts = as_f64(order_or_points)
        w = np.ones(len(pts))
        for i in range(len(pts)):
            w[i] = np.prod(pts[i] - pts[:i]) * np.prod(pts[i] - pts[i + 1:])
        return Lagrange(1.0 / w, pts)
    if t not in ("chebyshev2", "equidistant"):
        raise ValueError("Unknown node type %s" % t)
    p = int(order_or_points)
    nodes, w = np.zeros(p + 1), np.zeros(p + 1)
    if L.nls_lagrange(p, 0 if t == "chebyshev2" else 1, dptr(nodes), dptr(w)) != 0:
        raise ValueError("unsupported order %d" % p)
    return Lagrange(w, nodes)


def N(shape, xi):
    """Shape-function values at xi (shapefunctions.jl:50-61)."""
    out = np.zeros(len(shape.nodes))
    _lib.lib().nls_shape_eval(len(shape.nodes), dptr(shape.nodes), dptr(shape.w), float(xi), dptr(out), None)
    return out


def dN(shape, xi):
    """Shape-function derivatives at xi; the reference's `∇N` (shapefunctions.jl:64-81)."""
    out = np.zeros(len(shape.nodes))
    _lib.lib().nls_shape_eval(len(shape.nodes), dptr(shape.nodes), dptr(shape.w), float(xi), None, dptr(out))
    return out


def interpolate(shape, f, xi):
    vals = np.array([f(x) for x in shape.nodes]) if callable(f) else as_f64(f)
    return float(np.dot(vals, N(shape, xi)))


# ---- quadrature.jl ---------------------------------------------------------------------
class Quadrature:
    def __init__(self, points, weights):
        if len(points) != len(weights):
            raise ValueError("length(points) != length(weights)")
        self.points = as_f64(points)
        self.weights = as_f64(weights)


def gauss_quadrature(order):
    """Closed-form Gauss rules 1..5 (quadrature.jl:29-50)."""
    if not 1 <= order <= 5:
        raise ValueError("Gauss quadrature not implemented for n = %d" % order)
    x, w = np.zeros(order), np.zeros(order)
    _lib.lib().nls_gauss_quadrature(order, dptr(x), dptr(w))
    return Quadrature(x, w)


def integrate(q, f, mode="noindex"):
    """integrate(q, f[, mode]) (quadrature.jl:58-70): sum(f .* weights), q ascending."""
    if callable(f):
        if mode not in ("noindex", "index"):
            raise AssertionError("Mode must be :noindex or :index")
        f = [f(x, i) if mode == "index" else f(x) for i, x in enumerate(q.points)]
    acc = None
    for fq, wq in zip(f, q.weights):
        term = np.asarray(fq) * wq
        acc = term if acc is None else acc + term
    return acc


# ---- element.jl ------------------------------------------------------------------------
class Element:
    """Element(nodes, coords) (element.jl:6-12); nodes are 0-based here."""

    def __init__(self, nodes, coords):
        self.nodes = np.asarray(nodes, dtype=np.int64)
        self.x_loc = np.asarray(coords)[self.nodes]

    def __len__(self):
        return len(self.nodes)

    def length(self):
        """Length of the 1-D element (element.jl:15)."""
        return self.x_loc[-1] - self.x_loc[0]


class Mesh:
    """Mesh(dim, elements, nodes, coords) (element.jl:73-83).

    `elements` is a list of Element or a flattened [nel][nen] connectivity array.
    Multi-D meshes (the T2 extension) use Q1 elements in tensor order, x fastest.
    """

    def __init__(self, dim, elements, nodes=None, coords=None):
        if coords is None:
            raise ValueError("coords required")
        self.dim = int(dim)
        if len(elements) and isinstance(elements[0], Element):
            self.elements = list(elements)
            self.conn = np.ascontiguousarray([e.nodes for e in elements], dtype=np.int32)
        else:
            self.elements = None
            self.conn = np.ascontiguousarray(elements, dtype=np.int32)
        self.coords = as_f64(coords).reshape(-1, self.dim)
        self.nodes = np.arange(len(self.coords)) if nodes is None else np.asarray(nodes)

    @property
    def nel(self):
        return self.conn.shape[0]

    @property
    def nen(self):
        return self.conn.shape[1]

    @property
    def nn(self):
        return self.coords.shape[0]


# ---- boundary.jl -----------------------------------------------------------------------
class AbstractBoundary:
    def update(self, t, *a, **k):
        pass


class Dirichlet(AbstractBoundary):
    """Dirichlet(nodes, values) (boundary.jl:21-29); `nodes` are DoF indices (gdof = dim*node + c)."""
    isdirichlet = True

    def __init__(self, nodes, values):
        self.nodes = np.asarray(nodes, dtype=np.int64)
        self.values = np.array(values, dtype=np.float64)

    def apply(self, D):
        D[self.nodes] = self.values


class Neumann(AbstractBoundary):
    """Neumann(nodes, values) (boundary.jl:32-39)."""
    isdirichlet = False

    def __init__(self, nodes, values):
        self.nodes = np.asarray(nodes, dtype=np.int64)
        self.values = np.array(values, dtype=np.float64)

    def apply(self, F):
        np.add.at(F, self.nodes, self.values)


class ElementBoundary(AbstractBoundary):
    """Boundary condition restricted to one element (boundary.jl:42-61)."""

    def __init__(self, el, bc):
        if isinstance(bc, TimeDependent):
            bc = bc.bc
        idx_local = np.isin(el.nodes, bc.nodes)
        idx_vals = np.isin(bc.nodes, el.nodes)
        self.element = el
        self.bc_el = type(bc)(np.nonzero(idx_local)[0], bc.values[idx_vals])  # view semantics below
        self._parent, self._idx_vals = bc, idx_vals

    @property
    def isdirichlet(self):
        return self.bc_el.isdirichlet

    @property
    def nodes(self):
        return self.bc_el.nodes

    @property
    def values(self):
        return self._parent.values[self._idx_vals]  # @views in the reference: tracks updates

    def apply(self, evec):
        self.bc_el.values = self.values
        self.bc_el.apply(evec)


class TimeDependent(AbstractBoundary):
    """TimeDependent(bc, update_fn; pass_context=false) (boundary.jl:64-92)."""

    def __init__(self, bc, update_fn, pass_context=False):
        self.bc, self.update_fn, self.pass_context = bc, update_fn, pass_context

    @property
    def isdirichlet(self):
        return self.bc.isdirichlet

    @property
    def nodes(self):
        return self.bc.nodes

    @property
    def values(self):
        return self.bc.values

    def apply(self, v):
        self.bc.apply(v)

    def update(self, t, *a, **k):
        new = self.update_fn(self.bc, t, *a, **k) if self.pass_context else self.update_fn(t, *a, **k)
        self.bc.values[:] = new


# ---- fem.jl ----------------------------------------------------------------------------
class FEMModel:
    """FEMModel(mesh, q, p; data) (fem.jl:9-24) plus the device handle that runs its operators."""

    def __init__(self, mesh, q=2, p=1, data=None, device=0):
        self.mesh = mesh
        self.q, self.p = int(q), int(p)
        self.data = dict(data or {})
        self.boundaries = []
        self.ndof = mesh.nn * mesh.dim
        self.constrained_nodes = np.zeros(self.ndof, dtype=bool)
        self.device = device
        self._handle = None
        self._material = None
        self._bc_dirty = True
        self._n_owned = None
        self._halo = None
        self._comm = None

    # -- host bookkeeping, as in the reference
    def numdof(self):
        return self.ndof

    def addboundary(self, bc):
        self.boundaries.append(bc)
        if bc.isdirichlet:
            self.constrained_nodes[bc.nodes] = True
        self._bc_dirty = True

    def updateboundaries(self, t):
        for bc in self.boundaries:
            bc.update(t)
        self._bc_dirty = True

    def applydirichletboundaries(self, v):
        for bc in self.boundaries:
            if bc.isdirichlet:
                bc.apply(v)

    def applyneumannboundaries(self, v):
        for bc in self.boundaries:
            if not bc.isdirichlet:
                bc.apply(v)

    def dofs(self, v):
        v = np.asarray(v)
        free = ~self.constrained_nodes
        return v[free] if v.ndim == 1 else v[np.ix_(free, free)]

    def expand(self, v):
        out = np.zeros(self.ndof)
        out[~self.constrained_nodes] = v
        return out

    # -- device side
    def set_partition(self, n_owned_nodes, neigh_ranks, send_ptr, send_nodes, recv_ptr, recv_nodes):
        """Declare this model as one slab of a partitioned mesh (local numbering: owned, then ghosts)."""
        self._n_owned = int(n_owned_nodes)
        self._halo = (np.ascontiguousarray(neigh_ranks, dtype=np.int32), np.ascontiguousarray(send_ptr, dtype=np.int64),
                      np.ascontiguousarray(send_nodes, dtype=np.int32), np.ascontiguousarray(recv_ptr, dtype=np.int64),
                      np.ascontiguousarray(recv_nodes, dtype=np.int32))
        self._handle = None

    def set_comm(self, rank, nranks, unique_id):
        self._comm = (int(rank), int(nranks), bytes(unique_id))
        self._handle = None

    @property
    def nrows(self):
        return (self._n_owned if self._n_owned is not None else self.mesh.nn) * self.mesh.dim

    def handle(self, material=None):
        """Build (once) the device model: mesh, tables, pattern; then sync material and BCs."""
        if self._handle is None:
            H = _lib.Handle(self.device)
            m = self.mesh
            if self._comm is not None:
                rank, nranks, uid = self._comm
                idbuf = (C.c_uint8 * 128).from_buffer_copy(uid)
                H.call("nls_comm_init", _lib.nccl_lib_path().encode(), rank, nranks, idbuf)
            H.call("nls_set_mesh", m.dim, m.nn, m.nel, m.nen, iptr(m.conn), dptr(m.coords))
            H.call("nls_set_discretization", self.p if m.dim == 1 else 1, self.q if m.dim == 1 else 2)
            if self._halo is not None:
                nr, sp, sn, rp, rn = self._halo
                H.call("nls_set_halo", self._n_owned, len(nr), iptr(nr), lptr(sp), iptr(sn), lptr(rp), iptr(rn))
            H.call("nls_build_pattern")
            self._handle = H
            self._material = None
            self._bc_dirty = True
        H = self._handle
        if material is not None and material is not self._material:
            prm = as_f64(material.params())
            H.call("nls_set_material", material.kind, dptr(prm), len(prm))
            H.call("nls_set_area", float(self.data.get("A", 3e-4)))   # get(fem.data, :A, 3e-4)
            self._material = material
        if self._bc_dirty:
            dn = [bc for bc in self.boundaries if bc.isdirichlet]
            nn_ = [bc for bc in self.boundaries if not bc.isdirichlet]
            dd = np.concatenate([bc.nodes for bc in dn]) if dn else np.zeros(0, dtype=np.int64)
            dv = np.concatenate([bc.values for bc in dn]) if dn else np.zeros(0)
            nd = np.concatenate([bc.nodes for bc in nn_]) if nn_ else np.zeros(0, dtype=np.int64)
            nv = np.concatenate([bc.values for bc in nn_]) if nn_ else np.zeros(0)
            dd, nd = np.ascontiguousarray(dd, dtype=np.int64), np.ascontiguousarray(nd, dtype=np.int64)
            H.call("nls_set_dirichlet", len(dd), lptr(dd), dptr(as_f64(dv)))
            H.call("nls_set_neumann", len(nd), lptr(nd), dptr(as_f64(nv)))
            self._bc_dirty = False
        return H

    def pattern(self):
        """CSR pattern (rowptr int64, colind int32) of the Jacobian."""
        H = self.handle()
        nrows, nnz = C.c_int64(0), C.c_int64(0)
        H.call("nls_pattern_size", C.byref(nrows), C.byref(nnz))
        rowptr = np.zeros(nrows.value + 1, dtype=np.int64)
        colind = np.zeros(nnz.value, dtype=np.int32)
        H.call("nls_get_pattern", lptr(rowptr), iptr(colind))
        return rowptr, colind


def getstate(fem, name):
    """Per-(element, qp) state array of a stateful material (getqdata_el!, fem.jl:51-56)."""
    H = fem.handle()
    nq = fem.q if fem.mesh.dim == 1 else 2 ** fem.mesh.dim
    out = np.zeros((fem.mesh.nel, nq))
    H.call("nls_get_state", name.encode(), dptr(out))
    return out
