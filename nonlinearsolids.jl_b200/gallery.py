"""Host-side mirror of src/gallery/residual.jl and jacobian.jl: the functors NewtonSolver calls.

Calling them runs the element loop on the GPU through the C ABI with host buffers, with the
reference's in/out contract: U comes back with its Dirichlet entries overwritten.
"""
import numpy as np

from ._lib import dptr


class CSRMatrix:
    """Jacobian storage: CSR values over the model's precomputed pattern (replaces Matrix{Float64})."""

    def __init__(self, fem):
        self.rowptr, self.colind = fem.pattern()
        self.vals = np.zeros(len(self.colind))
        self.shape = (len(self.rowptr) - 1, fem.ndof)

    def toarray(self):
        A = np.zeros(self.shape)
        for i in range(self.shape[0]):
            s, e = self.rowptr[i], self.rowptr[i + 1]
            A[i, self.colind[s:e]] = self.vals[s:e]
        return A


class DistributedForce:
    """DistributedForce(fdist_fn) (distributedforce.jl:3-27): body-force density, a number (or a vector
    with one entry per spatial dimension) or a function of time `fdist_fn(time(ctx))`. The element
    integrals run on the device; the host only evaluates the function of time."""

    def __init__(self, fdist_fn):
        self.fdist_fn = fdist_fn

    def density(self, ctx, dim):
        f = self.fdist_fn(ctx.time()) if callable(self.fdist_fn) else self.fdist_fn
        f = np.atleast_1d(np.asarray(f, dtype=np.float64))
        if len(f) != dim:
            raise ValueError("DistributedForce needs one density value per spatial dimension")
        return f


def compute_internalforce(material, fem, U, Fint, ctx=None, mode="add"):
    """compute_internalforce(material, fem, U, Fint, ctx; mode) (defaults.jl:29-46): the element loop alone. U is not
    modified (no Dirichlet overwrite) and no external force enters; `mode` is "add" or "set", anything else raises
    like the reference's ArgumentError (defaults.jl:30-32)."""
    if mode not in ("add", "set"):
        raise ValueError("Invalid mode: %s" % mode)
    tmp = np.zeros(fem.nrows)
    fem.handle(material).call("nls_internal_force", dptr(np.ascontiguousarray(U, dtype=np.float64)), dptr(tmp))
    if mode == "set":
        Fint[:] = tmp
    else:
        Fint += tmp
    return Fint


def compute_dinternalforce(material, fem, U, K, ctx=None, mode="add"):
    """compute_dinternalforce(material, fem, U, K, ctx; mode) (defaults.jl:68-82) into the CSR values of K."""
    if mode not in ("add", "set"):
        raise ValueError("Invalid mode: %s" % mode)
    tmp = np.zeros(len(K.vals))
    fem.handle(material).call("nls_internal_force_jacobian", dptr(np.ascontiguousarray(U, dtype=np.float64)), dptr(tmp))
    if mode == "set":
        K.vals[:] = tmp
    else:
        K.vals += tmp
    return K


def _sync_external_forces(H, fem, forces, ctx):
    total = np.zeros(fem.mesh.dim)
    for f in forces:
        if not isinstance(f, DistributedForce):
            raise NotImplementedError("only DistributedForce external forces run on the device path")
        total += f.density(ctx, fem.mesh.dim)
    H.call("nls_set_body_force", len(total) if len(forces) else 0, dptr(total))


class Residual:
    """Residual(material, externalforces=[]) (residual.jl:5-9). R = F_int(U) - F_ext."""

    def __init__(self, material, externalforces=()):
        for f in externalforces:
            if not isinstance(f, DistributedForce):
                raise NotImplementedError("only DistributedForce external forces run on the device path")
        self.material = material
        self.externalforces = list(externalforces)

    def __call__(self, fem, U, *args):
        """(fem, U, R, ctx) static (residual.jl:12-26); (fem, U, Ud, R, ctx) rate form (:29-45, no shipped material
        is velocity dependent, so it equals the static one); (fem, U, Ud, Udd, R, ctx) dynamic (:48-54)."""
        if not (isinstance(U, np.ndarray) and U.dtype == np.float64 and U.flags.c_contiguous):
            raise TypeError("U must be a contiguous float64 array (it is updated in place)")
        arrs = [a for a in args if isinstance(a, np.ndarray)]
        ctx = args[len(arrs)] if len(args) > len(arrs) else None
        if len(arrs) not in (1, 2, 3):
            raise TypeError("Residual(fem, U, [Ud, [Udd,]] R, ctx)")
        H = fem.handle(self.material)
        _sync_external_forces(H, fem, self.externalforces, ctx)
        if len(arrs) == 3:
            _ensure_mass(H, fem, self.material)
            H.call("nls_dynamic_residual", dptr(U), dptr(np.ascontiguousarray(arrs[1], dtype=np.float64)), dptr(arrs[2]))
        else:
            H.call("nls_residual", dptr(U), dptr(arrs[-1]))


class Jacobian:
    """Jacobian(material, dexternalforces=[]) (jacobian.jl:3-7). K = dF_int/dU in CSR."""

    def __init__(self, material, dexternalforces=()):
        if len(dexternalforces):
            raise NotImplementedError("external force tangents are not on the device path yet")
        self.material = material
        self.dexternalforces = list(dexternalforces)

    def __call__(self, fem, U, *args):
        """(fem, U, K, ctx) static (jacobian.jl:10-22); (fem, U, Ud, K, ctx) rate form (:25-38); (fem, U, Ud, Udd, K, ctx)
        dynamic (:42-51): K(U) * dUdd_dU(ctx) + M, ctx a NewmarkTimeStepper."""
        if not (isinstance(U, np.ndarray) and U.dtype == np.float64 and U.flags.c_contiguous):
            raise TypeError("U must be a contiguous float64 array (it is updated in place)")
        mats = [a for a in args if isinstance(a, (np.ndarray, CSRMatrix))]
        ctx = args[len(mats)] if len(args) > len(mats) else None
        if len(mats) not in (1, 2, 3):
            raise TypeError("Jacobian(fem, U, [Ud, [Udd,]] K, ctx)")
        K = mats[-1]
        vals = K.vals if isinstance(K, CSRMatrix) else K
        H = fem.handle(self.material)
        if len(mats) == 3:
            if ctx is None or not hasattr(ctx, "dUdd_dU"):
                raise TypeError("the dynamic Jacobian needs a timestepper context (jacobian.jl:42)")
            _ensure_mass(H, fem, self.material)
            H.call("nls_dynamic_jacobian", dptr(U), float(ctx.dUdd_dU()), dptr(vals))
        else:
            H.call("nls_jacobian", dptr(U), dptr(vals))


def _ensure_mass(H, fem, material):
    key = (float(material.density()), float(fem.data.get("A", 1.0)))     # mass.jl:9,27: A defaults to 1.0 here
    if getattr(H, "_mass_key", None) != key:
        H.call("nls_set_mass", key[0], key[1])
        H._mass_key = key


def assemblemass(material, fem, K):
    """assemblemass!(material, fem, K) (mass.jl:8-17) for a zeroed K: K = M in CSR."""
    H = fem.handle(material)
    _ensure_mass(H, fem, material)
    H.call("nls_get_mass", dptr(K.vals if isinstance(K, CSRMatrix) else K))


def applymass(material, fem, Udd, R):
    """applymass!(material, fem, U̇̇, R) (mass.jl:26-37): R += M U̇̇."""
    H = fem.handle(material)
    _ensure_mass(H, fem, material)
    H.call("nls_apply_mass", dptr(np.ascontiguousarray(Udd, dtype=np.float64)), dptr(R))


def residual_jacobian(material, fem, U, R, K):
    """Both operators from one element pass (same results as Residual then Jacobian)."""
    vals = K.vals if isinstance(K, CSRMatrix) else K
    fem.handle(material).call("nls_residual_jacobian", dptr(U), dptr(R), dptr(vals))
