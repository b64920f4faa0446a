"""Host-side mirror of src/nonlinearsolvers/{types,newton_fem}.jl and src/timesteppers/{types,pseudotime}.jl.

`NewtonSolver.solve` hands the whole Newton loop to the GPU library (nls_newton_solve): assembly,
Jacobi-PCG (in place of the reference's dense `\\`), update and norms stay on the device; the
loop semantics (order of the rtol/atol tests, strict <, norm_r0 rule, dtol, the maxits quirk,
:modified) are those of newton_fem.jl:132-190.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import dptr
from .gallery import Jacobian, Residual, _ensure_mass, _sync_external_forces
from .materials import save_state


# ---- ConvergedReasons (types.jl:1-38) ---------------------------------------------------
class _Reason:
    def __init__(self, text, converged, code):
        self.repr, self.converged, self.code = text, converged, code

    def __repr__(self):
        return self.repr


REASON_NOT_YET_CONVERGED = _Reason("Not yet converged", None, 0)
REASON_CONVERGED_RELATIVE = _Reason("Converged: relative tolerance", True, 1)
REASON_CONVERGED_ABSOLUTE = _Reason("Converged: absolute tolerance", True, 2)
REASON_DIVERGED_MAXITS = _Reason("Diverged: maximum iterations", False, 3)
REASON_DIVERGED_STEP = _Reason("Diverged: step tolerance", False, 4)
REASON_DIVERGED_LINEAR_SOLVE = _Reason("Diverged: linear solve failed", False, 5)
_REASONS = [REASON_NOT_YET_CONVERGED, REASON_CONVERGED_RELATIVE, REASON_CONVERGED_ABSOLUTE,
            REASON_DIVERGED_MAXITS, REASON_DIVERGED_STEP, REASON_DIVERGED_LINEAR_SOLVE]


def is_converged(x, allow_maxits=None):
    """is_converged(reason | solver; allow_maxits) (types.jl:35-37, newton_fem.jl:123)."""
    if isinstance(x, NewtonSolver):
        return is_converged(x.converged_reason, x.allow_maxits if allow_maxits is None else allow_maxits)
    if x.converged is True:
        return True
    if x.converged is False:
        return x is REASON_DIVERGED_MAXITS and bool(allow_maxits)
    return False


class NewtonSolver:
    """NewtonSolver(fem; type, atol, rtol, dtol, maxits, ...) (newton_fem.jl:5-53).

    Extra keywords `lin_rtol`, `lin_maxits` control the Krylov solve that replaces `\\`."""

    def __init__(self, fem, type="standard", atol=1e-16, rtol=1e-16, dtol=1e6, maxits=100, ctx=None,
                 monitor_residual_norm=False, monitor_converged_reason=False, allow_maxits=False,
                 lin_rtol=1e-12, lin_maxits=100000):
        if type not in ("standard", "modified"):
            raise ValueError("type must be 'standard' or 'modified'")
        self.fem, self.type = fem, type
        self.atol, self.rtol, self.dtol, self.maxits = atol, rtol, dtol, maxits
        self.ctx = ctx
        self.monitor_residual_norm = monitor_residual_norm
        self.monitor_converged_reason = monitor_converged_reason
        self.allow_maxits = allow_maxits
        self.lin_rtol, self.lin_maxits = lin_rtol, lin_maxits
        self.residual_fn = None
        self.jacobian_fn = None
        self.U = np.zeros(fem.ndof)
        self.dU = np.zeros(fem.ndof)
        self.R = np.zeros(fem.nrows)
        self.converged_reason = REASON_NOT_YET_CONVERGED
        self.iterations = 0
        self.norm_history = np.zeros(0)
        self.lin_iterations = 0
        self.timing_ms = {}

    # accessors (newton_fem.jl:74-86)
    def solution(self):
        return self.U

    def residual(self):
        return self.R

    def model(self):
        return self.fem

    def numdof(self):
        return self.fem.ndof

    @property
    def material(self):
        return getattr(self.residual_fn, "material", None)

    def solve(self, U0=None):
        """solve!(solver, U0) (newton_fem.jl:132-190)."""
        if not isinstance(self.residual_fn, Residual) or not isinstance(self.jacobian_fn, Jacobian):
            raise RuntimeError("set_residual_fn/set_jacobian_fn with gallery Residual/Jacobian functors first")
        if self.jacobian_fn.material is not self.residual_fn.material:
            raise RuntimeError("Residual and Jacobian must share one material")
        fem = self.fem
        U0 = np.zeros(fem.ndof) if U0 is None else np.ascontiguousarray(U0, dtype=np.float64)
        if U0.shape != (fem.ndof,):
            raise ValueError("U0 has the wrong length")
        H = fem.handle(self.residual_fn.material)
        _sync_external_forces(H, fem, self.residual_fn.externalforces, self.ctx)   # fdist_fn(time(ctx))
        opts = _lib.NewtonOpts(int(self.type == "modified"), self.atol, self.rtol, self.dtol, self.maxits,
                               self.lin_rtol, self.lin_maxits)
        res = _lib.NewtonResult()
        hist = np.zeros(self.maxits + 2)
        H.call("nls_newton_solve", C.byref(opts), dptr(U0), dptr(self.U), dptr(self.R), dptr(hist), C.byref(res))
        self.converged_reason = _REASONS[res.reason]
        self.iterations = res.iterations
        self.lin_iterations = res.lin_its_total
        self.norm_history = hist[:res.nhist].copy()
        self.timing_ms = dict(assembly=res.ms_assembly, linear=res.ms_linear, total=res.ms_total)
        if self.monitor_residual_norm:
            for i, v in enumerate(self.norm_history):
                print("  Iteration %d: residual norm: %.16g" % (i, v))
        if self.monitor_converged_reason:
            print(self.converged_reason, "iterations =", self.iterations)


def set_residual_fn(solver, fn):
    (solver.solver if isinstance(solver, (PseudoTimeStepper, NewmarkTimeStepper)) else solver).residual_fn = fn


def set_jacobian_fn(solver, fn):
    (solver.solver if isinstance(solver, (PseudoTimeStepper, NewmarkTimeStepper)) else solver).jacobian_fn = fn


def set_ctx(solver, ctx):
    solver.ctx = ctx


# ---- timesteppers (types.jl, pseudotime.jl) ------------------------------------------------
class PseudoTimeStepper:
    """PseudoTimeStepper(solver; Δt, maxtime, ...) (pseudotime.jl:4-37): quasi-static load stepping."""

    def __init__(self, solver, dt=0.01, t=0.0, maxtime=1.0, maxsteps=None, allow_converged_max_its=False):
        self.solver = solver
        self.t, self.step, self.dt, self.t0, self.maxtime = float(t), 0, float(dt), float(t), float(maxtime)
        if maxsteps is None:
            ms = (self.maxtime - self.t) / self.dt + 1
            if ms != int(ms):
                raise ValueError("InexactError: Int(%r)" % ms)   # pseudotime.jl:16 uses Int(...)
            maxsteps = int(ms)
        self.maxsteps = maxsteps
        self.allow_converged_max_its = allow_converged_max_its
        self.U = np.zeros((self.maxsteps, solver.fem.ndof))
        self.postprocess = []
        self.output_data = {}

    # verbs (timesteppers/types.jl:5-14)
    def time(self):
        return self.t

    def model(self):
        return self.solver.fem

    def numdof(self):
        return self.solver.fem.ndof

    @property
    def material(self):
        return self.solver.material

    def getoutputarray(self, key, dims=None):
        if key not in self.output_data:
            shape = (self.maxsteps, self.solver.fem.nrows) if dims is None else \
                (self.maxsteps,) + (tuple(dims) if np.ndim(dims) else (int(dims),))
            self.output_data[key] = np.zeros(shape)
        return self.output_data[key]

    def __getattr__(self, key):   # output arrays surface as pseudo-properties (types.jl:31-39)
        od = self.__dict__.get("output_data", {})
        if key in od:
            return od[key]
        raise AttributeError(key)

    def trim(self):
        for k, v in self.output_data.items():
            self.output_data[k] = v[:self.step + 1]

    def addpostprocess(self, f):
        self.postprocess.append(f)

    def step_(self, dt):
        """step!(ts, Δt) (types.jl:68-72)."""
        self.t += dt
        self.step += 1
        self.solver.fem.updateboundaries(self.t)

    def poststep(self, u):
        """poststep!(ts, u) (types.jl:74-78): commit the material state of a converged step."""
        mat = self.material
        if mat is not None and mat.has_state:
            save_state(mat, self.solver.fem, u, self)

    def solution(self):
        return self.U

    def solve(self, U=None):
        """solve!(ts, U0) (pseudotime.jl:73-93). Steps are 0-based here: row 0 holds the initial state."""
        fem = self.solver.fem
        U = np.zeros(fem.ndof) if U is None else np.asarray(U, dtype=np.float64)
        r = self.getoutputarray("residual")
        num_its = self.getoutputarray("num_its", 1)
        self.U[self.step, :] = U
        self.t0 = self.t
        num_its[self.step] = 0
        while self.t < self.maxtime and self.step < self.maxsteps - 1:
            self.step_(self.dt)
            self.solver.solve(self.U[self.step - 1, :])
            if not is_converged(self.solver, self.allow_converged_max_its or self.solver.allow_maxits):
                self.trim()
                return
            self.U[self.step, :] = self.solver.solution()
            r[self.step, :] = self.solver.residual()
            num_its[self.step] = self.solver.iterations
            self.poststep(self.U[self.step, :])
            for f in self.postprocess:
                f(self, self.U[self.step, :])


class NewmarkTimeStepper(PseudoTimeStepper):
    """NewmarkTimeStepper(solver, material; beta, gamma, dt, maxtime) (timesteppers/newmark.jl:5-53): implicit
    dynamics. Newton runs on the acceleration: the closures of newmark.jl:75-107 (update_state! -> static
    Residual/Jacobian of U -> inertia terms M a and dt^2 beta K + M) and the Newton loop all stay on the device
    (nls_newmark_solve); the host keeps the histories U, Ud, Udd (maxsteps x ndof) like the reference.

    Kept from the reference: the solve is restricted to the free DoFs, so constrained entries of the acceleration
    keep their initial-guess values, and Dirichlet values only enter through the copy of U the operators act on
    (newmark.jl:78,103 apply them to a temporary). Two deviations from the text of solve! (newmark.jl:124-141),
    which as written discards its own results: U0 and Ud0 are kept (the reference's first residual call runs
    update_state! with zero predictors over them), and with consistent_a0 (default) the initial acceleration
    solves M a0 = -R(U0) -- the reference writes that residual into a temporary and so always starts from a0 = 0,
    which consistent_a0=False reproduces."""

    def __init__(self, solver, material, beta=0.25, gamma=0.5, dt=0.01, t=0.0, maxtime=1.0, maxsteps=None,
                 consistent_a0=True):
        ms = int(round((maxtime - t) / dt + 1)) if maxsteps is None else maxsteps     # newmark.jl:21 uses round
        super().__init__(solver, dt=dt, t=t, maxtime=maxtime, maxsteps=ms)
        self._material = material
        self.beta, self.gamma = float(beta), float(gamma)
        self.consistent_a0 = consistent_a0
        n = solver.fem.ndof
        self.Ut, self.Udt = np.zeros(n), np.zeros(n)
        self.Ud = np.zeros((self.maxsteps, n))
        self.Udd = np.zeros((self.maxsteps, n))

    @property
    def material(self):
        return self._material

    def dUdd_dU(self):
        """∂U̇̇∂U(ts) (newmark.jl:56-58)."""
        return self.dt ** 2 * self.beta

    def update_predictors(self, n=None):
        """update_predictors!(ts, n) (newmark.jl:109-112)."""
        n = self.step if n is None else n
        self.Ut[:] = self.U[n] + self.dt * self.Ud[n] + self.dt ** 2 * (0.5 - self.beta) * self.Udd[n]
        self.Udt[:] = self.Ud[n] + self.dt * (1 - self.gamma) * self.Udd[n]

    def velocity(self):
        return self.Ud

    def acceleration(self):
        return self.Udd

    def _handle(self):
        fem, mat = self.solver.fem, self._material
        H = fem.handle(mat)
        _ensure_mass(H, fem, mat)
        return H

    def solve(self, U0=None, Ud0=None, Udd0=None):
        """solve!(ts, U0, U̇0, U̇̇0) (newmark.jl:124-156). Steps are 0-based: row 0 holds the initial state."""
        sv = self.solver
        fem = sv.fem
        n = fem.ndof
        if not isinstance(sv.residual_fn, Residual) or not isinstance(sv.jacobian_fn, Jacobian):
            raise RuntimeError("set_residual_fn/set_jacobian_fn with gallery Residual/Jacobian functors first")
        H = self._handle()
        r = self.getoutputarray("residual")
        num_its = self.getoutputarray("num_its", 1)
        self.U[self.step] = 0.0 if U0 is None else U0
        self.Ud[self.step] = 0.0 if Ud0 is None else Ud0
        self.Udd[self.step] = 0.0 if Udd0 is None else Udd0
        self.t0 = self.t
        num_its[self.step] = 0
        if self.consistent_a0:      # M a0 = Fext(t0) - Fint(U0)  (newmark.jl:133-140)
            _sync_external_forces(H, fem, sv.residual_fn.externalforces, self)
            R0, Uc = np.zeros(fem.nrows), np.ascontiguousarray(self.U[self.step].copy())
            H.call("nls_residual", dptr(Uc), dptr(R0))
            a0, its, st = np.zeros(n), C.c_int32(0), C.c_int32(0)
            H.call("nls_mass_solve", 1e-14, 100000, dptr(np.ascontiguousarray(-R0)), dptr(a0), C.byref(its), C.byref(st))
            if st.value != 0:
                raise RuntimeError("mass solve for the initial acceleration did not converge")
            self.Udd[self.step] = a0
        opts = _lib.NewtonOpts(int(sv.type == "modified"), sv.atol, sv.rtol, sv.dtol, sv.maxits, sv.lin_rtol, sv.lin_maxits)
        while self.t < self.maxtime and self.step < self.maxsteps - 1:
            self.update_predictors(self.step)
            self.step_(self.dt)
            H = self._handle()                  # pushes the boundary values updateboundaries! just changed
            _sync_external_forces(H, fem, sv.residual_fn.externalforces, self)
            res = _lib.NewtonResult()
            hist = np.zeros(sv.maxits + 2)
            a, U, Ud, R = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(fem.nrows)
            H.call("nls_newmark_solve", C.byref(opts), self.beta, self.gamma, self.dt, dptr(self.Ut), dptr(self.Udt),
                   dptr(np.ascontiguousarray(self.Udd[self.step - 1])), dptr(a), dptr(U), dptr(Ud), dptr(R), dptr(hist), C.byref(res))
            sv.converged_reason = _REASONS[res.reason]
            sv.iterations, sv.lin_iterations = res.iterations, res.lin_its_total
            sv.norm_history = hist[:res.nhist].copy()
            sv.U[:], sv.R[:] = a, R
            if not is_converged(sv, sv.allow_maxits):
                self.trim()
                return
            self.Udd[self.step] = a
            r[self.step, :] = R
            num_its[self.step] = sv.iterations
            self.U[self.step], self.Ud[self.step] = U, Ud          # update_state!(ts, Udd) (newmark.jl:114-117)
            self.poststep(self.U[self.step, :])
            for f in self.postprocess:
                f(self, self.U[self.step, :])
