"""Host-side mirror of src/nonlinearsolvers/newtonarclength_fem.jl (+ the ArcLength functor of newtonarclength.jl:6-22)
and arclengthresults.jl: Newton-Raphson with arc-length continuation through limit points.

The continuation logic (predictors, the arc-length constraint, the bordered update) is host-side scalar and vector
algebra exactly as in the reference; what runs on the GPU is what costs: the internal force, the tangent assembly, and
the linear solves. The reference solves the bordered dense system `[K -Fext; df/dd df/dlambda] \\ r`
(newtonarclength_fem.jl:103-104); here it is the equivalent block elimination with two solves on the same CSR tangent,
    K v1 = r_d,  K v2 = F_ext,  dlambda = (r_l - g_d.v1) / (g_l + g_d.v2),  dd = v1 + dlambda v2,
each a Jacobi-preconditioned MINRES on the device (nls_linear_solve_indefinite): past a limit point the tangent is
indefinite, so PCG does not apply.

Deviations from the text of the reference, which is not runnable as shipped: calls to functions that do
not exist (`updateboundaries!(fem)` with one argument, `step!(fem, dlambda)`, `postprocess!(fem, d, res)`:
newtonarclength_fem.jl:49,155-156) are dropped; the sign of the first load increments follows the sign of
`dtilde . F_ext` (the current stiffness parameter) instead of the sign of det(K) (:64-70), which a Krylov solver does not
have -- both change sign together at a simple limit point (tests compare the two on a snap-through arch).
"""
import ctypes as C

import numpy as np

from ._lib import dptr
from .gallery import CSRMatrix


class ArcLengthResult:
    """ArcLengthResult(fem; maxsteps, maxits) (arclengthresults.jl:4-40), trimmed to the steps taken."""

    def __init__(self, ndof, maxsteps, maxits):
        self.d = np.zeros((maxsteps, ndof))
        self.lam = np.zeros(maxsteps)
        self.num_its = np.zeros(maxsteps, dtype=int)
        self.res_d = np.zeros((maxsteps, maxits + 1))
        self.res_lam = np.zeros((maxsteps, maxits + 1))
        self.num_steps = 1
        self.lin_iterations = 0
        self.reason = "completed"

    def trim(self):
        n = self.num_steps
        self.d, self.lam, self.num_its = self.d[:n], self.lam[:n], self.num_its[:n]
        self.res_d, self.res_lam = self.res_d[:n], self.res_lam[:n]
        return self


class ArcLength:
    """Arc length functor f(dd, dl) = sqrt((1-b) dd'Kdiag dd / den + b dl^2) and its gradient (newtonarclength.jl:6-22)."""

    def __init__(self, kdiag, dtilde, b=0.5):
        self.kdiag, self.b = kdiag, b
        self.den = float(dtilde @ (kdiag * dtilde))

    def __call__(self, dd, dl):
        return float(np.sqrt((1 - self.b) * (dd @ (self.kdiag * dd)) / self.den + self.b * dl ** 2))

    def grad(self, dd, dl):
        f = self(dd, dl)
        return (1 - self.b) * (dd * self.kdiag) / (f * self.den), self.b / f * dl


def newtonarclength(fem, material, Fext, stop=None, b=0.5, da=0.1, rtol=1e-16, ftol=1e-16, stol=0.0, maxsteps=20, maxits=100,
                    lin_rtol=1e-12, lin_maxits=200000):
    """newtonarclength(fem, Fint, dFint, Fext; stop, b, da, rtol, ftol, stol, maxsteps, maxits)
    (newtonarclength_fem.jl:2-169). Fint / dFint are the device operators of `material` on `fem` (which must not carry
    Neumann boundaries: the load enters through `Fext`, an array or a callable Fext(fem, F_ext) that fills one)."""
    if any(not bc.isdirichlet for bc in fem.boundaries):
        raise ValueError("arc-length continuation takes the load through Fext; remove the Neumann boundaries from fem")
    n = fem.ndof
    H = fem.handle(material)
    H.call("nls_set_body_force", 0, dptr(np.zeros(fem.mesh.dim)))
    if callable(Fext):
        F_ext = np.zeros(n)
        Fext(fem, F_ext)
    else:
        F_ext = np.array(Fext, dtype=np.float64)
    con = np.zeros(n, dtype=bool)
    for bc in fem.boundaries:
        con[np.asarray(bc.nodes, dtype=np.int64)] = True
    free = ~con
    res = ArcLengthResult(n, maxsteps, maxits)
    its, st, rr = C.c_int32(0), C.c_int32(0), C.c_double(0)

    def Fint(d):                                   # applydirichletboundaries!(fem, d); Fint(fem, d, F_int)
        d = np.ascontiguousarray(d)
        Fi = np.zeros(n)
        H.call("nls_residual", dptr(d), dptr(Fi))
        return d, Fi

    def tangent(d):                                # dFint(fem, d, Kd): the CSR values stay on the device for the solves
        H.call("nls_dev_upload_u", dptr(np.ascontiguousarray(d)))
        H.call("nls_dev_assemble", 0, 1)

    def solve(rhs_free):                           # dofs(Kd) \ rhs on the tangent assembled last
        bfull = np.zeros(n); bfull[free] = rhs_free
        x = np.zeros(n)
        H.call("nls_linear_solve_indefinite", lin_rtol, lin_maxits, dptr(bfull), dptr(x), C.byref(its), C.byref(rr), C.byref(st))
        res.lin_iterations += its.value
        if st.value != 0:
            raise ArithmeticError("linear solve failed (status %d, relres %.3e)" % (st.value, rr.value))
        return x[free]

    # initial tangent, its diagonal, the reference displacement increment (newtonarclength_fem.jl:37-45)
    K0 = CSRMatrix(fem)
    H.call("nls_jacobian", dptr(np.ascontiguousarray(res.d[0].copy())), dptr(K0.vals))
    rp, ci = K0.rowptr, K0.colind
    rows = np.repeat(np.arange(n), np.diff(rp))
    kdiag = np.zeros(n); dm = rows == ci
    kdiag[rows[dm]] = K0.vals[dm]
    kdiag = kdiag[free]
    Ff = F_ext[free]
    tangent(res.d[0])
    f_arc = ArcLength(kdiag, solve(Ff), b)
    ind_last, sign_last = None, 1.0
    expand = lambda v: (lambda o: (o.__setitem__(free, v), o)[1])(np.zeros(n))
    try:
        for s in range(maxsteps - 1):              # Julia: for n in 1:maxsteps-1 with n = s + 1
            d = res.d[s]
            if s + 1 < 3:                          # procedure 1 (newtonarclength_fem.jl:55-76)
                tangent(d)
                dtn = solve(Ff)
                dl = da / f_arc(dtn, 1.0)
                ind_n = np.sign(dtn @ Ff)
                if ind_last is None:
                    ind_last = ind_n               # the initial tangent is the tangent of step 1 (d = 0)
                sgn = -sign_last if ind_n == -ind_last else sign_last
                dl = sgn * abs(dl)
                sign_last, ind_last = sgn, ind_n
                dd = dl * dtn
                dk, lk = d + expand(dd), res.lam[s] + dl
            else:                                  # procedure 2: quadratic extrapolation (:77-83)
                lk = res.lam[s - 2] - 3 * res.lam[s - 1] + 3 * res.lam[s]
                dk = res.d[s - 2] - 3 * res.d[s - 1] + 3 * res.d[s]
                dl = lk - res.lam[s]
                dd = (dk - res.d[s])[free]
            dk, Fi = Fint(dk)
            r_d = (lk * F_ext - Fi)[free]; r_l = da - f_arc(dd, dl)
            norm_r0 = float(np.linalg.norm(r_d))
            if norm_r0 < np.finfo(float).eps:
                norm_r0 = 1.0
            res.res_d[s + 1, 0], res.res_lam[s + 1, 0] = norm_r0, r_l
            k = 1
            while k < maxits:
                if res.res_d[s + 1, k - 1] < rtol * norm_r0 and res.res_lam[s + 1, k - 1] / da < ftol:
                    break
                tangent(dk)
                g_d, g_l = f_arc.grad(dd, dl)
                v1, v2 = solve(r_d), solve(Ff)     # block elimination of [K -F; g_d g_l] [dd; dl] = [r_d; r_l]
                ddl = (r_l - g_d @ v1) / (g_l + g_d @ v2)
                ddd = v1 + ddl * v2
                if np.linalg.norm(ddd) < stol and abs(ddl) < stol:
                    break
                dk = dk + expand(ddd); dd = dd + ddd
                lk += ddl; dl += ddl
                dk, Fi = Fint(dk)
                r_d = (lk * F_ext - Fi)[free]; r_l = da - f_arc(dd, dl)
                res.res_d[s + 1, k], res.res_lam[s + 1, k] = float(np.linalg.norm(r_d)), r_l
                k += 1
            res.num_its[s + 1] = k
            res.d[s + 1], res.lam[s + 1] = dk, lk
            res.num_steps = s + 2
            if stop is not None and stop(res.d[s + 1], res.lam[s + 1]):
                break
    except ArithmeticError as e:                   # SingularException branch (:105-113): trim and return what was reached
        res.reason = "diverged linear solve: %s" % e
    return res.trim()
