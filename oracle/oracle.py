"""ctypes binding of the CPU ORACLE (oracle/libnls_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. Never imported by the product
package. See oracle/nls_oracle.h for what is restated and the reference
file:line each function follows.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAT_LINEAR, MAT_EXPONENTIAL, MAT_PLASTIC, MAT_NEOHOOKEAN = 0, 1, 2, 3
REASONS = ["NOT_YET_CONVERGED", "CONVERGED_RELATIVE", "CONVERGED_ABSOLUTE",
           "DIVERGED_MAXITS", "DIVERGED_STEP", "DIVERGED_LINEAR_SOLVE"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


class _Model(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nen", C.c_int32), ("nn", C.c_int64), ("nel", C.c_int64),
                ("conn", _ip), ("coords", _dp), ("p", C.c_int32), ("nq", C.c_int32),
                ("mat_kind", C.c_int32), ("mat", C.c_double * 8), ("area", C.c_double),
                ("ndir", C.c_int64), ("dir_dofs", _lp), ("dir_vals", _dp),
                ("nneu", C.c_int64), ("neu_dofs", _lp), ("neu_vals", _dp),
                ("sig", _dp), ("alp", _dp), ("kap", _dp), ("gam", _dp), ("dsig", _dp),
                ("sig_n", _dp), ("alp_n", _dp), ("kap_n", _dp), ("gam_n", _dp), ("u_n", _dp),
                ("nthreads", C.c_int32), ("has_body", C.c_int32), ("body", C.c_double * 3)]


class _Opts(C.Structure):
    _fields_ = [("type_modified", C.c_int32), ("atol", C.c_double), ("rtol", C.c_double),
                ("dtol", C.c_double), ("maxits", C.c_int32), ("lin_rtol", C.c_double),
                ("lin_maxits", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("reason", C.c_int32), ("iterations", C.c_int32), ("lin_its_total", C.c_int32),
                ("nhist", C.c_int32), ("norm_r", C.c_double), ("norm_r0", C.c_double)]


def build(force=False):
    """Compile oracle/libnls_oracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "libnls_oracle.so")
    src = os.path.join(_HERE, "nls_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_pattern.restype = C.c_int64
        _LIB.orc_integrate.restype = C.c_double
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


# ---- building blocks -------------------------------------------------------
def lagrange_nodes(p, kind="chebyshev2"):
    x = np.zeros(p + 1)
    lib().orc_lagrange_nodes(C.c_int32(p), C.c_int32(0 if kind == "chebyshev2" else 1), _d(x))
    return x


def lagrange_weights(nodes):
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    w = np.zeros_like(nodes)
    lib().orc_lagrange_weights(C.c_int32(len(nodes)), _d(nodes), _d(w))
    return w


def shape_N(nodes, w, xi):
    out = np.zeros(len(nodes))
    lib().orc_shape_N(C.c_int32(len(nodes)), _d(np.ascontiguousarray(nodes)), _d(np.ascontiguousarray(w)),
                      C.c_double(xi), _d(out))
    return out


def shape_dN(nodes, w, xi):
    out = np.zeros(len(nodes))
    lib().orc_shape_dN(C.c_int32(len(nodes)), _d(np.ascontiguousarray(nodes)), _d(np.ascontiguousarray(w)),
                       C.c_double(xi), _d(out))
    return out


def gauss(nq):
    x, w = np.zeros(nq), np.zeros(nq)
    if lib().orc_gauss(C.c_int32(nq), _d(x), _d(w)) != 0:
        raise ValueError("Gauss quadrature not implemented for n = %d" % nq)
    return x, w


def integrate(nq, fvals):
    f = np.ascontiguousarray(fvals, dtype=np.float64)
    return lib().orc_integrate(C.c_int32(nq), _d(f))


def dense_solve(A, b):
    A = np.array(A, dtype=np.float64, order="C")
    b = np.array(b, dtype=np.float64)
    rc = lib().orc_dense_solve(C.c_int64(len(b)), _d(A), _d(b))
    return rc, b


# ---- model -----------------------------------------------------------------
class OracleModel:
    """Holds the arrays an `orc_model` points at (FEMModel + material + BCs)."""

    def __init__(self, dim, conn, coords, p=1, nq=2, mat_kind=MAT_NEOHOOKEAN, mat=(1.0, 1.0), area=3e-4,
                 nthreads=1):
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, dim)
        self.dim = dim
        self.nel, self.nen = self.conn.shape
        self.nn = self.coords.shape[0]
        self.ndof = self.nn * dim
        self.m = _Model()
        m = self.m
        m.dim, m.nen, m.nn, m.nel = dim, self.nen, self.nn, self.nel
        m.conn = self.conn.ctypes.data_as(_ip)
        m.coords = _d(self.coords)
        m.p, m.nq, m.mat_kind = p, nq, mat_kind
        for i, v in enumerate(mat):
            m.mat[i] = v
        m.area = area
        m.nthreads = nthreads
        self.set_dirichlet([], [])
        self.set_neumann([], [])
        if mat_kind == MAT_PLASTIC:
            n = self.nel * nq
            self._state = {k: np.zeros(n) for k in
                           ("sig", "alp", "kap", "gam", "dsig", "sig_n", "alp_n", "kap_n", "gam_n")}
            self._state["u_n"] = np.zeros(self.nel * self.nen)
            for k, v in self._state.items():
                setattr(m, k, _d(v))
            lib().orc_init_state(C.byref(m))
        self._pattern = None

    def set_body_force(self, f):
        """DistributedForce density per component (gallery/distributedforce.jl); None removes it."""
        f = [] if f is None else list(np.atleast_1d(f))
        self.m.has_body = int(any(v != 0.0 for v in f))
        for i in range(3):
            self.m.body[i] = f[i] if i < len(f) else 0.0

    def set_threads(self, n):
        self.m.nthreads = n

    def set_dirichlet(self, dofs, vals):
        self.dir_dofs = np.ascontiguousarray(dofs, dtype=np.int64)
        self.dir_vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.m.ndir = len(self.dir_dofs)
        self.m.dir_dofs = self.dir_dofs.ctypes.data_as(_lp)
        self.m.dir_vals = _d(self.dir_vals)

    def set_neumann(self, dofs, vals):
        self.neu_dofs = np.ascontiguousarray(dofs, dtype=np.int64)
        self.neu_vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.m.nneu = len(self.neu_dofs)
        self.m.neu_dofs = self.neu_dofs.ctypes.data_as(_lp)
        self.m.neu_vals = _d(self.neu_vals)

    def state(self, name):
        return self._state[name].reshape(self.nel, -1)

    def commit_state(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        lib().orc_commit_state(C.byref(self.m), _d(U))

    def pattern(self):
        if self._pattern is None:
            rowptr = np.zeros(self.ndof + 1, dtype=np.int64)
            nnz = lib().orc_pattern(C.byref(self.m), rowptr.ctypes.data_as(_lp), None)
            colind = np.zeros(nnz, dtype=np.int32)
            lib().orc_pattern(C.byref(self.m), rowptr.ctypes.data_as(_lp), colind.ctypes.data_as(_ip))
            self._pattern = (rowptr, colind)
        return self._pattern

    def residual(self, U):
        """Returns (U_with_dirichlet, R) -- the reference mutates U in place (residual.jl:13)."""
        U = np.array(U, dtype=np.float64)
        R = np.zeros(self.ndof)
        lib().orc_residual(C.byref(self.m), _d(U), _d(R))
        return U, R

    def jacobian_dense(self, U):
        U = np.array(U, dtype=np.float64)
        K = np.zeros((self.ndof, self.ndof))
        lib().orc_jacobian_dense(C.byref(self.m), _d(U), _d(K))
        return K

    def jacobian_csr(self, U):
        U = np.array(U, dtype=np.float64)
        rowptr, colind = self.pattern()
        vals = np.zeros(len(colind))
        lib().orc_jacobian_csr(C.byref(self.m), _d(U), rowptr.ctypes.data_as(_lp),
                               colind.ctypes.data_as(_ip), _d(vals))
        return vals

    def contribution_scale(self, U):
        """(Rabs, Kabs): per entry of R and of the CSR Jacobian the sum of the absolute element contributions --
        the denominator of the entry-wise relative error (tests/helpers.entry_err). A test metric, not a
        reference function."""
        U = np.array(U, dtype=np.float64)
        rowptr, colind = self.pattern()
        Rabs, vabs = np.zeros(self.ndof), np.zeros(len(colind))
        lib().orc_assemble_abs(C.byref(self.m), _d(U), _d(Rabs), rowptr.ctypes.data_as(_lp),
                               colind.ctypes.data_as(_ip), _d(vabs))
        return Rabs, vabs

    def element(self, e, ue):
        ue = np.ascontiguousarray(ue, dtype=np.float64).ravel()
        nd = self.nen * self.dim
        fe, Ke = np.zeros(nd), np.zeros((nd, nd))
        lib().orc_element(C.byref(self.m), C.c_int64(e), _d(ue), _d(fe), _d(Ke))
        return fe, Ke

    def spmv(self, vals, x):
        rowptr, colind = self.pattern()
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.ndof)
        lib().orc_spmv(C.c_int64(self.ndof), rowptr.ctypes.data_as(_lp), colind.ctypes.data_as(_ip),
                       _d(np.ascontiguousarray(vals)), _d(x), _d(y), C.c_int32(self.m.nthreads))
        return y

    def constrained_mask(self):
        con = np.zeros(self.ndof, dtype=np.uint8)
        con[self.dir_dofs] = 1
        return con

    def pcg(self, vals, b, rtol=1e-12, maxits=10000):
        rowptr, colind = self.pattern()
        con = self.constrained_mask()
        x = np.zeros(self.ndof)
        relres, its = C.c_double(0), C.c_int32(0)
        rc = lib().orc_pcg(C.c_int64(self.ndof), rowptr.ctypes.data_as(_lp), colind.ctypes.data_as(_ip),
                           _d(np.ascontiguousarray(vals)), con.ctypes.data_as(_bp),
                           _d(np.ascontiguousarray(b, dtype=np.float64)), _d(x), C.c_double(rtol),
                           C.c_int32(maxits), C.byref(relres), C.byref(its), C.c_int32(self.m.nthreads))
        return rc, x, its.value, relres.value

    def newton(self, U0=None, rtol=1e-16, atol=1e-16, dtol=1e6, maxits=100, modified=False,
               linear="dense", lin_rtol=1e-12, lin_maxits=100000):
        """solve!(::NewtonSolver, U0) (newton_fem.jl:132-190). linear='dense' is the
        faithful variant (LU); 'pcg' the scalable one (CSR + Jacobi-PCG)."""
        U0 = np.zeros(self.ndof) if U0 is None else np.ascontiguousarray(U0, dtype=np.float64)
        U, R = np.zeros(self.ndof), np.zeros(self.ndof)
        hist = np.zeros(maxits + 2)
        o = _Opts(int(modified), atol, rtol, dtol, maxits, lin_rtol, lin_maxits)
        res = _Result()
        if linear == "dense":
            lib().orc_newton_dense(C.byref(self.m), C.byref(o), _d(U0), _d(U), _d(R), _d(hist), C.byref(res))
        else:
            rowptr, colind = self.pattern()
            vals = np.zeros(len(colind))
            lib().orc_newton_pcg(C.byref(self.m), C.byref(o), rowptr.ctypes.data_as(_lp),
                                 colind.ctypes.data_as(_ip), _d(vals), _d(U0), _d(U), _d(R), _d(hist),
                                 C.byref(res))
        return dict(U=U, R=R, reason=REASONS[res.reason], iterations=res.iterations,
                    lin_its=res.lin_its_total, hist=hist[:res.nhist].copy(),
                    norm_r=res.norm_r, norm_r0=res.norm_r0)

    # ---- dynamics: gallery/mass.jl, gallery/residual.jl:48-54, gallery/jacobian.jl:42-51, timesteppers/newmark.jl ----
    def mass_dense(self, rho, area=1.0):
        """assemblemass! into a zero matrix (mass.jl:8-17): dense ndof x ndof, M_{ai,bk} = delta_ik m_ab."""
        Mn = np.zeros((self.nn, self.nn))
        lib().orc_mass_nodes(C.byref(self.m), C.c_double(rho), C.c_double(area), _d(Mn))
        return np.kron(Mn, np.eye(self.dim))

    def mass_csr(self, rho, area=1.0):
        rowptr, colind = self.pattern()
        M = self.mass_dense(rho, area)
        vals = np.zeros(len(colind))
        for i in range(self.ndof):
            vals[rowptr[i]:rowptr[i + 1]] = M[i, colind[rowptr[i]:rowptr[i + 1]]]
        return vals

    def dynamic_residual(self, U, Udd, rho, area=1.0):
        """Residual dynamic method (residual.jl:48-54): static residual, then applymass!(U̇̇) added to R."""
        Ud, R = self.residual(U)
        return Ud, R + self.mass_dense(rho, area) @ np.asarray(Udd, dtype=np.float64)

    def dynamic_jacobian_dense(self, U, scale, rho, area=1.0):
        """Jacobian dynamic method (jacobian.jl:42-51): K .*= dUdd/dU scale, then assemblemass! added."""
        return self.jacobian_dense(U) * scale + self.mass_dense(rho, area)

    def newmark(self, rho, dt, maxtime, beta=0.25, gamma=0.5, area=1.0, U0=None, Ud0=None, bc_update=None,
                rtol=1e-16, atol=1e-16, dtol=1e6, maxits=100, consistent_a0=True):
        """solve!(ts::NewmarkTimeStepper, U0, U̇0, U̇̇0) (newmark.jl:124-156) with dense linear algebra (small cases).
        Newton runs on the acceleration (closures of newmark.jl:75-107, loop of newton_fem.jl:132-190).
        Semantics kept from the reference: the solve is on the free DoFs only, so constrained entries of U̇̇ keep
        their initial-guess values; Dirichlet values enter through the copy of U the operators work on.
        Deviations from the text of newmark.jl, both documented in DESIGN.md: U0/U̇0 are kept (the reference's first
        residual call runs update_state! with zero predictors and overwrites them), and with consistent_a0 the
        initial acceleration solves M a0 = -R(U0) (the reference writes that residual into a temporary and ends up
        with a0 = 0; consistent_a0=False reproduces that).
        bc_update(t) -> (dir_vals, neu_vals) or None updates the time-dependent boundaries (step!, types.jl:68-72)."""
        n = self.ndof
        nsteps = int(round((maxtime - 0.0) / dt + 1))
        U = np.zeros((nsteps, n)); Ud = np.zeros((nsteps, n)); Udd = np.zeros((nsteps, n))
        its = np.zeros(nsteps, dtype=int); res = np.zeros((nsteps, n))
        if U0 is not None: U[0] = U0
        if Ud0 is not None: Ud[0] = Ud0
        M = self.mass_dense(rho, area)
        con = self.constrained_mask().astype(bool)
        free = ~con
        if consistent_a0:
            _, R0 = self.residual(U[0])
            Udd[0] = np.linalg.solve(M, -R0)
        t, step = 0.0, 0
        s = dt * dt * beta
        while t < maxtime and step < nsteps - 1:
            Ut = U[step] + dt * Ud[step] + dt * dt * (0.5 - beta) * Udd[step]
            Udt = Ud[step] + dt * (1 - gamma) * Udd[step]
            t += dt; step += 1
            if bc_update is not None:
                dv, nv = bc_update(t)
                if dv is not None: self.dir_vals[:] = dv
                if nv is not None: self.neu_vals[:] = nv
            a = Udd[step - 1].copy()
            def resid(a):
                Uc, R = self.residual(Ut + s * a)
                return R + M @ a
            R = resid(a)
            nr = np.linalg.norm(R[free]); nr0 = nr if nr > np.finfo(float).eps else 1.0
            k = 0; reason = "NOT_YET_CONVERGED"
            while k < maxits:
                if nr / nr0 < rtol: reason = "CONVERGED_RELATIVE"; break
                if nr < atol: reason = "CONVERGED_ABSOLUTE"; break
                K = self.jacobian_dense(Ut + s * a) * s + M
                d = np.zeros(n)
                d[free] = np.linalg.solve(K[np.ix_(free, free)], -R[free])
                if np.linalg.norm(d) > dtol: reason = "DIVERGED_STEP"; break
                a = a + d
                R = resid(a)
                nr = np.linalg.norm(R[free]); k += 1
            else:
                reason = "DIVERGED_MAXITS"
            if k >= maxits and not reason.startswith("CONV") and reason != "DIVERGED_STEP": reason = "DIVERGED_MAXITS"
            if not reason.startswith("CONVERGED"):
                return dict(U=U[:step + 1], Ud=Ud[:step + 1], Udd=Udd[:step + 1], num_its=its[:step + 1], residual=res[:step + 1],
                            reason=reason, steps=step)
            Udd[step] = a; res[step] = R; its[step] = k
            U[step] = Ut + s * a
            Ud[step] = Udt + dt * gamma * a
        return dict(U=U[:step + 1], Ud=Ud[:step + 1], Udd=Udd[:step + 1], num_its=its[:step + 1], residual=res[:step + 1],
                    reason="CONVERGED", steps=step)

    def partition(self, nranks, node_off, rank):
        off = np.ascontiguousarray(node_off, dtype=np.int64)
        ne, ng = C.c_int64(0), C.c_int64(0)
        sp = np.zeros(nranks + 1, dtype=np.int64)
        lib().orc_partition(C.byref(self.m), C.c_int32(nranks), off.ctypes.data_as(_lp), C.c_int32(rank),
                            C.byref(ne), None, C.byref(ng), None, sp.ctypes.data_as(_lp), None)
        le = np.zeros(ne.value, dtype=np.int64)
        gn = np.zeros(ng.value, dtype=np.int64)
        sn = np.zeros(sp[-1], dtype=np.int64)
        lib().orc_partition(C.byref(self.m), C.c_int32(nranks), off.ctypes.data_as(_lp), C.c_int32(rank),
                            C.byref(ne), le.ctypes.data_as(_lp), C.byref(ng), gn.ctypes.data_as(_lp),
                            sp.ctypes.data_as(_lp), sn.ctypes.data_as(_lp))
        return dict(loc_elems=le, ghost_nodes=gn, send_ptr=sp, send_nodes=sn)


# ---- arc-length continuation: nonlinearsolvers/newtonarclength_fem.jl:2-169, dense restatement ----------
def newtonarclength(om, F_ext, b=0.5, da=0.1, rtol=1e-16, ftol=1e-16, stol=0.0, maxsteps=20, maxits=100, stop=None,
                    sign_rule="det"):
    """newtonarclength(fem, Fint, dFint, Fext; ...) restated with dense linear algebra for small models.
    `om`: OracleModel without Neumann boundaries (Fint = its residual), F_ext: load vector (what Fext(fem, F_ext) fills).
    The ArcLength functor and its gradient are newtonarclength.jl:6-22; the bordered system is solved as written
    (newtonarclength_fem.jl:103-104). Calls the reference makes to functions that do not exist (updateboundaries!(fem),
    step!(fem, dlambda), postprocess!(fem, d, res): newtonarclength_fem.jl:49,155-156) are dropped.
    sign_rule: "det" = sign change of det(K) as in the reference (:64-70); "stiffness" = sign of dtilde . F_ext (what the
    device path uses, which has no determinant)."""
    n = om.ndof
    free = ~om.constrained_mask().astype(bool)
    F_ext = np.asarray(F_ext, dtype=np.float64)
    d_hist = np.zeros((maxsteps, n)); lam = np.zeros(maxsteps); num_its = np.zeros(maxsteps, dtype=int)
    res_d = np.zeros((maxsteps, maxits + 1)); res_l = np.zeros((maxsteps, maxits + 1))
    Kf = lambda d: om.jacobian_dense(d)[np.ix_(free, free)]
    Fint = lambda d: om.residual(d)        # (d with Dirichlet values, F_int)
    K0 = Kf(d_hist[0])
    Ff = F_ext[free]
    dt0 = np.linalg.solve(K0, Ff)
    kdiag = np.diag(K0).copy()
    den = dt0 @ (kdiag * dt0)
    f_arc = lambda dd, dl: np.sqrt((1 - b) * (dd @ (kdiag * dd)) / den + b * dl ** 2)
    ind = lambda K, dt: (np.linalg.slogdet(K)[0] if sign_rule == "det" else np.sign(dt @ Ff))
    ind_last, sign_last = ind(K0, dt0), 1.0
    expand = lambda v: (lambda o: (o.__setitem__(free, v), o)[1])(np.zeros(n))
    nsteps = 1
    for s in range(maxsteps - 1):                          # Julia: for n in 1:maxsteps-1, n = s + 1
        d = d_hist[s]
        if s + 1 < 3:                                      # procedure 1
            Kd = Kf(d)
            dtn = np.linalg.solve(Kd, Ff)
            dl = da / f_arc(dtn, 1.0)
            ind_n = ind(Kd, dtn)
            sgn = -sign_last if ind_n == -ind_last else sign_last
            dl = sgn * abs(dl)
            sign_last, ind_last = sgn, ind_n
            dd = dl * dtn
            dk = d + expand(dd)
            lk = lam[s] + dl
        else:                                              # procedure 2: quadratic extrapolation
            lk = lam[s - 2] - 3 * lam[s - 1] + 3 * lam[s]
            dk = d_hist[s - 2] - 3 * d_hist[s - 1] + 3 * d_hist[s]
            dl = lk - lam[s]
            dd = (dk - d_hist[s])[free]
        dk, Fi = Fint(dk)
        r_d = (lk * F_ext - Fi)[free]; r_l = da - f_arc(dd, dl)
        norm_r0 = np.linalg.norm(r_d)
        if norm_r0 < np.finfo(float).eps:
            norm_r0 = 1.0
        res_d[s + 1, 0], res_l[s + 1, 0] = norm_r0, r_l
        k = 1
        while k < maxits:
            if res_d[s + 1, k - 1] < rtol * norm_r0 and res_l[s + 1, k - 1] / da < ftol:
                break
            Kd = Kf(dk)
            f = f_arc(dd, dl)
            g_d = (1 - b) * (dd * kdiag) / (f * den); g_l = b / f * dl
            Kb = np.block([[Kd, -Ff[:, None]], [g_d[None, :], np.array([[g_l]])]])
            delta = np.linalg.solve(Kb, np.concatenate([r_d, [r_l]]))
            if np.linalg.norm(delta[:-1]) < stol and abs(delta[-1]) < stol:
                break
            dk = dk + expand(delta[:-1]); dd = dd + delta[:-1]
            lk += delta[-1]; dl += delta[-1]
            dk, Fi = Fint(dk)
            r_d = (lk * F_ext - Fi)[free]; r_l = da - f_arc(dd, dl)
            res_d[s + 1, k], res_l[s + 1, k] = np.linalg.norm(r_d), r_l
            k += 1
        num_its[s + 1] = k
        d_hist[s + 1] = dk; lam[s + 1] = lk
        nsteps = s + 2
        if stop is not None and stop(d_hist[s + 1], lam[s + 1]):
            break
    return dict(d=d_hist[:nsteps], lam=lam[:nsteps], num_its=num_its[:nsteps], res_d=res_d[:nsteps], res_l=res_l[:nsteps],
                num_steps=nsteps)


class J2Oracle:
    """CPU restatement (test infrastructure) of the stateful-material protocol of the reference's LinearElastoPlasticity1D
    (src/materials/linearelastoplasticity.jl:60-117: return mapping from the committed state in the internal-force pass,
    stored tangent read by the tangent pass, commit on save_state!) for small-strain J2 plasticity with linear isotropic
    hardening on Q1 quads (plane strain) / hexahedra -- SURVEY 8 f-2. The reference has no 2-D / 3-D plastic material, so
    parity of this extension is pinned by known answers only (tests/test_oracle.py: uniaxial yield stress, consistency of the
    tangent with finite differences of the stress, elastic unloading). Written independently of the CUDA kernel: full 3 x 3
    tensors and the 3 x 3 x 3 x 3 tangent, contracted with einsum; plain numpy, small cases only."""

    def __init__(self, dim, conn, coords, K, G, sigma_y, H):
        self.dim, self.nen = dim, 2 ** dim
        self.conn = np.asarray(conn, dtype=np.int64).reshape(-1, self.nen)
        self.coords = np.asarray(coords, dtype=np.float64).reshape(-1, dim)
        self.nel, self.nn = len(self.conn), len(self.coords)
        self.K, self.G, self.sy, self.H = float(K), float(G), float(sigma_y), float(H)
        g = 1.0 / math.sqrt(3.0)
        # shapefunctions.jl / quadrature.jl for p = 1, two Gauss points per direction; nodes and points x-fastest
        self.dN = np.zeros((self.nen, self.nen, dim))           # [q][a][D]
        for q in range(self.nen):
            xi = [(-g if not (q >> d) & 1 else g) for d in range(dim)]
            for a in range(self.nen):
                s = [(-1.0 if not (a >> d) & 1 else 1.0) for d in range(dim)]
                for D in range(dim):
                    v = 1.0
                    for d in range(dim):
                        v *= 0.5 * s[d] if d == D else 0.5 * (1.0 + s[d] * xi[d])
                    self.dN[q, a, D] = v
        self.w = np.ones(self.nen)
        self.dir_dofs, self.dir_vals = np.zeros(0, dtype=np.int64), np.zeros(0)
        self.init_state()

    def init_state(self):
        nq = self.nen
        self.epsp_n = np.zeros((self.nel, nq, 3, 3)); self.alpha_n = np.zeros((self.nel, nq))
        self.epsp = np.zeros((self.nel, nq, 3, 3)); self.alpha = np.zeros((self.nel, nq))
        self.C = None                                            # consistent tangent of the last internal-force pass

    def set_dirichlet(self, dofs, vals):
        self.dir_dofs, self.dir_vals = np.asarray(dofs, dtype=np.int64), np.asarray(vals, dtype=np.float64)

    def commit_state(self):
        self.epsp_n[:] = self.epsp; self.alpha_n[:] = self.alpha

    def _kinematics(self, U):
        dim = self.dim
        X = self.coords[self.conn]                               # [e][a][d]
        u = U.reshape(-1, dim)[self.conn]
        J = np.einsum("ead,qaD->eqdD", X, self.dN)
        detJ = np.linalg.det(J)
        Gr = np.einsum("qaE,eqED->eqaD", self.dN, np.linalg.inv(J))       # dN/dX
        gu = np.zeros((self.nel, self.nen, 3, 3))
        gu[:, :, :dim, :dim] = np.einsum("eai,eqaj->eqij", u, Gr)
        return Gr, detJ * self.w[None, :], 0.5 * (gu + gu.transpose(0, 1, 3, 2))

    def return_map(self, eps):
        """stress, new plastic strain, new alpha, tangent [..., 3,3,3,3] at strains eps [..., 3, 3] from the committed state"""
        K, G, sy, H = self.K, self.G, self.sy, self.H
        I = np.eye(3)
        ee = eps - self.epsp_n
        tr = np.trace(ee, axis1=-2, axis2=-1)
        s_tr = 2.0 * G * (ee - tr[..., None, None] / 3.0 * I)
        qn = np.sqrt(np.einsum("...ij,...ij->...", s_tr, s_tr))
        f = qn - math.sqrt(2.0 / 3.0) * (sy + H * self.alpha_n)
        plastic = f > 0.0
        qsafe = np.where(plastic, qn, 1.0)
        dg = np.where(plastic, f / (2.0 * G + 2.0 * H / 3.0), 0.0)
        n = np.where(plastic[..., None, None], s_tr / qsafe[..., None, None], 0.0)
        sig = K * tr[..., None, None] * I + s_tr - 2.0 * G * dg[..., None, None] * n
        theta = 1.0 - 2.0 * G * dg / qsafe
        thetab = np.where(plastic, 1.0 / (1.0 + H / (3.0 * G)) - (1.0 - theta), 0.0)
        II = np.einsum("ij,kl->ijkl", I, I)
        Is = 0.5 * (np.einsum("ik,jl->ijkl", I, I) + np.einsum("il,jk->ijkl", I, I))
        C = (K * II + 2.0 * G * theta[..., None, None, None, None] * (Is - II / 3.0)
             - 2.0 * G * thetab[..., None, None, None, None] * np.einsum("...ij,...kl->...ijkl", n, n))
        return sig, self.epsp_n + dg[..., None, None] * n, self.alpha_n + math.sqrt(2.0 / 3.0) * dg, C

    def residual(self, U):
        """(U with the Dirichlet values written, internal force); runs the return mapping and stores the trial state."""
        U = np.array(U, dtype=np.float64)
        U[self.dir_dofs] = self.dir_vals
        dim = self.dim
        Gr, wd, eps = self._kinematics(U)
        sig, self.epsp, self.alpha, self.C = self.return_map(eps)
        fe = np.einsum("eq,eqij,eqaj->eai", wd, sig[:, :, :dim, :dim], Gr)
        R = np.zeros(self.nn * dim)
        np.add.at(R.reshape(-1, dim), self.conn, fe)
        return U, R

    def jacobian_dense(self, U):
        """tangent from the data stored by the last internal-force pass (the reference's protocol)"""
        assert self.C is not None, "residual first: the tangent pass reads the state of the internal-force pass"
        U = np.array(U, dtype=np.float64)
        U[self.dir_dofs] = self.dir_vals
        dim = self.dim
        Gr, wd, _ = self._kinematics(U)
        Ke = np.einsum("eq,eqaj,eqijkl,eqbl->eaibk", wd, Gr, self.C[:, :, :dim, :dim, :dim, :dim], Gr)
        n = self.nn * dim
        Kd = np.zeros((n, n))
        dofs = (self.conn[:, :, None] * dim + np.arange(dim)[None, None, :]).reshape(self.nel, -1)
        for e in range(self.nel):
            Kd[np.ix_(dofs[e], dofs[e])] += Ke[e].reshape(self.nen * dim, self.nen * dim)
        return Kd

    def newton(self, U0, rtol=1e-10, atol=1e-14, maxits=25):
        """newton_fem.jl:132-190 with a dense solve on the free DoFs; returns (U, iterations, residual norms)"""
        n = self.nn * self.dim
        free = np.ones(n, dtype=bool); free[self.dir_dofs] = False
        U, R = self.residual(U0)
        hist = [float(np.linalg.norm(R[free]))]
        r0 = hist[0] if hist[0] > 2.220446049250313e-16 else 1.0
        its = 0
        while its < maxits:
            if hist[-1] / r0 < rtol or hist[-1] < atol:
                break
            Kd = self.jacobian_dense(U)
            dU = np.zeros(n)
            dU[free] = np.linalg.solve(Kd[np.ix_(free, free)], -R[free])
            U, R = self.residual(U + dU)
            hist.append(float(np.linalg.norm(R[free])))
            its += 1
        return U, its, hist
