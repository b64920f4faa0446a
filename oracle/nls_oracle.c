/*
 * nls_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See nls_oracle.h.
 *
 * Every function cites the reference file:line it restates
 * (paths relative to /root/reference). Operation order follows SURVEY.md
 * Appendix C, because parity is checked at 1e-12.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC, pthreads)
 */
#include "nls_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* Minimal fork-join helper (this image's gcc has no libgomp): run fn(ctx, lo, hi)
 * over [0, n) split into `nth` contiguous chunks. nth <= 1 runs inline (serial,
 * deterministic -- the mode every parity check uses). */
typedef void (*orc_range_fn)(void *ctx, int64_t lo, int64_t hi, int tid);
typedef struct { orc_range_fn fn; void *ctx; int64_t lo, hi; int tid; } orc_task;
static void *orc_task_run(void *p) { orc_task *t = (orc_task *)p; t->fn(t->ctx, t->lo, t->hi, t->tid); return NULL; }
static void orc_parallel_for(int nth, int64_t n, orc_range_fn fn, void *ctx) {
  if (nth <= 1 || n < 2 * nth) { fn(ctx, 0, n, 0); return; }
  if (nth > 256) nth = 256;
  pthread_t th[256]; orc_task tk[256];
  for (int t = 0; t < nth; ++t) {
    tk[t].fn = fn; tk[t].ctx = ctx; tk[t].tid = t;
    tk[t].lo = n * t / nth; tk[t].hi = n * (t + 1) / nth;
    if (t > 0) pthread_create(&th[t], NULL, orc_task_run, &tk[t]);
  }
  orc_task_run(&tk[0]);
  for (int t = 1; t < nth; ++t) pthread_join(th[t], NULL);
}
static inline void orc_atomic_add(double *p, double v) {
  union { double d; uint64_t u; } old, nw;
  old.u = __atomic_load_n((uint64_t *)p, __ATOMIC_RELAXED);
  do { nw.d = old.d + v; } while (!__atomic_compare_exchange_n((uint64_t *)p, &old.u, nw.u, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

#define ORC_MAXN 16 /* max nodes per 1-D element (p <= 15) */
#define ORC_MAXQ 5

/* ------------------------------------------------------------------ */
/* Shape functions: src/fem/shapefunctions.jl                          */
/* ------------------------------------------------------------------ */

/* Julia isapprox(x, y) with default rtol = sqrt(eps), atol = 0 (used by `.≈`
 * in shapefunctions.jl:51,66). */
static int orc_isapprox(double x, double y) {
  if (x == y) return 1;
  if (!isfinite(x) || !isfinite(y)) return 0;
  double rtol = 1.4901161193847656e-8; /* sqrt(eps(Float64)) */
  double ax = fabs(x), ay = fabs(y);
  return fabs(x - y) <= rtol * (ax > ay ? ax : ay);
}

/* shapefunctions.jl:38-47  lagrange(order, :equidistant | :chebyshev2) node sets */
void orc_lagrange_nodes(int32_t p, int32_t kind, double *nodes) {
  int m = p + 1;
  if (kind == 1) { /* LinRange(-1, 1, order+1): lerp (1-t)*a + t*b, t = j/(len-1) */
    for (int j = 0; j < m; ++j) {
      double t = (m > 1) ? (double)j / (double)(m - 1) : 0.0;
      nodes[j] = (1.0 - t) * (-1.0) + t * 1.0;
    }
  } else { /* [cos(i*pi/order) for i in order:-1:0], shapefunctions.jl:43 */
    for (int j = 0; j < m; ++j) {
      int i = p - j;
      nodes[j] = cos((double)i * M_PI / (double)p);
    }
  }
}

/* shapefunctions.jl:23-30  barycentric weights: two separate products, then reciprocal */
void orc_lagrange_weights(int32_t m, const double *x, double *w) {
  for (int i = 0; i < m; ++i) {
    double p1 = 1.0, p2 = 1.0;
    for (int k = 0; k < i; ++k) p1 *= (x[i] - x[k]);
    for (int k = i + 1; k < m; ++k) p2 *= (x[i] - x[k]);
    w[i] = p1 * p2;
  }
  for (int i = 0; i < m; ++i) w[i] = 1.0 / w[i];
}

/* shapefunctions.jl:50-61  N(shape, xi) */
void orc_shape_N(int32_t m, const double *x, const double *w, double xi, double *N) {
  int any = 0;
  for (int i = 0; i < m; ++i) any |= orc_isapprox(xi, x[i]);
  if (any) {
    for (int i = 0; i < m; ++i) N[i] = orc_isapprox(xi, x[i]) ? 1.0 : 0.0;
    return;
  }
  double l = 1.0;
  for (int i = 0; i < m; ++i) l = (i == 0) ? (xi - x[0]) : l * (xi - x[i]);
  for (int i = 0; i < m; ++i) N[i] = (l * 1.0) * (w[i] / (xi - x[i]));
}

/* shapefunctions.jl:64-81  gradN(shape, xi) */
void orc_shape_dN(int32_t m, const double *x, const double *w, double xi, double *dN) {
  int idx = -1;
  for (int i = 0; i < m; ++i)
    if (orc_isapprox(xi, x[i])) { idx = i; break; }
  for (int i = 0; i < m; ++i) dN[i] = 0.0;
  if (idx >= 0) { /* on-node branch, :69-74 */
    for (int j = 0; j < m; ++j)
      if (j != idx) dN[j] = (w[j] / w[idx]) / (x[idx] - x[j]);
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += dN[j];
    dN[idx] = -s;
    return;
  }
  double N[ORC_MAXN];
  orc_shape_N(m, x, w, xi, N);
  for (int i = 0; i < m; ++i) { /* :77-80 */
    double s = 0.0;
    int first = 1;
    for (int j = 0; j < m; ++j) {
      if (j == i) continue;
      double t = 1.0 / (xi - x[j]);
      s = first ? t : s + t;
      first = 0;
    }
    dN[i] = N[i] * s;
  }
}

/* ------------------------------------------------------------------ */
/* Quadrature: src/fem/quadrature.jl:29-50 (closed forms as written)   */
/* ------------------------------------------------------------------ */
int32_t orc_gauss(int32_t nq, double *x, double *w) {
  switch (nq) {
  case 1: x[0] = 0.0; w[0] = 2.0; return 0;
  case 2:
    x[0] = -1.0 / sqrt(3.0); x[1] = 1.0 / sqrt(3.0);
    w[0] = 1.0; w[1] = 1.0; return 0;
  case 3:
    x[0] = -sqrt(3.0) / sqrt(5.0); x[1] = 0.0; x[2] = sqrt(3.0) / sqrt(5.0);
    w[0] = 5.0 / 9.0; w[1] = 8.0 / 9.0; w[2] = 5.0 / 9.0; return 0;
  case 4: {
    double x1 = sqrt(3.0 / 7.0 - 2.0 * sqrt(6.0) / (7.0 * sqrt(5.0)));
    double x2 = sqrt(3.0 / 7.0 + 2.0 * sqrt(6.0) / (7.0 * sqrt(5.0)));
    double w1 = (18.0 + sqrt(30.0)) / 36.0, w2 = (18.0 - sqrt(30.0)) / 36.0;
    x[0] = -x2; x[1] = -x1; x[2] = x1; x[3] = x2;
    w[0] = w2; w[1] = w1; w[2] = w1; w[3] = w2; return 0; }
  case 5: {
    double x1 = 1.0 / 3.0 * sqrt(5.0 - 2.0 * sqrt(10.0) / sqrt(7.0));
    double x2 = 1.0 / 3.0 * sqrt(5.0 + 2.0 * sqrt(10.0) / sqrt(7.0));
    double w1 = (322.0 + 13.0 * sqrt(70.0)) / 900.0, w2 = (322.0 - 13.0 * sqrt(70.0)) / 900.0;
    x[0] = -x2; x[1] = -x1; x[2] = 0.0; x[3] = x1; x[4] = x2;
    w[0] = w2; w[1] = w1; w[2] = 128.0 / 225.0; w[3] = w1; w[4] = w2; return 0; }
  default: return -1; /* quadrature.jl:49 error(...) */
  }
}

/* quadrature.jl:68-70  integrate(q, f::AbstractVecOrMat) = sum(f .* weights(q)) */
double orc_integrate(int32_t nq, const double *f) {
  double x[ORC_MAXQ], w[ORC_MAXQ];
  if (orc_gauss(nq, x, w)) return NAN;
  double s = f[0] * w[0];
  for (int q = 1; q < nq; ++q) s = s + f[q] * w[q];
  return s;
}

/* ------------------------------------------------------------------ */
/* Materials: src/materials/ (all .jl)                                     */
/* ------------------------------------------------------------------ */
static double orc_sigma(const orc_model *m, double eps) {
  if (m->mat_kind == ORC_MAT_LINEAR) return m->mat[0] * eps;          /* linearelasticity.jl:18 */
  return m->mat[0] * (1.0 - exp(-m->mat[1] * eps));                   /* exponential.jl:20 */
}
static double orc_dsigma(const orc_model *m, double eps) {
  if (m->mat_kind == ORC_MAT_LINEAR) return m->mat[0];                /* linearelasticity.jl:19 */
  return m->mat[1] * m->mat[0] * exp(-m->mat[1] * eps);               /* exponential.jl:21 */
}

/* ------------------------------------------------------------------ */
/* 1-D element kernels: src/materials/defaults.jl:2-26 and             */
/* linearelastoplasticity.jl:23-48                                     */
/* ------------------------------------------------------------------ */
typedef struct {
  int m, nq;
  double xq[ORC_MAXQ], wq[ORC_MAXQ];
  double dN[ORC_MAXQ][ORC_MAXN];
} orc_tab1d;

static void orc_tab1d_init(const orc_model *mo, orc_tab1d *t) {
  double nodes[ORC_MAXN], w[ORC_MAXN];
  t->m = mo->p + 1; t->nq = mo->nq;
  orc_lagrange_nodes(mo->p, 0, nodes); /* FEMModel always uses :chebyshev2, fem.jl:20 */
  orc_lagrange_weights(t->m, nodes, w);
  orc_gauss(mo->nq, t->xq, t->wq);
  for (int q = 0; q < t->nq; ++q) orc_shape_dN(t->m, nodes, w, t->xq[q], t->dN[q]);
}

static double orc_el_dxdxi(const orc_model *mo, int64_t e) {
  const int32_t *c = mo->conn + e * mo->nen;
  double len = mo->coords[c[mo->nen - 1]] - mo->coords[c[0]]; /* element.jl:15 */
  return len / 2.0;
}

/* f_e and/or K_e of element e for the 1-D path. fe, Ke may be NULL. */
static void orc_element_1d(const orc_model *mo, const orc_tab1d *t, int64_t e, const double *u,
                           double *fe, double *Ke) {
  int m = t->m;
  double A = mo->area;
  double dxdxi = orc_el_dxdxi(mo, e);
  double inv = 1.0 / dxdxi;
  double B[ORC_MAXN];
  for (int q = 0; q < t->nq; ++q) {
    for (int i = 0; i < m; ++i) B[i] = t->dN[q][i] * inv; /* defaults.jl:7 */
    double s, d;
    if (mo->mat_kind == ORC_MAT_PLASTIC) { /* stored qp data, linearelastoplasticity.jl:28-33,42-47 */
      s = mo->sig[e * mo->nq + q];
      d = mo->dsig[e * mo->nq + q];
    } else {
      double eps = 0.0; /* defaults.jl:8 */
      for (int i = 0; i < m; ++i) eps = (i == 0) ? B[0] * u[0] : eps + B[i] * u[i];
      s = orc_sigma(mo, eps);
      d = orc_dsigma(mo, eps);
    }
    double wq = t->wq[q];
    if (fe)
      for (int i = 0; i < m; ++i) { /* defaults.jl:9 then quadrature.jl:69 */
        double v = ((B[i] * s) * A) * dxdxi;
        fe[i] = (q == 0) ? v * wq : fe[i] + v * wq;
      }
    if (Ke)
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) { /* defaults.jl:23 */
          double v = (((B[i] * d) * B[j]) * A) * dxdxi;
          Ke[i * m + j] = (q == 0) ? v * wq : Ke[i * m + j] + v * wq;
        }
  }
}

/* linearelastoplasticity.jl:60-93  update_state_el! */
static void orc_update_state_el(orc_model *mo, const orc_tab1d *t, int64_t e, const double *u) {
  int m = t->m;
  double E = mo->mat[0], Hk = mo->mat[1], Ha = mo->mat[2], Eep = mo->mat[5];
  double du[ORC_MAXN], B[ORC_MAXN];
  for (int i = 0; i < m; ++i) du[i] = u[i] - mo->u_n[e * m + i];
  double dxdxi = orc_el_dxdxi(mo, e);
  double inv = 1.0 / dxdxi;
  for (int q = 0; q < t->nq; ++q) {
    int64_t k = e * mo->nq + q;
    for (int i = 0; i < m; ++i) B[i] = t->dN[q][i] * inv;
    double deps = 0.0;
    for (int i = 0; i < m; ++i) deps = (i == 0) ? B[0] * du[0] : deps + B[i] * du[i];
    double s_tr = mo->sig_n[k] + E * deps;
    double rel = s_tr - mo->alp_n[k];
    double sgn = (rel > 0.0) ? 1.0 : ((rel < 0.0) ? -1.0 : 0.0);
    double f_tr = fabs(rel) - mo->kap_n[k];
    if (f_tr > 1e-5) {
      double dg = f_tr / (E + Hk + Ha);
      mo->sig[k] = s_tr - E * dg * sgn;
      mo->alp[k] = mo->alp_n[k] + Ha * dg * sgn;
      mo->kap[k] = mo->kap_n[k] + Hk * dg;
      mo->gam[k] = mo->gam_n[k] + dg;
      mo->dsig[k] = Eep;
    } else {
      mo->sig[k] = s_tr;
      mo->alp[k] = mo->alp_n[k];
      mo->kap[k] = mo->kap_n[k];
      mo->gam[k] = mo->gam_n[k];
      mo->dsig[k] = E;
    }
  }
}

/* linearelastoplasticity.jl:106-117 */
void orc_init_state(orc_model *mo) {
  if (mo->mat_kind != ORC_MAT_PLASTIC) return;
  int64_t n = mo->nel * mo->nq;
  for (int64_t k = 0; k < n; ++k) {
    mo->sig[k] = mo->alp[k] = mo->kap[k] = mo->gam[k] = mo->dsig[k] = 0.0;
    mo->sig_n[k] = 0.0; mo->gam_n[k] = 0.0;
    mo->alp_n[k] = mo->mat[4];
    mo->kap_n[k] = mo->mat[3];
  }
  for (int64_t k = 0; k < mo->nel * mo->nen; ++k) mo->u_n[k] = 0.0;
}

/* linearelastoplasticity.jl:96-103 */
void orc_commit_state(orc_model *mo, const double *U) {
  if (mo->mat_kind != ORC_MAT_PLASTIC) return;
  for (int64_t e = 0; e < mo->nel; ++e) {
    for (int a = 0; a < mo->nen; ++a) mo->u_n[e * mo->nen + a] = U[mo->conn[e * mo->nen + a]];
    for (int q = 0; q < mo->nq; ++q) {
      int64_t k = e * mo->nq + q;
      mo->sig_n[k] = mo->sig[k]; mo->alp_n[k] = mo->alp[k];
      mo->kap_n[k] = mo->kap[k]; mo->gam_n[k] = mo->gam[k];
    }
  }
}

/* ------------------------------------------------------------------ */
/* T2 (builder-defined): Q1 tensor-product elements,     */
/* compressible Neo-Hookean. Same restrict -> qfunc -> integrate ->    */
/* unrestrict/assemble structure as defaults.jl:2-26.                  */
/* ------------------------------------------------------------------ */
typedef struct {
  int dim, nen, nqp;
  double w[8];
  double dN[8][8][3]; /* [qp][node][ref dir] */
} orc_tabq1;

static void orc_tabq1_init(const orc_model *mo, orc_tabq1 *t) {
  double nodes[2], bw[2], xq[ORC_MAXQ], wq[ORC_MAXQ];
  int dim = mo->dim;
  orc_lagrange_nodes(1, 0, nodes); /* p=1 chebyshev2 = [-1, +1] */
  orc_lagrange_weights(2, nodes, bw);
  orc_gauss(2, xq, wq);
  t->dim = dim; t->nen = 1 << dim; t->nqp = 1 << dim;
  double N1[2][2], D1[2][2]; /* [qp1d][node1d] */
  for (int q = 0; q < 2; ++q) { orc_shape_N(2, nodes, bw, xq[q], N1[q]); orc_shape_dN(2, nodes, bw, xq[q], D1[q]); }
  for (int q = 0; q < t->nqp; ++q) {
    int qi[3] = { q & 1, (q >> 1) & 1, (q >> 2) & 1 }; /* x fastest */
    double w = 1.0;
    for (int d = 0; d < dim; ++d) w = (d == 0) ? wq[qi[0]] : w * wq[qi[d]];
    t->w[q] = w;
    for (int a = 0; a < t->nen; ++a) {
      int ai[3] = { a & 1, (a >> 1) & 1, (a >> 2) & 1 };
      for (int D = 0; D < dim; ++D) {
        double v = 1.0;
        for (int d = 0; d < dim; ++d) {
          double f = (d == D) ? D1[qi[d]][ai[d]] : N1[qi[d]][ai[d]];
          v = (d == 0) ? f : v * f;
        }
        t->dN[q][a][D] = v;
      }
    }
  }
}

static double orc_inv(int dim, const double *M, double *Mi) { /* row-major dim x dim; returns det */
  if (dim == 2) {
    double det = M[0] * M[3] - M[1] * M[2];
    double id = 1.0 / det;
    Mi[0] = M[3] * id; Mi[1] = -M[1] * id; Mi[2] = -M[2] * id; Mi[3] = M[0] * id;
    return det;
  }
  double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  double id = 1.0 / det;
  Mi[0] = c00 * id; Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id; Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  Mi[3] = c01 * id; Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id; Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  Mi[6] = c02 * id; Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id; Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  return det;
}

/* ue: [nen][dim]; fe: [nen*dim]; Ke: [nen*dim][nen*dim] row-major */
static int orc_absmode = 0;   /* set only by orc_assemble_abs (test metric; one caller at a time): sums |contributions| */

static void orc_element_q1(const orc_model *mo, const orc_tabq1 *t, int64_t e, const double *ue,
                           double *fe, double *Ke) {
  int dim = t->dim, nen = t->nen, nd = nen * dim;
  double mu = mo->mat[0], lam = mo->mat[1];
  const int32_t *c = mo->conn + e * nen;
  if (fe) for (int i = 0; i < nd; ++i) fe[i] = 0.0;
  if (Ke) for (int i = 0; i < nd * nd; ++i) Ke[i] = 0.0;
  for (int q = 0; q < t->nqp; ++q) {
    double J[9] = {0}, Ji[9], F[9], Fi[9], G[8][3];
    for (int a = 0; a < nen; ++a)
      for (int d = 0; d < dim; ++d)
        for (int D = 0; D < dim; ++D) J[d * dim + D] += mo->coords[(int64_t)c[a] * dim + d] * t->dN[q][a][D];
    double detJ = orc_inv(dim, J, Ji);
    for (int a = 0; a < nen; ++a) /* dN/dX = dN/dxi * J^-1 */
      for (int D = 0; D < dim; ++D) {
        double s = 0.0;
        for (int E = 0; E < dim; ++E) s += t->dN[q][a][E] * Ji[E * dim + D];
        G[a][D] = s;
      }
    for (int i = 0; i < dim; ++i)
      for (int Jc = 0; Jc < dim; ++Jc) {
        double s = (i == Jc) ? 1.0 : 0.0;
        for (int a = 0; a < nen; ++a) s += ue[a * dim + i] * G[a][Jc];
        F[i * dim + Jc] = s;
      }
    double detF = orc_inv(dim, F, Fi);
    double lnJ = log(detF);
    double wd = t->w[q] * detJ;
    /* P_iJ = mu (F_iJ - Finv_Ji) + lam lnJ Finv_Ji */
    if (fe) {
      double P[9];
      for (int i = 0; i < dim; ++i)
        for (int Jc = 0; Jc < dim; ++Jc)
          P[i * dim + Jc] = mu * (F[i * dim + Jc] - Fi[Jc * dim + i]) + lam * lnJ * Fi[Jc * dim + i];
      for (int a = 0; a < nen; ++a)
        for (int i = 0; i < dim; ++i) {
          double s = 0.0;
          if (orc_absmode) {   /* every summand of the expression above by absolute value: mu F, mu F^-T and lam lnJ F^-T are
                                  added separately, and for small strains the first two cancel almost completely */
            for (int Jc = 0; Jc < dim; ++Jc)
              s += (fabs(mu * F[i * dim + Jc]) + fabs(mu * Fi[Jc * dim + i]) + fabs(lam * lnJ * Fi[Jc * dim + i])) * fabs(G[a][Jc]);
            fe[a * dim + i] += s * fabs(wd);
            continue;
          }
          for (int Jc = 0; Jc < dim; ++Jc) s += P[i * dim + Jc] * G[a][Jc];
          fe[a * dim + i] += s * wd;
        }
    }
    if (Ke) {
      /* A_iJkL = mu d_ik d_JL + lam Finv_Ji Finv_Lk + (mu - lam lnJ) Finv_Li Finv_Jk */
      double A4[3][3][3][3];
      double c3 = mu - lam * lnJ;
      for (int i = 0; i < dim; ++i) for (int Jc = 0; Jc < dim; ++Jc)
        for (int k = 0; k < dim; ++k) for (int L = 0; L < dim; ++L)
          A4[i][Jc][k][L] = ((i == k && Jc == L) ? mu : 0.0) + lam * Fi[Jc * dim + i] * Fi[L * dim + k]
                            + c3 * Fi[L * dim + i] * Fi[Jc * dim + k];
      /* T_{(a,i),(k,L)} = sum_J G_aJ A_iJkL ; K_{(a,i),(b,k)} += wd * sum_L T * G_bL */
      for (int a = 0; a < nen; ++a)
        for (int i = 0; i < dim; ++i) {
          double T[3][3];
          for (int k = 0; k < dim; ++k) for (int L = 0; L < dim; ++L) {
            double s = 0.0;
            for (int Jc = 0; Jc < dim; ++Jc) s += G[a][Jc] * A4[i][Jc][k][L];
            T[k][L] = s;
          }
          double *row = Ke + (int64_t)(a * dim + i) * nd;
          for (int b = 0; b < nen; ++b)
            for (int k = 0; k < dim; ++k) {
              double s = 0.0;
              for (int L = 0; L < dim; ++L) s += T[k][L] * G[b][L];
              row[b * dim + k] += orc_absmode ? fabs(s * wd) : s * wd;
            }
        }
    }
  }
}

/* one element through the public surface (no state update) */
void orc_element(const orc_model *mo, int64_t e, const double *ue, double *fe, double *Ke) {
  if (mo->dim == 1) { orc_tab1d t; orc_tab1d_init(mo, &t); orc_element_1d(mo, &t, e, ue, fe, Ke); }
  else { orc_tabq1 t; orc_tabq1_init(mo, &t); orc_element_q1(mo, &t, e, ue, fe, Ke); }
}

/* ------------------------------------------------------------------ */
/* Boundaries: src/fem/boundary.jl:29,39; src/fem/fem.jl:75-90         */
/* ------------------------------------------------------------------ */
static void orc_apply_dirichlet(const orc_model *mo, double *U) {
  for (int64_t k = 0; k < mo->ndir; ++k) U[mo->dir_dofs[k]] = mo->dir_vals[k];
}
static int64_t orc_ndof(const orc_model *mo) { return mo->nn * mo->dim; }

/* ------------------------------------------------------------------ */
/* Global loops: defaults.jl:29-46, 68-82; residual.jl:12-26;          */
/* jacobian.jl:10-22. `sink` abstracts dense vs CSR K storage.         */
/* ------------------------------------------------------------------ */
typedef struct {
  double *dense; int64_t ndof;
  const int64_t *rowptr; const int32_t *colind; double *vals;
  int atomic;
} orc_sink;

static inline void orc_sink_add(const orc_sink *s, int64_t gi, int64_t gj, double v) {
  double *p;
  if (s->dense) p = s->dense + gi * s->ndof + gj;
  else {
    int64_t lo = s->rowptr[gi], hi = s->rowptr[gi + 1] - 1;
    int32_t key = (int32_t)gj;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (s->colind[mid] < key) lo = mid + 1; else hi = mid; }
    p = s->vals + lo;
  }
  if (s->atomic) orc_atomic_add(p, v); else *p += v;
}

/* DistributedForce: distributedforce.jl:7-27. 1-D: f_e = integrate(Q, xi -> J * fdist * N(P, xi)), J = length(e)/2;
 * Q1 (builder-defined): f_{a,i} = sum_q w_q det(dX/dxi) N_a f_i. Fext[nodes] += f_e in element order. */
static void orc_body_force(const orc_model *mo, double *Fext) {
  int dim = mo->dim, nen = mo->nen;
  if (dim == 1) {
    double nodes[ORC_MAXN], w[ORC_MAXN], xq[ORC_MAXQ], wq[ORC_MAXQ], N[ORC_MAXQ][ORC_MAXN];
    orc_lagrange_nodes(mo->p, 0, nodes); orc_lagrange_weights(nen, nodes, w); orc_gauss(mo->nq, xq, wq);
    for (int q = 0; q < mo->nq; ++q) orc_shape_N(nen, nodes, w, xq[q], N[q]);
    for (int64_t e = 0; e < mo->nel; ++e) {
      double J = orc_el_dxdxi(mo, e);
      for (int a = 0; a < nen; ++a) {
        double s = 0.0;
        for (int q = 0; q < mo->nq; ++q) { double v = (J * mo->body[0] * N[q][a]) * wq[q]; s = (q == 0) ? v : s + v; }
        Fext[mo->conn[e * nen + a]] += s;
      }
    }
    return;
  }
  orc_tabq1 t; orc_tabq1_init(mo, &t);
  double nodes[2], bw[2], xq[ORC_MAXQ], wq[ORC_MAXQ], N1[2][2];
  orc_lagrange_nodes(1, 0, nodes); orc_lagrange_weights(2, nodes, bw); orc_gauss(2, xq, wq);
  for (int q = 0; q < 2; ++q) orc_shape_N(2, nodes, bw, xq[q], N1[q]);
  for (int64_t e = 0; e < mo->nel; ++e) {
    const int32_t *c = mo->conn + e * nen;
    for (int q = 0; q < t.nqp; ++q) {
      double J[9] = {0}, Ji[9];
      for (int a = 0; a < nen; ++a)
        for (int d = 0; d < dim; ++d)
          for (int D = 0; D < dim; ++D) J[d * dim + D] += mo->coords[(int64_t)c[a] * dim + d] * t.dN[q][a][D];
      double wd = t.w[q] * orc_inv(dim, J, Ji);
      for (int a = 0; a < nen; ++a) {
        double Na = 1.0;
        for (int d = 0; d < dim; ++d) Na = (d == 0) ? N1[(q >> d) & 1][(a >> d) & 1] : Na * N1[(q >> d) & 1][(a >> d) & 1];
        for (int d = 0; d < dim; ++d) Fext[(int64_t)c[a] * dim + d] += wd * Na * mo->body[d];
      }
    }
  }
}

/* Consistent mass: gallery/mass.jl:2-17 (assemblemass!) and :19-37 (applymass!).
 * 1-D: m_e = integrate(Q, xi -> N N' * A * rho * J), J = length(e)/2, product order ((N_a N_b) A) rho) J, then * w_q,
 * summed over q ascending; M[nodes,nodes] += m_e in element order. `area` is the caller's fem.data[:A] for the
 * mass (mass.jl:9,27 default 1.0 -- not the 3e-4 default of the internal force).
 * Q1 (builder-defined, same structure): m_ab = sum_q ((N_a N_b) rho) (w_q det(dX/dxi)), M_{ai,bk} = delta_ik m_ab.
 * Output: node-level dense matrix Mn[nn][nn] (the dim x dim expansion is the caller's). */
void orc_mass_nodes(const orc_model *mo, double rho, double area, double *Mn) {
  int dim = mo->dim, nen = mo->nen;
  int64_t nn = mo->nn;
  for (int64_t i = 0; i < nn * nn; ++i) Mn[i] = 0.0;
  if (dim == 1) {
    double nodes[ORC_MAXN], w[ORC_MAXN], xq[ORC_MAXQ], wq[ORC_MAXQ], N[ORC_MAXQ][ORC_MAXN];
    orc_lagrange_nodes(mo->p, 0, nodes); orc_lagrange_weights(nen, nodes, w); orc_gauss(mo->nq, xq, wq);
    for (int q = 0; q < mo->nq; ++q) orc_shape_N(nen, nodes, w, xq[q], N[q]);
    for (int64_t e = 0; e < mo->nel; ++e) {
      double J = orc_el_dxdxi(mo, e);
      for (int a = 0; a < nen; ++a)
        for (int b = 0; b < nen; ++b) {
          double s = 0.0;
          for (int q = 0; q < mo->nq; ++q) { double v = ((((N[q][a] * N[q][b]) * area) * rho) * J) * wq[q]; s = (q == 0) ? v : s + v; }
          Mn[(int64_t)mo->conn[e * nen + a] * nn + mo->conn[e * nen + b]] += s;
        }
    }
    return;
  }
  orc_tabq1 t; orc_tabq1_init(mo, &t);
  double nodes[2], bw[2], xq[ORC_MAXQ], wq[ORC_MAXQ], N1[2][2];
  orc_lagrange_nodes(1, 0, nodes); orc_lagrange_weights(2, nodes, bw); orc_gauss(2, xq, wq);
  for (int q = 0; q < 2; ++q) orc_shape_N(2, nodes, bw, xq[q], N1[q]);
  for (int64_t e = 0; e < mo->nel; ++e) {
    const int32_t *c = mo->conn + e * nen;
    double me[8][8];
    for (int a = 0; a < nen; ++a) for (int b = 0; b < nen; ++b) me[a][b] = 0.0;
    for (int q = 0; q < t.nqp; ++q) {
      double J[9] = {0}, Ji[9], Nq[8];
      for (int a = 0; a < nen; ++a)
        for (int d = 0; d < dim; ++d)
          for (int D = 0; D < dim; ++D) J[d * dim + D] += mo->coords[(int64_t)c[a] * dim + d] * t.dN[q][a][D];
      double wd = t.w[q] * orc_inv(dim, J, Ji);
      for (int a = 0; a < nen; ++a) {
        double Na = 1.0;
        for (int d = 0; d < dim; ++d) Na = (d == 0) ? N1[(q >> d) & 1][(a >> d) & 1] : Na * N1[(q >> d) & 1][(a >> d) & 1];
        Nq[a] = Na;
      }
      for (int a = 0; a < nen; ++a) for (int b = 0; b < nen; ++b) me[a][b] += ((Nq[a] * Nq[b]) * rho) * wd;
    }
    for (int a = 0; a < nen; ++a) for (int b = 0; b < nen; ++b) Mn[(int64_t)c[a] * nn + c[b]] += me[a][b];
  }
}

typedef struct { orc_model *mo; double *U, *R; const orc_sink *K; const orc_tab1d *t1; const orc_tabq1 *tq; int nth; int absmode; } orc_asm_ctx;

static void orc_assemble_range(void *vc, int64_t lo, int64_t hi, int tid) {
  orc_asm_ctx *c_ = (orc_asm_ctx *)vc; (void)tid;
  orc_model *mo = c_->mo; double *U = c_->U, *R = c_->R; const orc_sink *K = c_->K;
  int dim = mo->dim, nen = mo->nen, nd = nen * dim, nth = c_->nth;
  const int ab = c_->absmode;                      /* test metric: sum of |element contributions| per entry */
  for (int64_t e = lo; e < hi; ++e) {              /* defaults.jl:38 / :77 */
    double ue[24], fe[24], Ke[24 * 24];
    const int32_t *c = mo->conn + e * nen;
    for (int a = 0; a < nen; ++a)                  /* restrict!, element.jl:64 */
      for (int d = 0; d < dim; ++d) ue[a * dim + d] = U[(int64_t)c[a] * dim + d];
    if (dim == 1) {
      if (R && mo->mat_kind == ORC_MAT_PLASTIC) orc_update_state_el(mo, c_->t1, e, ue); /* defaults.jl:40-42 */
      orc_element_1d(mo, c_->t1, e, ue, R ? fe : NULL, K ? Ke : NULL);
    } else {
      orc_element_q1(mo, c_->tq, e, ue, R ? fe : NULL, K ? Ke : NULL);
    }
    if (R)
      for (int a = 0; a < nen; ++a)                /* unrestrict!, element.jl:67 */
        for (int d = 0; d < dim; ++d) {
          double *p = R + (int64_t)c[a] * dim + d;
          const double v = ab ? fabs(fe[a * dim + d]) : fe[a * dim + d];
          if (nth > 1) orc_atomic_add(p, v); else *p += v;
        }
    if (K)
      for (int i = 0; i < nd; ++i)                 /* assemble!, element.jl:70 */
        for (int j = 0; j < nd; ++j)
          orc_sink_add(K, (int64_t)c[i / dim] * dim + i % dim, (int64_t)c[j / dim] * dim + j % dim, ab ? fabs(Ke[i * nd + j]) : Ke[i * nd + j]);
  }
}

static void orc_assemble(orc_model *mo, double *U, double *R, const orc_sink *K) {
  int dim = mo->dim;
  int64_t ndof = orc_ndof(mo);
  int nth = (mo->nthreads > 1 && !(K && K->dense)) ? mo->nthreads : 1;
  orc_apply_dirichlet(mo, U);                     /* residual.jl:13 / jacobian.jl:11 */
  if (R) memset(R, 0, sizeof(double) * ndof);     /* residual.jl:14 */
  orc_tab1d t1; orc_tabq1 tq;
  if (dim == 1) orc_tab1d_init(mo, &t1); else orc_tabq1_init(mo, &tq);
  orc_asm_ctx ctx = { mo, U, R, K, &t1, &tq, nth, orc_absmode };
  orc_parallel_for(nth, mo->nel, orc_assemble_range, &ctx);
  if (R && !orc_absmode) {                         /* residual.jl:15,24-25 */
    double *Fext = (double *)calloc((size_t)ndof, sizeof(double));
    if (mo->has_body) orc_body_force(mo, Fext);     /* residual.jl:19-23 */
    for (int64_t k = 0; k < mo->nneu; ++k) Fext[mo->neu_dofs[k]] += mo->neu_vals[k];
    for (int64_t i = 0; i < ndof; ++i) R[i] -= Fext[i];
    free(Fext);
  }
}

void orc_residual(orc_model *mo, double *U, double *R) { orc_assemble(mo, U, R, NULL); }

void orc_jacobian_dense(orc_model *mo, double *U, double *K) {
  int64_t n = orc_ndof(mo);
  memset(K, 0, sizeof(double) * n * n);            /* jacobian.jl:13 */
  orc_sink s = { K, n, NULL, NULL, NULL, 0 };
  orc_assemble(mo, U, NULL, &s);
}

void orc_jacobian_csr(orc_model *mo, double *U, const int64_t *rowptr, const int32_t *colind, double *vals) {
  int64_t n = orc_ndof(mo);
  memset(vals, 0, sizeof(double) * rowptr[n]);
  orc_sink s = { NULL, n, rowptr, colind, vals, mo->nthreads > 1 };
  orc_assemble(mo, U, NULL, &s);
}

/* Test metric, not a reference function: per entry of R and of the CSR Jacobian the sum of the ABSOLUTE contributions
   (Q1: per quadrature point and element, R: per summand of the stress expression; 1-D: per element) -- the scale an entry-wise relative error is judged against (entries that cancel to ~0 are made of
   contributions of this size). Not for stateful materials (the element pass runs twice). */
void orc_assemble_abs(orc_model *mo, double *U, double *Rabs, const int64_t *rowptr, const int32_t *colind, double *vabs) {
  int64_t n = orc_ndof(mo);
  orc_absmode = 1;
  if (Rabs) orc_assemble(mo, U, Rabs, NULL);
  if (vabs) {
    memset(vabs, 0, sizeof(double) * rowptr[n]);
    orc_sink s = { NULL, n, rowptr, colind, vabs, mo->nthreads > 1 };
    orc_assemble(mo, U, NULL, &s);
  }
  orc_absmode = 0;
}

/* ------------------------------------------------------------------ */
/* Pattern: {(i,j): exists e, i in dofs(e) and j in dofs(e)}           */
/* (= the entries element.jl:70 can touch), rows/cols ascending.       */
/* ------------------------------------------------------------------ */
static int orc_cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* node -> element CSR (counting sort) */
static void orc_node_elems(const orc_model *mo, int64_t **ptr_out, int64_t **idx_out) {
  int64_t nn = mo->nn, nel = mo->nel; int nen = mo->nen;
  int64_t *ptr = (int64_t *)calloc((size_t)nn + 1, sizeof(int64_t));
  for (int64_t k = 0; k < nel * nen; ++k) ptr[mo->conn[k] + 1]++;
  for (int64_t i = 0; i < nn; ++i) ptr[i + 1] += ptr[i];
  int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ptr[nn] ? ptr[nn] : 1));
  int64_t *fill = (int64_t *)calloc((size_t)nn, sizeof(int64_t));
  for (int64_t e = 0; e < nel; ++e)
    for (int a = 0; a < nen; ++a) { int32_t n = mo->conn[e * nen + a]; idx[ptr[n] + fill[n]++] = e; }
  free(fill);
  *ptr_out = ptr; *idx_out = idx;
}

int64_t orc_pattern(const orc_model *mo, int64_t *rowptr, int32_t *colind) {
  int64_t nn = mo->nn; int nen = mo->nen, dim = mo->dim;
  int64_t *nptr, *nidx;
  orc_node_elems(mo, &nptr, &nidx);
  int32_t *buf = NULL; int64_t cap = 0;
  int64_t nnz = 0;
  rowptr[0] = 0;
  for (int64_t n = 0; n < nn; ++n) {
    int64_t cnt = (nptr[n + 1] - nptr[n]) * nen;
    if (cnt > cap) { cap = cnt * 2; buf = (int32_t *)realloc(buf, sizeof(int32_t) * (size_t)cap); }
    int64_t k = 0;
    for (int64_t p = nptr[n]; p < nptr[n + 1]; ++p)
      for (int a = 0; a < nen; ++a) buf[k++] = mo->conn[nidx[p] * nen + a];
    qsort(buf, (size_t)k, sizeof(int32_t), orc_cmp_i32);
    int64_t u = 0;
    for (int64_t i = 0; i < k; ++i) if (i == 0 || buf[i] != buf[i - 1]) buf[u++] = buf[i];
    for (int c = 0; c < dim; ++c) {
      if (colind)
        for (int64_t i = 0; i < u; ++i)
          for (int d = 0; d < dim; ++d) colind[nnz + i * dim + d] = buf[i] * dim + d;
      nnz += u * dim;
      rowptr[n * dim + c + 1] = nnz;
    }
  }
  free(buf); free(nptr); free(nidx);
  return nnz;
}

/* ------------------------------------------------------------------ */
/* Partition: contiguous owned node ranges; a rank      */
/* computes every element touching an owned node; ghosts = the rest    */
/* of those elements' nodes. Integer work, bit-exact contract.         */
/* ------------------------------------------------------------------ */
static int orc_cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}
static int32_t orc_owner(int32_t nranks, const int64_t *off, int64_t node) {
  int32_t lo = 0, hi = nranks - 1;
  while (lo < hi) { int32_t mid = (lo + hi + 1) >> 1; if (off[mid] <= node) lo = mid; else hi = mid - 1; }
  return lo;
}

static int64_t orc_ghosts_of(const orc_model *mo, const int64_t *off, int32_t rank,
                             int64_t *n_loc, int64_t *loc_elems, int64_t **ghost_out) {
  int nen = mo->nen;
  int64_t lo = off[rank], hi = off[rank + 1];
  int64_t ne = 0, ng = 0, cap = 1024;
  int64_t *g = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
  for (int64_t e = 0; e < mo->nel; ++e) {
    int touch = 0;
    for (int a = 0; a < nen; ++a) { int64_t n = mo->conn[e * nen + a]; if (n >= lo && n < hi) touch = 1; }
    if (!touch) continue;
    if (loc_elems) loc_elems[ne] = e;
    ne++;
    for (int a = 0; a < nen; ++a) {
      int64_t n = mo->conn[e * nen + a];
      if (n >= lo && n < hi) continue;
      if (ng == cap) { cap *= 2; g = (int64_t *)realloc(g, sizeof(int64_t) * (size_t)cap); }
      g[ng++] = n;
    }
  }
  qsort(g, (size_t)ng, sizeof(int64_t), orc_cmp_i64);
  int64_t u = 0;
  for (int64_t i = 0; i < ng; ++i) if (i == 0 || g[i] != g[i - 1]) g[u++] = g[i];
  *n_loc = ne; *ghost_out = g;
  return u;
}

int32_t orc_partition(const orc_model *mo, int32_t nranks, const int64_t *off, int32_t rank,
                      int64_t *n_loc_elems, int64_t *loc_elems, int64_t *n_ghost, int64_t *ghost_nodes,
                      int64_t *send_ptr, int64_t *send_nodes) {
  int64_t *g;
  int64_t ng = orc_ghosts_of(mo, off, rank, n_loc_elems, loc_elems, &g);
  *n_ghost = ng;
  if (ghost_nodes) memcpy(ghost_nodes, g, sizeof(int64_t) * (size_t)ng);
  free(g);
  /* send lists: my owned nodes that appear as ghosts on rank q */
  int64_t pos = 0;
  send_ptr[0] = 0;
  for (int32_t q = 0; q < nranks; ++q) {
    if (q != rank) {
      int64_t nq_loc, *gq;
      int64_t ngq = orc_ghosts_of(mo, off, q, &nq_loc, NULL, &gq);
      for (int64_t i = 0; i < ngq; ++i)
        if (orc_owner(nranks, off, gq[i]) == rank) { if (send_nodes) send_nodes[pos] = gq[i]; pos++; }
      free(gq);
    }
    send_ptr[q + 1] = pos;
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* Linear algebra                                                      */
/* ------------------------------------------------------------------ */

/* Dense LU with partial pivoting (what `\` does for a square Matrix{Float64}:
 * LAPACK getrf/getrs, newton_fem.jl:158). Returns 1 on an exactly zero pivot
 * (SingularException). */
int32_t orc_dense_solve(int64_t n, double *A, double *b) {
  for (int64_t k = 0; k < n; ++k) {
    int64_t piv = k; double mx = fabs(A[k * n + k]);
    for (int64_t i = k + 1; i < n; ++i) { double v = fabs(A[i * n + k]); if (v > mx) { mx = v; piv = i; } }
    if (mx == 0.0 || mx != mx) return 1;
    if (piv != k) {
      for (int64_t j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
      double t = b[k]; b[k] = b[piv]; b[piv] = t;
    }
    double inv = 1.0 / A[k * n + k];
    for (int64_t i = k + 1; i < n; ++i) {
      double l = A[i * n + k] * inv;
      if (l == 0.0) continue;
      A[i * n + k] = l;
      for (int64_t j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
      b[i] -= l * b[k];
    }
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int64_t j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j];
    b[i] = s / A[i * n + i];
  }
  return 0;
}

typedef struct {
  const int64_t *rowptr; const int32_t *colind; const double *vals; const uint8_t *con;
  const double *x; double *y;                 /* spmv */
  double *xx, *r, *z, *p, *Ap; const double *dinv; double alpha, beta; /* pcg */
  double part[256][4];
} orc_la_ctx;

static void orc_spmv_range(void *vc, int64_t lo, int64_t hi, int tid) {
  orc_la_ctx *c = (orc_la_ctx *)vc;
  double pAp = 0.0;
  for (int64_t i = lo; i < hi; ++i) {
    double s = 0.0;
    for (int64_t k = c->rowptr[i]; k < c->rowptr[i + 1]; ++k) s += c->vals[k] * c->x[c->colind[k]];
    if (c->con && c->con[i]) s = 0.0;
    c->y[i] = s;
    pAp += c->x[i] * s;
  }
  c->part[tid][0] = pAp;
}
static void orc_update_range(void *vc, int64_t lo, int64_t hi, int tid) {
  orc_la_ctx *c = (orc_la_ctx *)vc;
  double rz = 0.0, rr = 0.0;
  for (int64_t i = lo; i < hi; ++i) {
    c->xx[i] += c->alpha * c->p[i];
    c->r[i] -= c->alpha * c->Ap[i];
    c->z[i] = c->dinv[i] * c->r[i];
    rz += c->r[i] * c->z[i]; rr += c->r[i] * c->r[i];
  }
  c->part[tid][0] = rz; c->part[tid][1] = rr;
}
static void orc_pupdate_range(void *vc, int64_t lo, int64_t hi, int tid) {
  orc_la_ctx *c = (orc_la_ctx *)vc; (void)tid;
  for (int64_t i = lo; i < hi; ++i) c->p[i] = c->z[i] + c->beta * c->p[i];
}
static int orc_nchunks(int nth, int64_t n) { return (nth <= 1 || n < 2 * nth) ? 1 : (nth > 256 ? 256 : nth); }

void orc_spmv(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals,
              const double *x, double *y, int32_t nthreads) {
  orc_la_ctx *c = (orc_la_ctx *)calloc(1, sizeof(orc_la_ctx));
  c->rowptr = rowptr; c->colind = colind; c->vals = vals; c->x = x; c->y = y;
  orc_parallel_for(nthreads, n, orc_spmv_range, c);
  free(c);
}

/* Jacobi-preconditioned CG on the free-free block: constrained rows/cols act as
 * identity (x_c = 0). Builder-defined (the reference uses dense LU,
 * newton_fem.jl:158). Converged when ||r|| <= rtol ||b||.
 * Returns 0 ok, 1 maxits, 2 breakdown. */
int32_t orc_pcg(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals,
                const uint8_t *con, const double *b, double *x, double rtol, int32_t maxits,
                double *relres, int32_t *its_out, int32_t nthreads) {
  int nth = nthreads > 1 ? nthreads : 1;
  double *r = (double *)malloc(sizeof(double) * n), *z = (double *)malloc(sizeof(double) * n);
  double *p = (double *)malloc(sizeof(double) * n), *Ap = (double *)malloc(sizeof(double) * n);
  double *dinv = (double *)malloc(sizeof(double) * n);
  orc_la_ctx *c = (orc_la_ctx *)calloc(1, sizeof(orc_la_ctx));
  c->rowptr = rowptr; c->colind = colind; c->vals = vals; c->con = con;
  c->xx = x; c->r = r; c->z = z; c->p = p; c->Ap = Ap; c->dinv = dinv;
  int32_t it = 0, status = 1;
  double bb = 0.0, rz = 0.0, rr = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) if (colind[k] == i) d = vals[k];
    dinv[i] = (con && con[i]) ? 0.0 : 1.0 / d;
    x[i] = 0.0;
    r[i] = (con && con[i]) ? 0.0 : b[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += r[i] * r[i]; rz += r[i] * z[i];
  }
  rr = bb;
  if (bb == 0.0) status = 0;
  while (status == 1 && it < maxits) {
    int nc = orc_nchunks(nth, n);
    c->x = p; c->y = Ap;
    orc_parallel_for(nth, n, orc_spmv_range, c);
    double pAp = 0.0;
    for (int t = 0; t < nc; ++t) pAp += c->part[t][0];
    if (!(pAp > 0.0)) { status = 2; break; }
    c->alpha = rz / pAp;
    orc_parallel_for(nth, n, orc_update_range, c);
    double rz_new = 0.0; rr = 0.0;
    for (int t = 0; t < nc; ++t) { rz_new += c->part[t][0]; rr += c->part[t][1]; }
    ++it;
    if (!(rr == rr)) { status = 2; break; }
    if (sqrt(rr) <= rtol * sqrt(bb)) { status = 0; break; }
    c->beta = rz_new / rz;
    rz = rz_new;
    orc_parallel_for(nth, n, orc_pupdate_range, c);
  }
  if (relres) *relres = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
  if (its_out) *its_out = it;
  free(r); free(z); free(p); free(Ap); free(dinv); free(c);
  return status;
}

/* ------------------------------------------------------------------ */
/* Newton: src/nonlinearsolvers/newton_fem.jl:132-190                  */
/* ------------------------------------------------------------------ */
static void orc_con_mask(const orc_model *mo, uint8_t *con) { /* fem.jl:59-64 constrained_nodes */
  memset(con, 0, (size_t)orc_ndof(mo));
  for (int64_t k = 0; k < mo->ndir; ++k) con[mo->dir_dofs[k]] = 1;
}
static double orc_norm_free(int64_t n, const uint8_t *con, const double *v) { /* norm(dofs(fem, R)) */
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) if (!con[i]) s += v[i] * v[i];
  return sqrt(s);
}

typedef int (*orc_linsolve_fn)(orc_model *mo, void *ctx, double *U, const double *R, const uint8_t *con,
                               double *dU, int refresh_K, int32_t *lin_its);

static void orc_newton_loop(orc_model *mo, const orc_newton_opts *o, orc_linsolve_fn lin, void *ctx,
                            const double *U0, double *U, double *R, double *hist, orc_newton_result *res) {
  int64_t n = orc_ndof(mo);
  uint8_t *con = (uint8_t *)malloc((size_t)n);
  double *dU = (double *)malloc(sizeof(double) * n);
  orc_con_mask(mo, con);
  memcpy(U, U0, sizeof(double) * n);                 /* :135 */
  orc_residual(mo, U, R);                            /* :137 */
  int refresh = 1;                                   /* :138 jacobian evaluated at the initial U */
  double norm_r = orc_norm_free(n, con, R);          /* :140-141 */
  double norm_r0 = norm_r > 2.220446049250313e-16 ? norm_r : 1.0; /* :143 */
  int32_t its = 0, lin_total = 0, nh = 0;
  res->reason = ORC_REASON_NOT_YET;
  if (hist) hist[nh++] = norm_r;
  while (its < o->maxits) {                          /* :145 */
    if (norm_r / norm_r0 < o->rtol) { res->reason = ORC_REASON_CONVERGED_RELATIVE; break; } /* :146-150 */
    if (norm_r < o->atol) { res->reason = ORC_REASON_CONVERGED_ABSOLUTE; break; }           /* :151-155 */
    int32_t li = 0;
    if (lin(mo, ctx, U, R, con, dU, refresh, &li)) { /* :157-167 */
      res->reason = ORC_REASON_DIVERGED_LINEAR_SOLVE;
      goto out;
    }
    lin_total += li;
    double nd = 0.0;
    for (int64_t i = 0; i < n; ++i) nd += dU[i] * dU[i];
    if (sqrt(nd) > o->dtol) { res->reason = ORC_REASON_DIVERGED_STEP; goto out; } /* :168-172 */
    for (int64_t i = 0; i < n; ++i) U[i] += dU[i];   /* :173 */
    orc_residual(mo, U, R);                          /* :175 */
    refresh = !o->type_modified;                     /* :176-178 */
    norm_r = orc_norm_free(n, con, R);               /* :180 */
    its++;                                           /* :181 */
    if (hist) hist[nh++] = norm_r;
  }
  if (its >= o->maxits) res->reason = ORC_REASON_DIVERGED_MAXITS; /* :184-187 */
out:
  res->iterations = its; res->lin_its_total = lin_total; res->nhist = nh;
  res->norm_r = norm_r; res->norm_r0 = norm_r0;
  free(con); free(dU);
}

typedef struct { double *K; double *Kff; int64_t nfree; int have; } orc_dense_ctx;

static int orc_lin_dense(orc_model *mo, void *vctx, double *U, const double *R, const uint8_t *con,
                         double *dU, int refresh, int32_t *lin_its) {
  orc_dense_ctx *c = (orc_dense_ctx *)vctx;
  int64_t n = orc_ndof(mo);
  if (refresh || !c->have) { orc_jacobian_dense(mo, U, c->K); c->have = 1; }
  int64_t nf = 0;
  for (int64_t i = 0; i < n; ++i) nf += !con[i];
  double *A = (double *)malloc(sizeof(double) * (size_t)(nf > 0 ? nf * nf : 1));
  double *b = (double *)malloc(sizeof(double) * (size_t)(nf ? nf : 1));
  int64_t r = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (con[i]) continue;
    int64_t cidx = 0;
    for (int64_t j = 0; j < n; ++j) if (!con[j]) A[r * nf + cidx++] = c->K[i * n + j];
    b[r++] = -R[i];
  }
  int fail = orc_dense_solve(nf, A, b);
  r = 0;
  for (int64_t i = 0; i < n; ++i) dU[i] = con[i] ? 0.0 : b[r++]; /* expand(), fem.jl:98-102 */
  free(A); free(b);
  *lin_its = 1;
  return fail;
}

void orc_newton_dense(orc_model *mo, const orc_newton_opts *o, const double *U0, double *U, double *R,
                      double *hist, orc_newton_result *res) {
  int64_t n = orc_ndof(mo);
  orc_dense_ctx c = { (double *)malloc(sizeof(double) * n * n), NULL, 0, 0 };
  orc_newton_loop(mo, o, orc_lin_dense, &c, U0, U, R, hist, res);
  free(c.K);
}

typedef struct { const int64_t *rowptr; const int32_t *colind; double *vals; const orc_newton_opts *o; int have; } orc_csr_ctx;

static int orc_lin_pcg(orc_model *mo, void *vctx, double *U, const double *R, const uint8_t *con,
                       double *dU, int refresh, int32_t *lin_its) {
  orc_csr_ctx *c = (orc_csr_ctx *)vctx;
  int64_t n = orc_ndof(mo);
  if (refresh || !c->have) { orc_jacobian_csr(mo, U, c->rowptr, c->colind, c->vals); c->have = 1; }
  double *b = (double *)malloc(sizeof(double) * n);
  for (int64_t i = 0; i < n; ++i) b[i] = -R[i];
  double relres;
  int32_t rc = orc_pcg(n, c->rowptr, c->colind, c->vals, con, b, dU, c->o->lin_rtol, c->o->lin_maxits,
                       &relres, lin_its, mo->nthreads);
  free(b);
  return rc != 0;
}

void orc_newton_pcg(orc_model *mo, const orc_newton_opts *o, const int64_t *rowptr, const int32_t *colind,
                    double *vals, const double *U0, double *U, double *R, double *hist, orc_newton_result *res) {
  orc_csr_ctx c = { rowptr, colind, vals, o, 0 };
  orc_newton_loop(mo, o, orc_lin_pcg, &c, U0, U, R, hist, res);
}
