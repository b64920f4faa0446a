/*
 * nls_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the Newton hot path of NonlinearSolids.jl
 * (reference: /root/reference, pure Julia; `julia` is not installed in this
 * image, so the reference itself cannot be executed here -- parity of the
 * assembled residual/Jacobian/Newton path is "unpinned" by reference-owned
 * tests; the building blocks (shape functions, quadrature, boundary apply) are
 * pinned against the assertions of test/runtests.jl:54-207, and the path as a
 * whole against the known answers KAT-1..5 derived from the reference formulas).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product library
 * (libnls_b200.so) never links, loads or calls it.
 *
 * T1 (reference parity): 1-D bar, Lagrange order p on Chebyshev-2 nodes, Gauss
 *     1..5 points, the three reference materials, dense K + dense LU Newton.
 * T2 (builder-defined extension): 2-D plane-strain /
 *     3-D Q1 compressible Neo-Hookean, CSR Jacobian, Jacobi-PCG.
 * Dynamics and continuation: orc_mass_nodes here; the Newmark stepper and the arc-length
 *     continuation are restated in oracle.py on top of these primitives with dense linear algebra (small cases).
 *     The reference has no tests for mass.jl, newmark.jl or newtonarclength_fem.jl, and the last two are not
 *     runnable as shipped, so their parity is unpinned too: the restatements are pinned by known answers (element
 *     mass rho A L/6 [2 1;1 2], total mass, energy conservation of the average-acceleration rule, equilibrium of the
 *     converged arc-length states), tests/test_oracle.py.
 */
#ifndef NLS_ORACLE_H
#define NLS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MAT_LINEAR = 0, ORC_MAT_EXPONENTIAL = 1, ORC_MAT_PLASTIC = 2, ORC_MAT_NEOHOOKEAN = 3 };

enum {
  ORC_REASON_NOT_YET = 0,
  ORC_REASON_CONVERGED_RELATIVE = 1,
  ORC_REASON_CONVERGED_ABSOLUTE = 2,
  ORC_REASON_DIVERGED_MAXITS = 3,
  ORC_REASON_DIVERGED_STEP = 4,
  ORC_REASON_DIVERGED_LINEAR_SOLVE = 5
};

/* A finite-element model: the data FEMModel + Mesh + material + boundaries hold
 * (src/fem/fem.jl:9-24, src/fem/element.jl:73-83). All indices 0-based. */
typedef struct {
  int32_t dim;          /* 1, 2 or 3; DoFs per node = dim, gdof = dim*node + c */
  int32_t nen;          /* nodes per element: p+1 (1-D), 4 (2-D Q1), 8 (3-D Q1) */
  int64_t nn;           /* nodes */
  int64_t nel;          /* elements */
  const int32_t *conn;  /* [nel][nen] */
  const double *coords; /* [nn][dim] */
  int32_t p;            /* Lagrange order per direction (1-D: any; multi-D: 1) */
  int32_t nq;           /* Gauss points per direction */
  int32_t mat_kind;
  double mat[8];        /* LINEAR: E | EXP: sigma_sat, b | PLASTIC: E,Hk,Ha,k0,a0,Eep | NH: mu, lambda */
  double area;          /* fem.data[:A] (1-D only; defaults.jl:3 default 3e-4) */
  int64_t ndir;         /* Dirichlet list, applied in order (fem.jl:75-81) */
  const int64_t *dir_dofs;
  const double *dir_vals;
  int64_t nneu;         /* Neumann list, added in order (fem.jl:84-90) */
  const int64_t *neu_dofs;
  const double *neu_vals;
  /* per-(element,qp) plastic state (linearelastoplasticity.jl:60-117); may be NULL otherwise */
  double *sig, *alp, *kap, *gam, *dsig;
  double *sig_n, *alp_n, *kap_n, *gam_n;
  double *u_n;          /* [nel][nen] */
  int32_t nthreads;     /* worker threads (pthreads) for the scalable (CSR) paths; 1 = serial & deterministic */
  int32_t has_body;     /* DistributedForce (gallery/distributedforce.jl): density per component */
  double body[3];
} orc_model;

typedef struct {
  int32_t type_modified; /* 0 :standard, 1 :modified (newton_fem.jl:7,176) */
  double atol, rtol, dtol;
  int32_t maxits;
  double lin_rtol;       /* PCG variant only */
  int32_t lin_maxits;
} orc_newton_opts;

typedef struct {
  int32_t reason;
  int32_t iterations;
  int32_t lin_its_total;
  int32_t nhist;          /* number of entries written to norm_hist (<= maxits+1) */
  double norm_r, norm_r0;
} orc_newton_result;

/* ---- building blocks (shapefunctions.jl, quadrature.jl) ---- */
void orc_lagrange_nodes(int32_t p, int32_t kind /*0 :chebyshev2, 1 :equidistant*/, double *nodes);
void orc_lagrange_weights(int32_t m, const double *nodes, double *w);
void orc_shape_N(int32_t m, const double *nodes, const double *w, double xi, double *N);
void orc_shape_dN(int32_t m, const double *nodes, const double *w, double xi, double *dN);
int32_t orc_gauss(int32_t nq, double *pts, double *wts);
double orc_integrate(int32_t nq, const double *fvals); /* sum(f .* w), quadrature.jl:68-70 */

/* ---- pattern / partition (integer work, bit-exact contract) ---- */
int64_t orc_pattern(const orc_model *m, int64_t *rowptr /*ndof+1*/, int32_t *colind /*NULL = count only*/);
/* slab partition by contiguous node ranges: node_off[nranks+1]. For `rank`:
 * returns counts; arrays (may be NULL for a counting pass) sized by a previous call.
 *   loc_elems : global ids of elements touching an owned node, ascending
 *   ghost_nodes: global ids of non-owned nodes of those elements, ascending
 *   send lists: for every other rank q (ascending), owned nodes that are ghosts on q, ascending */
int32_t orc_partition(const orc_model *m, int32_t nranks, const int64_t *node_off, int32_t rank,
                      int64_t *n_loc_elems, int64_t *loc_elems,
                      int64_t *n_ghost, int64_t *ghost_nodes,
                      int64_t *send_ptr /*nranks+1*/, int64_t *send_nodes);

/* ---- operators (gallery/residual.jl:12-26, gallery/jacobian.jl:10-22) ---- */
void orc_init_state(orc_model *m);                       /* initialize_state! */
void orc_commit_state(orc_model *m, const double *U);    /* save_state! */
void orc_residual(orc_model *m, double *U, double *R);   /* U is mutated (Dirichlet overwrite) */
void orc_jacobian_dense(orc_model *m, double *U, double *K /*[ndof][ndof] row-major*/);
void orc_jacobian_csr(orc_model *m, double *U, const int64_t *rowptr, const int32_t *colind, double *vals);
void orc_assemble_abs(orc_model *m, double *U, double *Rabs, const int64_t *rowptr, const int32_t *colind, double *vabs); /* test metric: sum |element contributions| */
void orc_element(const orc_model *m, int64_t e, const double *ue, double *fe, double *Ke); /* one element, no state update */

/* ---- linear algebra ---- */
int32_t orc_dense_solve(int64_t n, double *A /*row-major, destroyed*/, double *b /*in: rhs, out: x*/);
/* consistent mass at node level, gallery/mass.jl:2-17 (dense nn x nn; small cases only) */
void orc_mass_nodes(const orc_model *mo, double rho, double area, double *Mn);
void orc_spmv(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals, const double *x, double *y, int32_t nthreads);
int32_t orc_pcg(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals,
                const uint8_t *constrained, const double *b, double *x, double rtol, int32_t maxits,
                double *relres, int32_t *its, int32_t nthreads); /* 0 ok, 1 maxits, 2 breakdown */

/* ---- Newton (nonlinearsolvers/newton_fem.jl:132-190) ---- */
void orc_newton_dense(orc_model *m, const orc_newton_opts *o, const double *U0, double *U, double *R,
                      double *norm_hist, orc_newton_result *res);
void orc_newton_pcg(orc_model *m, const orc_newton_opts *o, const int64_t *rowptr, const int32_t *colind,
                    double *vals, const double *U0, double *U, double *R, double *norm_hist,
                    orc_newton_result *res);

#ifdef __cplusplus
}
#endif
#endif
