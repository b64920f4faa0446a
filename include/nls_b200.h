/*
 * nls_b200.h -- C ABI of libnls_b200.so: the B200-native (sm_100a) Newton hot path
 * of NonlinearSolids.jl. Flat `extern "C"`, plain pointers and sizes, no torch or
 * CUDA types. This is what a Julia `ccall` shim binds; in this
 * repository the same symbols are exercised from Python ctypes.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * the NonlinearSolids.jl source tree).
 *
 * Conventions
 *  - every function returns an nls_status (0 = ok); on failure nls_last_error(h)
 *    holds a message. CUDA/NCCL errors are sticky on the handle.
 *  - host pointers are borrowed for the duration of the call only (synchronous
 *    copy-in/copy-out); the library owns all device memory behind the handle.
 *  - indices are 0-based. DoF numbering is node-major interleaved:
 *    gdof = dim*node + component (1-D: identical to the reference's node ids - 1).
 *  - one handle = one GPU = one host thread at a time. Multi-GPU = one process per
 *    GPU, each with its own handle holding its slab of the mesh (nls_set_halo +
 *    nls_comm_init).
 *  - there is NO CPU fallback: nls_create fails with NLS_ERR_NOGPU without a device.
 */
#ifndef NLS_B200_H
#define NLS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nls_context *nls_handle;

typedef enum {
  NLS_OK = 0,
  NLS_ERR_ARG = 1,         /* ArgumentError on the Julia side */
  NLS_ERR_STATE = 2,       /* call order violated (e.g. assemble before nls_build_pattern) */
  NLS_ERR_CUDA = 3,
  NLS_ERR_NCCL = 4,
  NLS_ERR_NOGPU = 5,
  NLS_ERR_UNSUPPORTED = 6
} nls_status;

/* src/materials/{linearelasticity,exponential,linearelastoplasticity}.jl + the T2 extension */
typedef enum {
  NLS_MAT_LINEAR_ELASTIC_1D = 0,   /* params: E                       (linearelasticity.jl:11-19) */
  NLS_MAT_EXPONENTIAL_1D = 1,      /* params: sigma_sat, b            (exponential.jl:11-21) */
  NLS_MAT_ELASTOPLASTIC_1D = 2,    /* params: E, Hk, Ha, k0, a0, Eep  (linearelastoplasticity.jl:5-20) */
  NLS_MAT_NEOHOOKEAN = 3,          /* params: mu, lambda; dim 2 (plane strain) or 3 */
  NLS_MAT_J2_SMALL_STRAIN = 4      /* params: K (bulk), G (shear), sigma_y, H (isotropic hardening); dim 2 (plane strain) or 3.
                                      The stateful-material protocol of linearelastoplasticity.jl:60-117 on Q1 elements
                                      (SURVEY 8 f-2): per-qp committed / trial plastic strain and alpha, radial return in the
                                      internal-force pass, stored consistent-tangent data read by the tangent pass */
} nls_material_kind;

/* src/nonlinearsolvers/types.jl:22-32 REASON_* singletons, 1:1 */
typedef enum {
  NLS_REASON_NOT_YET_CONVERGED = 0,
  NLS_REASON_CONVERGED_RELATIVE = 1,
  NLS_REASON_CONVERGED_ABSOLUTE = 2,
  NLS_REASON_DIVERGED_MAXITS = 3,
  NLS_REASON_DIVERGED_STEP = 4,
  NLS_REASON_DIVERGED_LINEAR_SOLVE = 5
} nls_reason;

/* NewtonSolver kwdef fields (newton_fem.jl:5-47) + the Krylov controls that replace `\` */
typedef struct {
  int32_t type_modified;   /* 0 = :standard, 1 = :modified */
  double atol, rtol, dtol;
  int32_t maxits;
  double lin_rtol;         /* PCG: ||r|| <= lin_rtol * ||b|| */
  int32_t lin_maxits;
} nls_newton_opts;

typedef struct {
  int32_t reason;          /* nls_reason */
  int32_t iterations;      /* solver.iterations */
  int32_t lin_its_total;   /* PCG iterations summed over Newton iterations */
  int32_t nhist;           /* entries written to norm_hist */
  double norm_r, norm_r0;
  double ms_assembly, ms_linear, ms_total; /* device time (CUDA events) */
} nls_newton_result;

/* ---- lifecycle ----------------------------------------------------------- */
int32_t nls_create(int32_t device, nls_handle *h);
int32_t nls_destroy(nls_handle h);
const char *nls_last_error(nls_handle h);
const char *nls_version(void);

/* ---- host-only building blocks (no GPU needed) ------------------------------ */
/* lagrange(order, :chebyshev2 | :equidistant) (src/fem/shapefunctions.jl:23-47): nodes and
 * barycentric weights; kind 0 = :chebyshev2, 1 = :equidistant. Arrays hold p+1 entries. */
int32_t nls_lagrange(int32_t p_order, int32_t kind, double *nodes, double *weights);
/* N(shape, xi) and gradN(shape, xi) (shapefunctions.jl:50-81); either output may be NULL */
int32_t nls_shape_eval(int32_t m, const double *nodes, const double *weights, double xi, double *N, double *dN);
/* gauss_quadrature(order) (src/fem/quadrature.jl:29-50), order 1..5 */
int32_t nls_gauss_quadrature(int32_t nq, double *points, double *weights);

/* The CSR pattern builder on its own (what nls_build_pattern computes on the host): rows of the
 * first n_owned nodes, columns ascending. Two-pass: colind == NULL returns only rowptr / nnz. */
int32_t nls_pattern_host(int32_t dim, int32_t nen, int64_t nelem, const int32_t *conn0, int64_t nnodes,
                         int64_t n_owned_nodes, int64_t *rowptr, int32_t *colind, int64_t *nnz);

/* Statistics of the cluster plan the gather (row-owner) assembly would use for this mesh:
 * out4 = {clusters, max elements per cluster, sum of elements over clusters, max elements around a node}.
 * sum/nelem is the redundancy factor of the element computation. */
int32_t nls_plan_stats(int32_t dim, int32_t nen, int64_t nelem, const int32_t *conn0, int64_t nnodes,
                       int64_t n_owned_nodes, const double *coords, int32_t nc, int64_t *out4);

/* ---- model description ---------------------------------------------------- */
/* Mesh / Element (src/fem/element.jl:6-12,73-83): flattened connectivity + coordinates.
 * nen = p+1 (1-D), 4 (2-D Q1, tensor order x-fastest), 8 (3-D Q1). */
int32_t nls_set_mesh(nls_handle h, int32_t dim, int64_t nnodes, int64_t nelem, int32_t nen,
                     const int32_t *conn0, const double *coords);
/* FEMModel(mesh, q, p) (src/fem/fem.jl:18-23): Gauss points per direction, Lagrange order
 * (:chebyshev2 nodes). Shape tables are tabulated on the host with the reference's own
 * formulas (shapefunctions.jl:23-81, quadrature.jl:29-50). */
int32_t nls_set_discretization(nls_handle h, int32_t p_order, int32_t nq);
/* Setting a stateful material (kinds 2 and 4) also (re)initialises its per-qp state: initialize_state!. */
int32_t nls_set_material(nls_handle h, int32_t kind, const double *params, int32_t nparams);
/* fem.data[:A] (materials/defaults.jl:3,16; default 3e-4) */
int32_t nls_set_area(nls_handle h, double area);
/* addboundary!(fem, Dirichlet(nodes, values)) (fem.jl:59-64, boundary.jl:21-29); replaces the list */
int32_t nls_set_dirichlet(nls_handle h, int64_t n, const int64_t *dofs0, const double *vals);
/* TimeDependent update! (boundary.jl:86-92): new values for the same DoF list */
int32_t nls_set_dirichlet_values(nls_handle h, int64_t n, const double *vals);
/* addboundary!(fem, Neumann(nodes, values)) (boundary.jl:32-39). The list is the concatenation of all Neumann boundaries
 * of the model; a DoF listed more than once receives the SUM of its values (fem.jl:84-90 adds boundary after boundary; inside
 * ONE reference boundary `F[nodes] .+= values` would keep only the last duplicate -- do not list a DoF twice in one boundary). */
int32_t nls_set_neumann(nls_handle h, int64_t n, const int64_t *dofs0, const double *vals);
int32_t nls_set_neumann_values(nls_handle h, int64_t n, const double *vals);
/* DistributedForce(fdist) (src/gallery/distributedforce.jl:3-27): body-force density per component at the
 * current time (the host evaluates fdist_fn(time(ctx))). Adds F_ext,a = sum_q (J f N_a) w_q, i.e. R -= F_ext
 * (residual.jl:19-25). ncomp = dim, or 0 to remove it. */
int32_t nls_set_body_force(nls_handle h, int32_t ncomp, const double *f);
/* initialize_state!(material, fem) (linearelastoplasticity.jl:106-117) */
int32_t nls_init_state(nls_handle h);
/* save_state!(material, fem, U, ctx), called from poststep! (timesteppers/types.jl:74-78) */
int32_t nls_commit_state(nls_handle h, const double *U);
/* read back one per-qp state array by name: 1-D elastoplasticity "sig","alp","kap","gam","dsig" (+"_n");
 * J2: "epsp_xx","epsp_yy","epsp_zz","epsp_xy","epsp_yz","epsp_xz","alpha" (trial; +"_n" = committed) */
int32_t nls_get_state(nls_handle h, const char *name, double *out /*[nelem][nq]*/);

/* ---- CSR pattern (host-built; replaces the dense K of newton_fem.jl:41,52) ---- */
/* pattern = {(i,j): some element holds both i and j} = what assemble! can touch (element.jl:70) */
int32_t nls_build_pattern(nls_handle h);
int32_t nls_pattern_size(nls_handle h, int64_t *nrows, int64_t *nnz);
int32_t nls_get_pattern(nls_handle h, int64_t *rowptr, int32_t *colind);

/* ---- operators, host buffers (drop-in for the Residual / Jacobian functors) ---- */
/* (::Residual)(fem, U, R, ctx) (gallery/residual.jl:12-26). U is in/out: Dirichlet entries
 * are overwritten, as the reference does. R has nrows entries. */
int32_t nls_residual(nls_handle h, double *U, double *R);
/* (::Jacobian)(fem, U, K, ctx) (gallery/jacobian.jl:10-22); vals in CSR order */
int32_t nls_jacobian(nls_handle h, double *U, double *vals);
/* both in one element pass (same results as residual-then-jacobian, newton_fem.jl:137-138) */
int32_t nls_residual_jacobian(nls_handle h, double *U, double *R, double *vals);
/* The material-level dispatch point alone -- compute_internalforce / compute_dinternalforce(material, fem, U, ...;
   mode = :set), src/materials/defaults.jl:29-46, 68-82: F_int(U) and dF_int/dU with U left untouched (no Dirichlet
   overwrite) and no external force subtracted. What a Residual / Jacobian functor of the reference calls underneath. */
int32_t nls_internal_force(nls_handle h, const double *U, double *Fint);
int32_t nls_internal_force_jacobian(nls_handle h, const double *U, double *vals);

/* ---- Krylov pieces, host buffers (test hooks; replace dofs(K) \ dofs(-R), newton_fem.jl:158) */
int32_t nls_set_values(nls_handle h, const double *vals);           /* load CSR values */
int32_t nls_spmv(nls_handle h, const double *x, double *y);          /* y = K x (all rows, unmasked) */
int32_t nls_dot(nls_handle h, int64_t n, const double *x, const double *y, double *out);
int32_t nls_axpy(nls_handle h, int64_t n, double a, const double *x, double *y); /* y += a x */
/* Jacobi-PCG on the unconstrained block of the handle's current values; x is zero on
 * constrained DoFs (expand(), fem.jl:98-102). status_out: 0 ok, 1 maxits, 2 breakdown */
int32_t nls_linear_solve(nls_handle h, double rtol, int32_t maxits, const double *b, double *x,
                         int32_t *its, double *relres, int32_t *status_out);

/* Same contract for a symmetric INDEFINITE matrix (the tangent past a limit point, arc-length continuation,
 * nonlinearsolvers/newtonarclength_fem.jl:104): Jacobi-preconditioned MINRES. status_out: 0 ok, 1 maxits, 2 breakdown */
int32_t nls_linear_solve_indefinite(nls_handle h, double rtol, int32_t maxits, const double *b, double *x,
                                    int32_t *its, double *relres, int32_t *status_out);

/* ---- Newton (solve!(::NewtonSolver, U0), newton_fem.jl:132-190) ------------- */
int32_t nls_newton_solve(nls_handle h, const nls_newton_opts *opts, const double *U0, double *U, double *R,
                         double *norm_hist /*maxits+1 or NULL*/, nls_newton_result *res);

/* ---- dynamics: consistent mass, dynamic operators, Newmark step -- */
/* Builds the consistent mass once (assemblemass!, src/gallery/mass.jl:2-17): rho = density(material),
 * area = get(fem.data, :A, 1.0) of mass.jl:9,27 (1-D only; ignored for Q1). Needs the pattern. */
int32_t nls_set_mass(nls_handle h, double rho, double area);
/* assemblemass!(material, fem, K) into zeroed CSR values (K = M) */
int32_t nls_get_mass(nls_handle h, double *vals);
/* applymass!(material, fem, a, R) (mass.jl:26-37): y += M x; y in/out, nrows entries */
int32_t nls_apply_mass(nls_handle h, const double *x, double *y);
/* Residual dynamic method (gallery/residual.jl:48-54): static residual of U (Dirichlet overwrite, in/out) + M Udd */
int32_t nls_dynamic_residual(nls_handle h, double *U, const double *Udd, double *R);
/* Jacobian dynamic method (gallery/jacobian.jl:42-51): K(U) * scale + M, scale = dUdd/dU(ctx) = dt^2 beta */
int32_t nls_dynamic_jacobian(nls_handle h, double *U, double scale, double *vals);
/* M x = b on all DoFs (initial acceleration, timesteppers/newmark.jl:136-140; Jacobi-PCG in place of `\`) */
int32_t nls_mass_solve(nls_handle h, double rtol, int32_t maxits, const double *b, double *x, int32_t *its, int32_t *status_out);
/* One step of solve!(ts::NewmarkTimeStepper) (newmark.jl:142-155): Newton on the acceleration with the closures
 * of newmark.jl:75-107 (update_state! -> static operators -> inertia terms), device-resident, from the
 * predictors Ut, Udt (update_predictors!, :109-112) and the initial guess Udd0 (previous acceleration).
 * Returns the converged acceleration and U, Ud = update_state!(Udd) (:114-117). */
int32_t nls_newmark_solve(nls_handle h, const nls_newton_opts *opts, double beta, double gamma, double dt,
                          const double *Ut, const double *Udt, const double *Udd0,
                          double *Udd, double *U, double *Ud, double *R, double *norm_hist, nls_newton_result *res);

/* ---- device-resident variants (no host<->device copies; U/R/vals stay in HBM) ---- */
int32_t nls_dev_upload_u(nls_handle h, const double *U);
int32_t nls_dev_download_u(nls_handle h, double *U);
int32_t nls_dev_download_r(nls_handle h, double *R);
int32_t nls_dev_download_vals(nls_handle h, double *vals);
/* Rows [row0, row1) of the device-resident operators: R[row0:row1], and of the CSR Jacobian the row pointers
   (row1 - row0 + 1 entries, as stored: offsets into the whole matrix), column indices and values of those rows.
   Any output may be NULL. For spot checks of large models whose whole Jacobian should not cross PCIe. */
int32_t nls_dev_download_rows(nls_handle h, int64_t row0, int64_t row1, double *R, int64_t *rowptr, int32_t *colind, double *vals);
int32_t nls_dev_assemble(nls_handle h, int32_t want_residual, int32_t want_jacobian);
int32_t nls_dev_spmv(nls_handle h);          /* work vector Ap = K * work vector p (p := U for the hook) */
int32_t nls_dev_pcg_iterations(nls_handle h, int32_t niter); /* exactly niter PCG iterations on b = -R */
int32_t nls_dev_newton_step(nls_handle h, double lin_rtol, int32_t lin_maxits, int32_t *lin_its, double *norm_r);
int32_t nls_dev_newton_solve(nls_handle h, const nls_newton_opts *opts, double *norm_hist, nls_newton_result *res);
int32_t nls_sync(nls_handle h);
int32_t nls_timer_start(nls_handle h);       /* CUDA event on the library's stream */
int32_t nls_timer_stop(nls_handle h, float *ms);
int32_t nls_flush_l2(nls_handle h);          /* overwrite a >L2-sized scratch buffer */
int64_t nls_kernel_launches(nls_handle h);   /* kernels launched by this handle so far */
/* which assembly scheme the next nls_*assemble / residual / jacobian call uses for this model:
 * 0 atomic scatter, 1 shared-memory gather (2-D), 2 element records + row owners; -1 without a pattern */
int32_t nls_assembly_scheme(nls_handle h);
/* debug: cycle counters of the gather kernel phases after nls_set_option(h, "profile_phases", 1):
 * {staging, element phase, 0, block-task phase, clusters, cp.async wait, barrier, issue of the next stage}, summed over CTAs */
int32_t nls_debug_counters(nls_handle h, int64_t *out8);
/* preconditioner of the next Krylov solve: 0 Jacobi, 1 block-Jacobi, 2 geometric multigrid (nls_set_option "precond").
 * COLLECTIVE after nls_comm_init: the ranks agree (a rank whose slab has no grid hierarchy vetoes the multigrid for all). */
int32_t nls_krylov_preconditioner(nls_handle h);
/* optional, collective after nls_comm_init: allocate the solver's work space (the multigrid hierarchy of the Krylov
 * preconditioner on grid-topology 3-D meshes) now instead of inside the first solve */
int32_t nls_solver_prepare(nls_handle h);
int32_t nls_debug_counters_ext(nls_handle h, int64_t *out24); /* + per-warp phase cycles of the 3-D line kernel */
/* selects a kernel variant for A/B measurement: key "assembly" -> 0 atomic scatter, 1 the default row-owner scheme
 * (2-D: shared-memory gather; 3-D grid-topology meshes: row-owner line kernel with TMA row stores, other 3-D meshes: atomic
 * scatter unless "twophase_3d" = 1), 2 element records + row owners;
 * "scratch_mb" -> byte budget of the element-record scratch; "gather_ctas" -> persistent CTAs per SM;
 * "spmv" -> 0 scalar CSR, 1 block-aware; "spmv_lanes" -> lanes per node (0 = default);
 * "precond" -> preconditioner of the Krylov solve that replaces `\`: -1 automatic (multigrid on grid-topology 3-D meshes of
 * >= 10 000 nodes, else Jacobi), 0 Jacobi, 1 block-Jacobi, 2 multigrid; "mg_degree" / "mg_ratio_x10" / "mg_coarse_degree" /
 * "mg_coarse_ratio" -> Chebyshev smoother (degree, eigenvalue interval), "mg_power_its" -> power iterations of the eigenvalue
 * bound, "mg_fp32" -> 1: the smoother's level-0 products read a single-precision copy of the CSR values (default),
 * "mg_max_levels" -> cap on the number of grid levels (default: coarsen until the thinnest axis has <= 6 nodes),
 * "mg_graph" -> 1: on one GPU the V-cycle is replayed as a CUDA graph (default), 0: plain launches */
int32_t nls_set_option(nls_handle h, const char *key, int32_t value);

/* ---- multi-GPU: one process per GPU, element-partitioned slabs ---------------- */
/* Host-only integer work (no GPU needed). Contiguous owned node ranges node_off[nranks+1];
 * a rank computes every element touching an owned node; ghosts = the remaining nodes of
 * those elements. Two-pass: call with NULL arrays to get the counts. */
int32_t nls_partition(int32_t nen, int64_t nelem, const int32_t *conn0, int32_t nranks,
                      const int64_t *node_off, int32_t rank,
                      int64_t *n_loc_elems, int64_t *loc_elems, int64_t *n_ghost, int64_t *ghost_nodes,
                      int64_t *send_ptr /*nranks+1*/, int64_t *send_nodes);
/* Local numbering: owned nodes [0, n_owned) then ghosts. Rows/R exist for owned nodes only.
 * Lists are local node ids grouped by neighbour rank (ptr arrays have nneigh+1 entries); the k-th node a rank
 * sends to a neighbour is the k-th node that neighbour receives from it. After nls_comm_init this call is
 * COLLECTIVE over the communicator (every rank sets its halo): it exchanges CUDA IPC handles of the peer-memory
 * windows (see nls_comm_path). */
int32_t nls_set_halo(nls_handle h, int64_t n_owned_nodes, int32_t nneigh, const int32_t *neigh_ranks,
                     const int64_t *send_ptr, const int32_t *send_nodes,
                     const int64_t *recv_ptr, const int32_t *recv_nodes);
/* NCCL bootstrap: rank 0 creates the 128-byte id, the host language broadcasts it. */
int32_t nls_comm_unique_id(const char *nccl_lib_path, uint8_t *id128);
int32_t nls_comm_init(nls_handle h, const char *nccl_lib_path, int32_t rank, int32_t nranks, const uint8_t *id128);
int32_t nls_dev_halo_exchange_u(nls_handle h); /* test hook: fill ghost entries of U */
/* Data path of the halo exchange / reductions: 0 single GPU, 1 NCCL (ncclSend/Recv, ncclAllReduce), 2 peer memory
 * (CUDA IPC windows over NVLink, kernels_p2p.cu; chosen at nls_set_halo when every rank can open every window;
 * NLS_B200_P2P=0 in the environment keeps NCCL) */
int32_t nls_comm_path(nls_handle h);
/* device-side barrier: one reduction over all ranks enqueued on the library stream (no host sync); lines the ranks'
 * streams up before a timed region */
int32_t nls_dev_barrier(nls_handle h);
/* average device time (us) of one halo exchange of U and of one 2-double allreduce over `iters` back-to-back calls */
int32_t nls_dev_comm_probe(nls_handle h, int32_t iters, float *us_halo, float *us_allreduce);

#ifdef __cplusplus
}
#endif
#endif
