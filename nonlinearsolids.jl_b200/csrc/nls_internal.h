// nls_internal.h -- shared declarations of libnls_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/nls_b200.h"

#define NLS_MAX_NEN1D 8   // 1-D Lagrange order p <= 7
#define NLS_MAX_NQ 5      // gauss_quadrature(1..5), quadrature.jl:29-50

// ---- shape/quadrature tables, passed to kernels by value (__grid_constant__) ----
struct Tab1D {            // dN_i/dxi at each Gauss point, shapefunctions.jl:64-81
  int m, nq;
  double w[NLS_MAX_NQ];
  double dN[NLS_MAX_NQ][NLS_MAX_NEN1D];
  double N[NLS_MAX_NQ][NLS_MAX_NEN1D];   // N_i at each Gauss point (shapefunctions.jl:50-61), for body forces
};
struct TabQ1 {            // tensor-product Q1: [qp][node][ref dir], x fastest
  int dim, nen, nqp;
  double w[8];
  double dN[8][8][3];
  double N[8][8];           // [qp][node]
};
struct MatParams {
  int kind;
  double p[8];
  double area;
};

// ---- host-side tabulation with the reference formulas (host_tables.cpp) ----
void nls_host_lagrange_nodes(int p, int kind, double *nodes);
void nls_host_lagrange_weights(int m, const double *nodes, double *w);
void nls_host_shape_N(int m, const double *x, const double *w, double xi, double *N);
void nls_host_shape_dN(int m, const double *x, const double *w, double xi, double *dN);
int nls_host_gauss(int nq, double *x, double *w);
int nls_host_tab1d(int p, int nq, Tab1D *t);
int nls_host_tabq1(int dim, TabQ1 *t);

// node-level sparsity pattern for the owned nodes (block rows); scalar CSR is its
// dim x dim expansion. bcol ascending within a row.
void nls_host_block_pattern(int nen, int64_t nel, const int32_t *conn, int64_t nn, int64_t n_owned,
                            std::vector<int64_t> &browptr, std::vector<int32_t> &bcol);

// cluster plan of the gather (row-owner) assembly, built on the host (host_tables.cpp)
struct AsmPlan {
  int nc = 0;                     // nodes per cluster
  int64_t ncl = 0;                // clusters
  int max_elems = 0;              // largest element count of a cluster
  int max_adj = 0;                // largest number of elements around a node
  std::vector<int32_t> nodes;     // [ncl*nc] owned node id or -1
  std::vector<int32_t> eptr;      // [ncl+1] into elems
  std::vector<int32_t> elems;     // element ids touching the cluster, ascending per cluster
};
void nls_host_cluster_plan(int dim, int nen, int64_t nel, const int32_t *conn, int64_t nn, int64_t n_owned,
                           const double *coords, int nc, AsmPlan &plan, const int *tile = nullptr);
void nls_gather_tile(int *tile3);

// packed, fixed-stride device copy of the plan (kernels_gather.cu)
struct GatherPlanDev {
  int64_t ncl = 0, ncl_interior = 0;      // clusters [0, ncl_interior) touch no ghost node
  int maxt = 0, maxe = 0, maxb = 0, nsw = 0;
  int32_t *tn = nullptr;
  uint32_t *lconn = nullptr;
  int32_t *cn = nullptr;
  int32_t *b0 = nullptr;
  int32_t *deg = nullptr;
  uint16_t *btask = nullptr;
  uint32_t *bsrc = nullptr;
  uint16_t *rsrc = nullptr;
};

// plan of the two-phase (element records -> row owners) assembly (kernels_twophase.cu)
struct TwoPhaseChunk { int64_t n0, n1, e_lo, e_hi; };
struct TwoPhasePlan {
  bool ok = false;                // plan built and the mesh fits the packed formats
  int64_t estride = 0, np = 0;
  int maxadj = 0, rs = 0, nent = 0;
  int32_t *esrc = nullptr;
  uint64_t *asrc = nullptr;
  double *scratch = nullptr;      // allocated on first use
  size_t scratch_bytes = 0;
  std::vector<TwoPhaseChunk> chunks;
};

// plan of the pair-owner assembly of 3-D meshes with grid topology (kernels_lines3d.cu)
struct Lines3Plan {
  bool ok = false;                 // the mesh has grid topology and the closed-form row addresses match the pattern (3-D: the line kernel runs)
  bool tables = false;             // the addressing tables exist (3-D as above; 2-D: for the multigrid preconditioner only)
  bool lap_valid = false;          // the Laplacian table holds the current mesh
  int nx = 0, ny = 0, nzp = 0, nunits = 0;
  int64_t lap_doubles = 0;
  int32_t *d_pbase = nullptr;
  uint8_t *d_pown = nullptr;
  void *d_info = nullptr;          // L3Line [nzp * ny]
  void *d_units = nullptr;         // int4 [nunits]
  int64_t *d_lapoff = nullptr;
  double *d_lap = nullptr;
  unsigned *d_counter = nullptr;
  double *d_carry = nullptr;       // [sm_count][18][384] per-CTA carry scratch of the line kernel
  std::vector<uint8_t> h_pown;     // host copies for the multigrid hierarchy (kernels_mg.cu)
  std::vector<int32_t> h_pbase;
};

// geometric multigrid preconditioner on grid-topology meshes (kernels_mg.cu)
#define MG_MAX_LEVELS 10
struct MgLevel {
  int nx = 0, ny = 0, NZ = 0;      // global grid of the level
  int nzb = 0, zbase = 0;          // planes stored here: global [zbase, zbase + nzb) (levels >= 2: the whole grid)
  int zown0 = 0, zown1 = 0;        // planes this rank works on (levels >= 2: all)
  int rown0 = 0, rown1 = 0;        // planes whose fine planes this rank owns (its share of a replicated level's sums)
  int64_t nn = 0;                  // nodes stored
  double *K = nullptr;             // [nn][27][9] Galerkin operator (levels >= 1; level 0 is the CSR Jacobian)
  double *dinv = nullptr;          // [nn][9] inverse node blocks (level 0: by node id)
  double *x = nullptr, *b = nullptr, *y = nullptr, *d = nullptr;
  double lmax = 1.0;               // largest eigenvalue of D^-1 K (power iteration, inflated)
};
struct MgPlan {
  bool ok = false;                 // the mesh can have a hierarchy
  bool laid_out = false;           // levels and buffers exist (first set-up; again when the communicator changes)
  bool dist = false;
  double *d_vote = nullptr;        // multi-GPU: scratch of the per-solve agreement on the preconditioner
  bool ghost_below = false, ghost_above = false;
  int nlev = 0, z0 = 0, nzo = 0;   // line-table index of the first owned plane, owned planes
  int p0 = 0, NZ0 = 0;             // global index of the first owned plane, global plane count
  int rank_below = -1, rank_above = -1;
  int64_t version = -1;            // vals_version the operators were built for
  int mode = 2;                    // 1 block-Jacobi alone, 2 V-cycle
  int degree = 2, coarse_degree = 24, power_its = 10, max_levels = MG_MAX_LEVELS;
  double ratio = 8.0, coarse_ratio = 60.0;
  int fp32 = 1;                    // the smoother's level-0 products read a single-precision copy of the CSR values
  float *vals32 = nullptr;
  int64_t power_version = -1000;   // vals_version of the last full power iteration
  double *gbelow = nullptr, *gabove = nullptr, *gsend = nullptr;   // level-0 boundary row planes in stencil format
  double *rf = nullptr;            // level-0 residual with ghost entries (restriction input)
  uint8_t *conf = nullptr;         // constrained mask of all local nodes, ghost entries from their owners
  MgLevel lev[MG_MAX_LEVELS];
  // single GPU: the V-cycle (~100 launches, most of them microseconds long on the coarse levels) replayed as one CUDA graph;
  // captured again whenever the operators, the eigenvalue bounds, the smoother options or the vectors it was captured on change
  int use_graph = 1;
  bool graph_off = false;          // capture failed once: plain launches from then on
  cudaGraphExec_t vgraph = nullptr;
  int64_t vgraph_version = -1, vgraph_launches = 0;
  const double *vgraph_r = nullptr;
  double *vgraph_z = nullptr;
  double vgraph_key[6] = {0, 0, 0, 0, 0, 0};
};

// ---- device control block of the PCG loop (lives in device memory) ----
struct PcgCtl {
  double rz, rz_new, pAp, rr, bb;   // reductions (global, after allreduce if distributed)
  double alpha, beta;
  double tol2;                      // lin_rtol^2
  int iters;
  int done;                         // 0 running, 1 converged, 2 maxits(not set on device), 3 breakdown
  int maxits;
  int pad;
  double red[4];                    // scratch for allreduce: [0]=pAp partial, [1]=rz partial, [2]=rr partial
};

// captured chunk of 16 PCG iterations (single GPU), valid while the kernels' fixed arguments are unchanged
struct PcgGraph {
  cudaGraphExec_t exec = nullptr;
  const void *vals = nullptr, *con = nullptr;
  int64_t n = 0, nnz = 0;
  int spmv = 0, lanes = 0;
};

struct NcclApi;                     // nccl_dyn.h

// peer-memory (CUDA IPC over NVLink) data path of the halo exchange and the Krylov reductions (kernels_p2p.cu)
struct P2PState {
  bool on = false;
  double *win = nullptr;                       // my window: [2][H] halo areas, then [2][nranks][4] reduction slots
  unsigned long long *flags = nullptr;         // my flags: [0][src] halo epochs, [1][src] reduction epochs
  int64_t H = 0;
  double *dst[2][8] = {};                      // per parity, per neighbour: where my planes go in its window
  unsigned long long *dst_flag[8] = {};
  const unsigned long long *my_flag[8] = {};
  double **d_peer_red[2] = {nullptr, nullptr}; // device arrays [nranks]
  unsigned long long **d_peer_rflag = nullptr;
  const double *my_red[2] = {nullptr, nullptr};
  unsigned long long halo_epoch = 0, red_epoch = 0;
  unsigned int *d_done = nullptr;
  int *d_err = nullptr;                        // device-side sticky error flag
  volatile int *h_err = nullptr;               // its twin in mapped pinned memory (host pointer)
  volatile int *d_herr = nullptr;              // ... and the device pointer to it
  std::vector<void *> opened;
};

struct nls_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;   // second stream: halo exchange overlapped with the interior clusters of the 2-D assembly
  cudaStream_t halo_stream = nullptr;   // when set, nls_halo_exchange enqueues there instead of `stream`
  cudaEvent_t ev_halo0 = nullptr, ev_halo1 = nullptr;
  int opt_overlap_halo = 1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;

  // mesh
  int dim = 0, nen = 0;
  int64_t nn = 0, nel = 0, n_owned = 0;
  std::vector<int32_t> h_conn;
  int32_t *d_conn = nullptr;
  double *d_coords = nullptr;
  bool have_mesh = false;

  // discretisation / material
  int p = 1, nq = 2;
  bool have_disc = false;
  Tab1D tab1{};
  TabQ1 tabq{};
  MatParams mat{};
  bool have_mat = false;

  // boundaries (deduplicated, last write wins like sequential apply!)
  int64_t ndir = 0, ndir_user = 0;
  std::vector<int64_t> dir_user_dofs;   // as given (may hold duplicates)
  std::vector<int64_t> dir_slot;        // user index -> unique slot or -1 if overridden
  int64_t *d_dir_dofs = nullptr;
  double *d_dir_vals = nullptr;
  uint8_t *d_con = nullptr;             // constrained mask, ndof_local
  double body[3] = {0.0, 0.0, 0.0};     // DistributedForce density per component at the current time
  bool have_body = false;
  int64_t nneu = 0;
  int64_t *d_neu_dofs = nullptr;
  double *d_neu_vals = nullptr;

  // pattern
  bool have_pattern = false;
  int64_t nrows = 0, nnz = 0, nnzb = 0;
  std::vector<int64_t> h_browptr;
  std::vector<int32_t> h_bcol;
  int64_t *d_browptr = nullptr;
  int32_t *d_bcol = nullptr;
  int64_t *d_rowptr = nullptr;
  int32_t *d_colind = nullptr;
  double *d_vals = nullptr;
  uint16_t *d_pos = nullptr;            // [nel][nen][nen] position of node b in block row of node a
  int32_t *d_diag = nullptr;            // [nrows] offset of the diagonal entry within its row

  // 3-D: deformation-independent (Laplacian) part of the tangent per element, built once per mesh (hex_element.cuh)
  double *d_lap = nullptr;
  int64_t lap_stride = 0;
  bool lap_valid = false;
  int opt_lap3d = 1;      // 1: use it when the 320 B/element fit in HBM

  // symmetric product of the Krylov solvers: packed upper blocks + index (kernels_krylov.cu)
  int32_t *d_dpos = nullptr;
  int64_t *d_symptr = nullptr;
  uint32_t *d_lowslot = nullptr;
  double *d_symvals = nullptr;
  int64_t nsym = 0;
  bool have_sym_index = false, sym_active = false;
  int64_t vals_version = 0, sym_version = -1;   // bumped whenever the CSR values change
  int opt_sym_spmv = 0;   // experimental (nls_set_option "sym_spmv"): measured slower than the full product, see DESIGN.md 4

  // gather-assembly plan (device copy)
  bool have_plan = false;
  GatherPlanDev gplan;
  TwoPhasePlan tplan;
  Lines3Plan lplan;
  MgPlan mg;
  // exact banded solve of small systems (kernels_direct.cu)
  double *d_band = nullptr;
  int *d_ipiv = nullptr;
  size_t direct_cap = 0;
  int direct_bw = -1;        // half bandwidth of the scalar matrix (-1: not computed yet)
  int opt_direct = 1;        // Newton's linear solve: exact banded LU when the system is small enough
  int opt_precond = -1;    // -1 auto (multigrid on large grid-topology 3-D meshes, else Jacobi), 0 Jacobi, 1 block-Jacobi, 2 multigrid
  std::vector<double> h_coords;

  // vectors (ndof_local unless noted)
  double *d_U = nullptr, *d_R = nullptr, *d_dU = nullptr;
  double *d_r = nullptr, *d_p = nullptr, *d_Ap = nullptr, *d_dinv = nullptr, *d_b = nullptr;
  double *d_tmp = nullptr;
  PcgGraph pcg_graph;
  int opt_pcg_graph = 1;
  int opt_pcg_small = 1;   // systems of <= 8192 rows: the whole PCG solve in one thread block
  double *d_mr[7] = {nullptr};          // MINRES work vectors (allocated on first use)
  void *d_mctl = nullptr, *h_mctl = nullptr;   // MINRES control block (device, pinned mirror)
  double *d_partial = nullptr;          // block partial sums
  PcgCtl *d_ctl = nullptr;
  PcgCtl *h_ctl = nullptr;              // pinned mirror
  double *d_scal = nullptr;             // small scalar scratch (8 doubles)
  double *h_scal = nullptr;             // pinned

  // dynamics: node-level consistent mass [nnzb], Newmark vectors
  double *d_mass = nullptr;
  bool have_mass = false;
  double mass_rho = 0.0, mass_area = 1.0;
  double *d_Ut = nullptr, *d_Udt = nullptr, *d_Udd = nullptr, *d_Ud = nullptr;   // predictors, acceleration, velocity
  bool dyn_on = false;                  // the Newton unknown is the acceleration (newmark.jl:75-107)
  double dyn_cu = 0.0, dyn_cv = 0.0;    // dt^2 beta, dt gamma
  uint8_t *d_con_zero = nullptr;        // all-free mask for the unconstrained mass solve

  // plastic state [nel][nq]
  double *d_state[9] = {nullptr};       // sig, alp, kap, gam, dsig, sig_n, alp_n, kap_n, gam_n
  double *d_j2 = nullptr;               // J2 material: [22][nel * nq] committed (7), trial (7), tangent data (8)  (kernels_j2.cu)
  int64_t j2_nqp = 0;
  double *d_u_n = nullptr;              // [nel][nen]

  // per-handle (= per-device) record of the dynamic shared-memory limits already raised with cudaFuncSetAttribute
  // (always raised to the device's opt-in maximum, so handles sharing a device cannot lower each other's limit)
  bool attr_asm3d = false, attr_tp3d = false, attr_gather = false, attr_lines3 = false, attr_rows[2][2][2] = {};   // rows: [dim-2][DO_R][DO_K]
  int max_optin_smem = 227 * 1024;

  // L2 flush scratch
  void *d_flush = nullptr;
  size_t flush_bytes = 0;

  // options
  int opt_spmv = 1;        // 0: scalar CSR SpMV; 1: block-aware traversal (dim >= 2)
  int opt_spmv_lanes = 0;  // lanes per node for the block-aware SpMV (0 = default)
  int opt_profile_phases = 0; // debug: per-phase cycle counters of the gather kernel
  uint64_t *d_prof = nullptr;
  int opt_gather_ctas = 0; // persistent CTAs per SM of the gather kernel (0 = default)
  int opt_assembly = 1;   // 0: atomic scatter everywhere; 1: best atomic-free scheme available (2-D: shared-memory
                          // gather, 3-D: line pairs on grid-topology meshes, else the atomic scatter); 2: element records + row owners
  int opt_twophase_3d = 0; // 1: the 3-D default becomes element records + row owners instead of the atomic scatter
  int opt_scratch_mb = 0; // byte budget of the element-record scratch in MiB (0 = default 8192)

  // distributed
  bool distributed = false;
  int rank = 0, nranks = 1;
  int nneigh = 0;
  std::vector<int> neigh;
  std::vector<int64_t> send_ptr, recv_ptr;   // in nodes
  int32_t *d_send_nodes = nullptr, *d_recv_nodes = nullptr;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
  NcclApi *nccl = nullptr;
  void *comm = nullptr;
  P2PState p2p;

  int64_t ndof_local() const { return nn * (int64_t)dim; }
};

// error helpers
#define NLS_FAIL(h, code, msg) do { (h)->err = (msg); return (code); } while (0)
#define NLS_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return NLS_ERR_CUDA; } } while (0)

// ---- kernel launchers (kernels_assembly.cu / kernels_krylov.cu) ----
int nls_launch_assembly(nls_context *h, bool want_r, bool want_k);
enum { NLS_ASM_ATOMIC = 0, NLS_ASM_GATHER = 1, NLS_ASM_TWOPHASE = 2, NLS_ASM_LINES = 3 };
int nls_assembly_mode(nls_context *h);          // which scheme the next assembly uses (allocates the record scratch if needed)
bool nls_assembly_overwrites(nls_context *h);   // true when the scheme writes every row itself (no zero-fill)
int nls_twophase_build(nls_context *h);
void nls_twophase_free(nls_context *h);
int nls_launch_twophase(nls_context *h, bool want_r, bool want_k);
int nls_lines3_build(nls_context *h);
int nls_lines2_build(nls_context *h);
void nls_lines3_free(nls_context *h);
int nls_launch_lines3(nls_context *h, bool want_r, bool want_k);
int nls_gather_nc(void);
int nls_gather_build(nls_context *h, const AsmPlan &plan);
void nls_gather_free(nls_context *h);
int nls_launch_gather_2d(nls_context *h, bool want_r, bool want_k, int64_t c_begin = 0, int64_t c_end = -1);
int nls_launch_apply_dirichlet(nls_context *h, double *U);
int nls_launch_neumann(nls_context *h, double *R);
int nls_launch_body_force(nls_context *h, double *R);
int nls_launch_build_pos(nls_context *h);
int nls_launch_expand_pattern(nls_context *h);
int nls_launch_commit_state(nls_context *h, const double *dU);

int nls_launch_spmv(nls_context *h, const double *x, double *y, bool masked, bool with_dot);
int nls_launch_extract_dinv(nls_context *h);
int nls_sym_build_index(nls_context *h);
int nls_sym_prepare(nls_context *h);
double nls_dev_masked_norm2(nls_context *h, const double *v, bool masked, int *rc);  // sqrt(sum v_i^2) over owned (free) DoFs
int nls_dev_dot(nls_context *h, int64_t n, const double *x, const double *y, double *out);
int nls_launch_axpy(nls_context *h, int64_t n, double a, const double *x, double *y);
int nls_pcg_solve(nls_context *h, double rtol, int maxits, int fixed_iters, int *its, double *relres, int *status);
int nls_minres_solve(nls_context *h, double rtol, int maxits, int *its, double *relres, int *status);

int nls_launch_mass_build(nls_context *h, double rho, double area);
int nls_launch_apply_mass(nls_context *h, const double *x, double *y, bool accumulate);
int nls_launch_scale_add_mass(nls_context *h, double scale, bool scale_existing);
int nls_launch_newmark_state(nls_context *h, double c_u, double c_v);

int nls_halo_exchange(nls_context *h, double *v);
int nls_p2p_setup(nls_context *h);
void nls_p2p_free(nls_context *h);
int nls_p2p_halo_exchange(nls_context *h, double *v);
int nls_p2p_allreduce(nls_context *h, double *dptr, int count);
int nls_p2p_check(nls_context *h);
int nls_p2p_failed(nls_context *h);   // after a stream synchronisation: NLS_ERR_NCCL if an exchange has timed out
int nls_allreduce_sum(nls_context *h, double *dptr, int count);
int nls_dev_dot_local(nls_context *h, int64_t n, const double *x, const double *y, double *out);   // this rank's part only
int nls_launch_vals_to_f32(nls_context *h, float *out);
int nls_launch_spmv_f32(nls_context *h, const float *vals32, const double *x, double *y);
struct AsmArgs;
void nls_j2_free(nls_context *h);
int nls_j2_init(nls_context *h);
int nls_j2_commit(nls_context *h);
int nls_j2_get(nls_context *h, const char *name, double *out);
int nls_launch_j2(nls_context *h, const AsmArgs &a, bool want_r, bool want_k);
bool nls_direct_available(nls_context *h);
int nls_direct_solve(nls_context *h, int *status);
void nls_direct_free(nls_context *h);
void nls_mg_free(nls_context *h);
void nls_mg_invalidate(nls_context *h);
int nls_mg_build(nls_context *h);
int nls_mg_agreed(nls_context *h, bool *use);   // collective when distributed
int nls_mg_prepare(nls_context *h, bool collective);
int nls_mg_setup(nls_context *h);
int nls_pcg_solve_mg(nls_context *h, double rtol, int maxits, int *its, double *relres, int *status);
