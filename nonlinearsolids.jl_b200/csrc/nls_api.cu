// nls_api.cu -- the C ABI of libnls_b200.so (see include/nls_b200.h) and the host-side
// control flow around the kernels: model setup, CSR pattern, operator calls and the
// Newton loop of solve!(::NewtonSolver) (src/nonlinearsolvers/newton_fem.jl:132-190).
// There is no CPU fallback anywhere in this file: without a CUDA device nls_create fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include "nls_internal.h"
#include "nccl_dyn.h"

static std::string g_noctx_error = "no error";
static NcclApi g_nccl;

#define RED_BLOCKS_MAX 2048

#define NLS_ENTER(h) do { if (!(h)) { g_noctx_error = "null handle"; return NLS_ERR_ARG; } \
  cudaError_t e0_ = cudaSetDevice((h)->device); if (e0_ != cudaSuccess) { \
  (h)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e0_); return NLS_ERR_CUDA; } } while (0)
#define NLS_TRY(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

template <typename T>
static int dev_alloc(nls_context *h, T **p, size_t count) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (count == 0) count = 1;
  NLS_CUDA(h, cudaMalloc((void **)p, count * sizeof(T)));
  return NLS_OK;
}
template <typename T>
static void dev_free(T **p) { if (*p) { cudaFree(*p); *p = nullptr; } }

__global__ void k_set_mask(int64_t n, const int64_t *dofs, uint8_t *mask) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n) mask[dofs[k]] = 1;
}
__global__ void k_fill(int64_t n, double *p, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_negate_into(int64_t n, const double *x, double *y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = -x[i];
}

static void free_pattern(nls_context *h) {
  nls_direct_free(h);
  nls_j2_free(h);
  dev_free(&h->d_browptr); dev_free(&h->d_bcol); dev_free(&h->d_rowptr); dev_free(&h->d_colind);
  dev_free(&h->d_vals); dev_free(&h->d_pos); dev_free(&h->d_diag);
  nls_gather_free(h);
  nls_twophase_free(h);
  nls_lines3_free(h);
  dev_free(&h->d_mass); h->have_mass = false;
  dev_free(&h->d_dpos); dev_free(&h->d_symptr); dev_free(&h->d_lowslot); dev_free(&h->d_symvals);
  h->have_sym_index = false; h->sym_active = false; h->nsym = 0;
  if (h->pcg_graph.exec) { cudaGraphExecDestroy(h->pcg_graph.exec); h->pcg_graph = PcgGraph(); }
  h->h_browptr.clear(); h->h_bcol.clear();
  h->have_pattern = false; h->nrows = h->nnz = h->nnzb = 0;
}

static int alloc_state(nls_context *h) {
  if (h->mat.kind == NLS_MAT_J2_SMALL_STRAIN && h->have_mesh) return nls_j2_init(h);      // initialize_state! on every (re)set of the material
  if (h->mat.kind != NLS_MAT_ELASTOPLASTIC_1D || !h->have_mesh || !h->have_disc) return NLS_OK;
  if (h->d_state[0]) return NLS_OK;
  for (int i = 0; i < 9; ++i) NLS_TRY(dev_alloc(h, &h->d_state[i], (size_t)(h->nel * h->nq)));
  NLS_TRY(dev_alloc(h, &h->d_u_n, (size_t)(h->nel * h->nen)));
  return nls_init_state(h);
}

// ---- lifecycle ----------------------------------------------------------------------
extern "C" const char *nls_version(void) { return "nls_b200 0.1 (sm_100a)"; }

extern "C" const char *nls_last_error(nls_handle h) { return h ? h->err.c_str() : g_noctx_error.c_str(); }

extern "C" int32_t nls_create(int32_t device, nls_handle *out) {
  if (!out) { g_noctx_error = "nls_create: null output pointer"; return NLS_ERR_ARG; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_noctx_error = std::string("nls_create: no CUDA device (") + cudaGetErrorString(e) +
                    "); libnls_b200 has no CPU fallback";
    return NLS_ERR_NOGPU;
  }
  if (device < 0 || device >= ndev) { g_noctx_error = "nls_create: device index out of range"; return NLS_ERR_ARG; }
  nls_context *h = new nls_context();
  h->device = device;
  h->err = "no error";
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_noctx_error = std::string(#call) + ": " + \
  cudaGetErrorString(e_); delete h; return NLS_ERR_CUDA; } } while (0)
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_halo0, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_halo1, cudaEventDisableTiming));
  CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1));
  CK(cudaEventCreate(&h->ev2)); CK(cudaEventCreate(&h->ev3));
  CK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
  CK(cudaDeviceGetAttribute(&h->max_optin_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  CK(cudaMalloc((void **)&h->d_ctl, sizeof(PcgCtl)));
  CK(cudaMemset(h->d_ctl, 0, sizeof(PcgCtl)));
  CK(cudaMallocHost((void **)&h->h_ctl, sizeof(PcgCtl)));
  CK(cudaMalloc((void **)&h->d_scal, 8 * sizeof(double)));
  CK(cudaMallocHost((void **)&h->h_scal, 8 * sizeof(double)));
  CK(cudaMalloc((void **)&h->d_partial, (3 * RED_BLOCKS_MAX + 2) * sizeof(double)));
  CK(cudaMemset(h->d_partial, 0, (3 * RED_BLOCKS_MAX + 2) * sizeof(double)));
#undef CK
  h->mat.area = 3e-4;  // get(fem.data, :A, 3e-4), defaults.jl:3
  if (const char *ov = std::getenv("NLS_B200_OVERLAP")) h->opt_overlap_halo = (ov[0] != '0');
  *out = h;
  return NLS_OK;
}

extern "C" int32_t nls_destroy(nls_handle h) {
  if (!h) return NLS_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  nls_p2p_free(h);
  if (h->comm && h->nccl) h->nccl->CommDestroy(h->comm);
  free_pattern(h);
  dev_free(&h->d_conn); dev_free(&h->d_coords);
  dev_free(&h->d_dir_dofs); dev_free(&h->d_dir_vals); dev_free(&h->d_con);
  dev_free(&h->d_neu_dofs); dev_free(&h->d_neu_vals);
  dev_free(&h->d_U); dev_free(&h->d_R); dev_free(&h->d_dU); dev_free(&h->d_r); dev_free(&h->d_p);
  dev_free(&h->d_Ap); dev_free(&h->d_dinv); dev_free(&h->d_b); dev_free(&h->d_tmp);
  dev_free(&h->d_partial); dev_free(&h->d_ctl); dev_free(&h->d_scal); dev_free(&h->d_prof);
  for (int i = 0; i < 7; ++i) dev_free(&h->d_mr[i]);
  if (h->d_mctl) { cudaFree(h->d_mctl); h->d_mctl = nullptr; }
  if (h->h_mctl) { cudaFreeHost(h->h_mctl); h->h_mctl = nullptr; }
  for (int i = 0; i < 9; ++i) dev_free(&h->d_state[i]);
  dev_free(&h->d_u_n);
  dev_free(&h->d_send_nodes); dev_free(&h->d_recv_nodes); dev_free(&h->d_sendbuf); dev_free(&h->d_recvbuf);
  dev_free(&h->d_Ut); dev_free(&h->d_Udt); dev_free(&h->d_Udd); dev_free(&h->d_Ud); dev_free(&h->d_con_zero);
  dev_free(&h->d_lap);
  if (h->d_flush) cudaFree(h->d_flush);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->h_scal) cudaFreeHost(h->h_scal);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3);
  cudaEventDestroy(h->ev_halo0); cudaEventDestroy(h->ev_halo1);
  cudaStreamDestroy(h->comm_stream);
  cudaStreamDestroy(h->stream);
  delete h;
  return NLS_OK;
}

// ---- model description ------------------------------------------------------------------
extern "C" int32_t nls_set_mesh(nls_handle h, int32_t dim, int64_t nnodes, int64_t nelem, int32_t nen,
                                const int32_t *conn0, const double *coords) {
  NLS_ENTER(h);
  if (dim < 1 || dim > 3) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_mesh: dim must be 1, 2 or 3");
  if (nnodes < 1 || nelem < 0 || !coords || (nelem > 0 && !conn0)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_mesh: bad sizes or null arrays");
  if (nnodes * (int64_t)dim > 2147483647LL) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_mesh: local DoF count exceeds int32 column indices");
  if ((dim == 1 && (nen < 2 || nen > NLS_MAX_NEN1D)) || (dim == 2 && nen != 4) || (dim == 3 && nen != 8))
    NLS_FAIL(h, NLS_ERR_ARG, "nls_set_mesh: nen must be p+1 (2..8) in 1-D, 4 in 2-D, 8 in 3-D");
  for (int64_t k = 0; k < nelem * nen; ++k)
    if (conn0[k] < 0 || conn0[k] >= nnodes) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_mesh: connectivity entry out of range (expects 0-based node ids)");
  free_pattern(h);
  dev_free(&h->d_lap); h->lap_valid = false;
  for (int i = 0; i < 7; ++i) dev_free(&h->d_mr[i]);
  h->dim = dim; h->nen = nen; h->nn = nnodes; h->nel = nelem; h->n_owned = nnodes;
  h->h_conn.assign(conn0, conn0 + nelem * nen);
  h->h_coords.assign(coords, coords + nnodes * dim);
  NLS_TRY(dev_alloc(h, &h->d_conn, (size_t)(nelem * nen)));
  NLS_TRY(dev_alloc(h, &h->d_coords, (size_t)(nnodes * dim)));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_conn, conn0, sizeof(int32_t) * nelem * nen, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_coords, coords, sizeof(double) * nnodes * dim, cudaMemcpyHostToDevice, h->stream));
  const size_t nd = (size_t)h->ndof_local();
  NLS_TRY(dev_alloc(h, &h->d_U, nd)); NLS_TRY(dev_alloc(h, &h->d_R, nd)); NLS_TRY(dev_alloc(h, &h->d_dU, nd));
  NLS_TRY(dev_alloc(h, &h->d_r, nd)); NLS_TRY(dev_alloc(h, &h->d_p, nd)); NLS_TRY(dev_alloc(h, &h->d_Ap, nd));
  NLS_TRY(dev_alloc(h, &h->d_dinv, nd)); NLS_TRY(dev_alloc(h, &h->d_b, nd)); NLS_TRY(dev_alloc(h, &h->d_tmp, nd));
  NLS_TRY(dev_alloc(h, &h->d_con, nd));
  NLS_CUDA(h, cudaMemsetAsync(h->d_con, 0, nd, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_U, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_dU, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_p, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->ndir = h->ndir_user = 0; h->nneu = 0; h->have_body = false;
  dev_free(&h->d_Ut); dev_free(&h->d_Udt); dev_free(&h->d_Udd); dev_free(&h->d_Ud); dev_free(&h->d_con_zero);
  h->dir_user_dofs.clear(); h->dir_slot.clear();
  for (int i = 0; i < 9; ++i) dev_free(&h->d_state[i]);
  dev_free(&h->d_u_n);
  h->have_mesh = true;
  h->have_disc = false;
  h->nneigh = 0;
  if (dim > 1) { h->p = 1; h->nq = 2; nls_host_tabq1(dim, &h->tabq); h->have_disc = true; }
  return NLS_OK;
}

extern "C" int32_t nls_set_discretization(nls_handle h, int32_t p_order, int32_t nq) {
  NLS_ENTER(h);
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_discretization: call nls_set_mesh first");
  if (h->dim == 1) {
    if (p_order + 1 != h->nen) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_discretization: element has nen != p+1 nodes");
    if (nq < 1 || nq > NLS_MAX_NQ) NLS_FAIL(h, NLS_ERR_ARG, "Gauss quadrature not implemented for this n (1..5 supported)");
    if (nls_host_tab1d(p_order, nq, &h->tab1)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_discretization: unsupported p/nq");
  } else {
    if (p_order != 1 || nq != 2) NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "multi-D elements are Q1 with 2 Gauss points per direction");
    nls_host_tabq1(h->dim, &h->tabq);
  }
  if (h->have_disc && (h->p != p_order || h->nq != nq)) { for (int i = 0; i < 9; ++i) dev_free(&h->d_state[i]); dev_free(&h->d_u_n); }
  h->p = p_order; h->nq = nq; h->have_disc = true;
  return alloc_state(h);
}

extern "C" int32_t nls_set_material(nls_handle h, int32_t kind, const double *params, int32_t nparams) {
  NLS_ENTER(h);
  if (!params) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_material: null params");
  const int need[5] = {1, 2, 5, 2, 4};
  if (kind < 0 || kind > 4) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_material: unknown material kind");
  if (nparams < need[kind] || nparams > 8) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_material: wrong number of parameters");
  if (h->have_mesh) {
    const bool solid = kind == NLS_MAT_NEOHOOKEAN || kind == NLS_MAT_J2_SMALL_STRAIN;
    if (solid && h->dim == 1) NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "Neo-Hookean and J2 need a 2-D or 3-D mesh");
    if (!solid && h->dim != 1) NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "1-D materials need a 1-D mesh");
  }
  const double keep_area = h->mat.area;
  std::memset(&h->mat, 0, sizeof(h->mat));
  h->mat.area = keep_area;
  h->mat.kind = kind;
  for (int i = 0; i < nparams; ++i) h->mat.p[i] = params[i];
  if (kind == NLS_MAT_ELASTOPLASTIC_1D && nparams < 6) {  // Eep default, linearelastoplasticity.jl:17
    const double E = params[0], Hk = params[1], Ha = params[2];
    h->mat.p[5] = E * (Hk + Ha) / (E + Hk + Ha);
  }
  h->have_mat = true;
  // a (new) stateful material starts from its initial state, like the reference's initialize_state! -- also when the handle
  // already holds state arrays from a previous material object with other hardening parameters
  if (kind == NLS_MAT_ELASTOPLASTIC_1D && h->d_state[0]) return nls_init_state(h);
  return alloc_state(h);
}

extern "C" int32_t nls_set_area(nls_handle h, double area) {
  NLS_ENTER(h);
  h->mat.area = area;
  return NLS_OK;
}

extern "C" int32_t nls_set_dirichlet(nls_handle h, int64_t n, const int64_t *dofs0, const double *vals) {
  NLS_ENTER(h);
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_dirichlet: call nls_set_mesh first");
  if (n < 0 || (n > 0 && (!dofs0 || !vals))) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_dirichlet: bad arguments");
  const int64_t nd = h->ndof_local();
  for (int64_t k = 0; k < n; ++k)
    if (dofs0[k] < 0 || dofs0[k] >= nd) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_dirichlet: DoF index out of range");
  // sequential apply! semantics: for a repeated DoF the last value wins
  h->dir_user_dofs.assign(dofs0, dofs0 + n);
  h->dir_slot.assign(n, -1);
  std::unordered_map<int64_t, int64_t> last;
  last.reserve((size_t)n * 2);
  for (int64_t k = 0; k < n; ++k) last[dofs0[k]] = k;
  std::vector<int64_t> udofs; std::vector<double> uvals;
  for (int64_t k = 0; k < n; ++k)
    if (last[dofs0[k]] == k) { h->dir_slot[k] = (int64_t)udofs.size(); udofs.push_back(dofs0[k]); uvals.push_back(vals[k]); }
  h->ndir_user = n; h->ndir = (int64_t)udofs.size();
  NLS_TRY(dev_alloc(h, &h->d_dir_dofs, (size_t)h->ndir));
  NLS_TRY(dev_alloc(h, &h->d_dir_vals, (size_t)h->ndir));
  NLS_CUDA(h, cudaMemsetAsync(h->d_con, 0, (size_t)nd, h->stream));
  if (h->ndir > 0) {
    NLS_CUDA(h, cudaMemcpyAsync(h->d_dir_dofs, udofs.data(), sizeof(int64_t) * h->ndir, cudaMemcpyHostToDevice, h->stream));
    NLS_CUDA(h, cudaMemcpyAsync(h->d_dir_vals, uvals.data(), sizeof(double) * h->ndir, cudaMemcpyHostToDevice, h->stream));
    k_set_mask<<<(unsigned)((h->ndir + 255) / 256), 256, 0, h->stream>>>(h->ndir, h->d_dir_dofs, h->d_con);
    h->launches++;
  }
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_set_dirichlet_values(nls_handle h, int64_t n, const double *vals) {
  NLS_ENTER(h);
  if (n != h->ndir_user || (n > 0 && !vals)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_dirichlet_values: length differs from the Dirichlet list");
  std::vector<double> uvals((size_t)h->ndir);
  for (int64_t k = 0; k < n; ++k) if (h->dir_slot[k] >= 0) uvals[h->dir_slot[k]] = vals[k];
  if (h->ndir > 0) {
    NLS_CUDA(h, cudaMemcpyAsync(h->d_dir_vals, uvals.data(), sizeof(double) * h->ndir, cudaMemcpyHostToDevice, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return NLS_OK;
}

extern "C" int32_t nls_set_neumann(nls_handle h, int64_t n, const int64_t *dofs0, const double *vals) {
  NLS_ENTER(h);
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_neumann: call nls_set_mesh first");
  if (n < 0 || (n > 0 && (!dofs0 || !vals))) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_neumann: bad arguments");
  for (int64_t k = 0; k < n; ++k)
    if (dofs0[k] < 0 || dofs0[k] >= h->ndof_local()) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_neumann: DoF index out of range");
  h->nneu = n;
  NLS_TRY(dev_alloc(h, &h->d_neu_dofs, (size_t)n));
  NLS_TRY(dev_alloc(h, &h->d_neu_vals, (size_t)n));
  if (n > 0) {
    NLS_CUDA(h, cudaMemcpyAsync(h->d_neu_dofs, dofs0, sizeof(int64_t) * n, cudaMemcpyHostToDevice, h->stream));
    NLS_CUDA(h, cudaMemcpyAsync(h->d_neu_vals, vals, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return NLS_OK;
}

extern "C" int32_t nls_set_neumann_values(nls_handle h, int64_t n, const double *vals) {
  NLS_ENTER(h);
  if (n != h->nneu || (n > 0 && !vals)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_neumann_values: length differs from the Neumann list");
  if (n > 0) {
    NLS_CUDA(h, cudaMemcpyAsync(h->d_neu_vals, vals, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return NLS_OK;
}

extern "C" int32_t nls_set_body_force(nls_handle h, int32_t ncomp, const double *f) {
  NLS_ENTER(h);
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_body_force: call nls_set_mesh first");
  if (ncomp != 0 && (ncomp != h->dim || !f)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_body_force: need one value per spatial dimension");
  h->have_body = false;
  for (int d = 0; d < 3; ++d) h->body[d] = 0.0;
  for (int d = 0; d < ncomp; ++d) { h->body[d] = f[d]; h->have_body = h->have_body || f[d] != 0.0; }
  return NLS_OK;
}

extern "C" int32_t nls_init_state(nls_handle h) {
  NLS_ENTER(h);
  if (h->mat.kind == NLS_MAT_J2_SMALL_STRAIN) return h->have_mesh ? nls_j2_init(h) : NLS_OK;
  if (h->mat.kind != NLS_MAT_ELASTOPLASTIC_1D) return NLS_OK;
  if (!h->d_state[0]) return alloc_state(h);  // allocates, then re-enters here
  const int64_t n = h->nel * h->nq;
  if (n == 0) return NLS_OK;
  const unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, 4096);
  for (int i = 0; i < 9; ++i) {
    double v = 0.0;
    if (i == 6) v = h->mat.p[4];  // alpha_n <- alpha_0
    if (i == 7) v = h->mat.p[3];  // kappa_n <- kappa_0
    k_fill<<<g, 256, 0, h->stream>>>(n, h->d_state[i], v);
    h->launches++;
  }
  NLS_CUDA(h, cudaMemsetAsync(h->d_u_n, 0, sizeof(double) * h->nel * h->nen, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_commit_state(nls_handle h, const double *U) {
  NLS_ENTER(h);
  if (h->mat.kind == NLS_MAT_J2_SMALL_STRAIN) {      // the trial state of the last internal-force pass becomes the committed one
    NLS_TRY(nls_j2_commit(h));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    return NLS_OK;
  }
  if (h->mat.kind != NLS_MAT_ELASTOPLASTIC_1D) return NLS_OK;
  if (!h->d_state[0]) NLS_FAIL(h, NLS_ERR_STATE, "nls_commit_state: state not initialised");
  if (U) NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(nls_launch_commit_state(h, h->d_U));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_get_state(nls_handle h, const char *name, double *out) {
  NLS_ENTER(h);
  static const char *names[9] = {"sig", "alp", "kap", "gam", "dsig", "sig_n", "alp_n", "kap_n", "gam_n"};
  if (!name || !out) NLS_FAIL(h, NLS_ERR_ARG, "nls_get_state: null argument");
  if (h->mat.kind == NLS_MAT_J2_SMALL_STRAIN) return nls_j2_get(h, name, out);
  if (!h->d_state[0]) NLS_FAIL(h, NLS_ERR_STATE, "nls_get_state: material has no state");
  for (int i = 0; i < 9; ++i)
    if (!std::strcmp(name, names[i])) {
      NLS_CUDA(h, cudaMemcpyAsync(out, h->d_state[i], sizeof(double) * h->nel * h->nq, cudaMemcpyDeviceToHost, h->stream));
      NLS_CUDA(h, cudaStreamSynchronize(h->stream));
      return NLS_OK;
    }
  NLS_FAIL(h, NLS_ERR_ARG, "nls_get_state: unknown state name");
}

// ---- pattern --------------------------------------------------------------------------
extern "C" int32_t nls_build_pattern(nls_handle h) {
  NLS_ENTER(h);
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_build_pattern: call nls_set_mesh first");
  free_pattern(h);
  nls_host_block_pattern(h->nen, h->nel, h->h_conn.data(), h->nn, h->n_owned, h->h_browptr, h->h_bcol);
  h->nnzb = (int64_t)h->h_bcol.size();
  h->nrows = h->n_owned * h->dim;
  h->nnz = h->nnzb * h->dim * h->dim;
  for (int64_t a = 0; a < h->n_owned; ++a)
    if (h->h_browptr[a + 1] - h->h_browptr[a] > 65535) NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "node with more than 65535 neighbours");
  NLS_TRY(dev_alloc(h, &h->d_browptr, (size_t)(h->n_owned + 1)));
  NLS_TRY(dev_alloc(h, &h->d_bcol, (size_t)h->nnzb));
  NLS_TRY(dev_alloc(h, &h->d_rowptr, (size_t)(h->nrows + 1)));
  NLS_TRY(dev_alloc(h, &h->d_colind, (size_t)h->nnz));
  NLS_TRY(dev_alloc(h, &h->d_vals, (size_t)h->nnz));
  NLS_TRY(dev_alloc(h, &h->d_diag, (size_t)h->nrows));
  NLS_TRY(dev_alloc(h, &h->d_pos, (size_t)(h->nel * h->nen * h->nen)));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_browptr, h->h_browptr.data(), sizeof(int64_t) * (h->n_owned + 1), cudaMemcpyHostToDevice, h->stream));
  if (h->nnzb) NLS_CUDA(h, cudaMemcpyAsync(h->d_bcol, h->h_bcol.data(), sizeof(int32_t) * h->nnzb, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_vals, 0, sizeof(double) * h->nnz, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_diag, 0, sizeof(int32_t) * h->nrows, h->stream));
  NLS_TRY(nls_launch_expand_pattern(h));
  NLS_TRY(nls_launch_build_pos(h));
  NLS_TRY(nls_sym_build_index(h));
  h->vals_version++;
  // plan of the gather (row-owner) assembly: 2-D only for now; falls back to the atomic scatter
  // when a cluster would not fit in shared memory
  if (h->dim == 2 && h->n_owned > 0 && h->nel > 0) {
    AsmPlan plan;
    int tile[3];
    nls_gather_tile(tile);
    nls_host_cluster_plan(h->dim, h->nen, h->nel, h->h_conn.data(), h->nn, h->n_owned, h->h_coords.data(), nls_gather_nc(), plan, tile);
    NLS_TRY(nls_gather_build(h, plan));
  }
  NLS_TRY(nls_twophase_build(h));
  NLS_TRY(h->dim == 2 ? nls_lines2_build(h) : nls_lines3_build(h));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_pattern = true;
  return NLS_OK;
}

extern "C" int32_t nls_pattern_size(nls_handle h, int64_t *nrows, int64_t *nnz) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_pattern_size: call nls_build_pattern first");
  if (nrows) *nrows = h->nrows;
  if (nnz) *nnz = h->nnz;
  return NLS_OK;
}

extern "C" int32_t nls_get_pattern(nls_handle h, int64_t *rowptr, int32_t *colind) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_get_pattern: call nls_build_pattern first");
  if (rowptr) NLS_CUDA(h, cudaMemcpyAsync(rowptr, h->d_rowptr, sizeof(int64_t) * (h->nrows + 1), cudaMemcpyDeviceToHost, h->stream));
  if (colind && h->nnz) NLS_CUDA(h, cudaMemcpyAsync(colind, h->d_colind, sizeof(int32_t) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

// ---- operators ------------------------------------------------------------------------
static int check_ready(nls_context *h, const char *who) {
  if (!h->have_mesh || !h->have_disc || !h->have_mat || !h->have_pattern) {
    h->err = std::string(who) + ": model incomplete (need mesh, discretization, material and pattern)";
    return NLS_ERR_STATE;
  }
  if ((h->mat.kind == NLS_MAT_NEOHOOKEAN || h->mat.kind == NLS_MAT_J2_SMALL_STRAIN) != (h->dim > 1)) {
    h->err = std::string(who) + ": material kind does not match the mesh dimension";
    return NLS_ERR_UNSUPPORTED;
  }
  return NLS_OK;
}

// (::Residual)(fem,U,R,ctx) and/or (::Jacobian)(fem,U,K,ctx) on the device-resident U
// pure = compute_internalforce / compute_dinternalforce (defaults.jl:29-46, 68-82) alone: U is left as given (no Dirichlet
// overwrite) and no external force is subtracted -- what the material-level dispatch point of the reference means
static int assemble(nls_context *h, bool want_r, bool want_k, bool pure = false) {
  if (pure) {
    NLS_TRY(nls_halo_exchange(h, h->d_U));
    if (!nls_assembly_overwrites(h)) {
      if (want_r) NLS_CUDA(h, cudaMemsetAsync(h->d_R, 0, sizeof(double) * h->nrows, h->stream));
      if (want_k) NLS_CUDA(h, cudaMemsetAsync(h->d_vals, 0, sizeof(double) * h->nnz, h->stream));
    }
    NLS_TRY(nls_launch_assembly(h, want_r, want_k));
    if (want_k) h->vals_version++;
    return NLS_OK;
  }
  NLS_TRY(nls_launch_apply_dirichlet(h, h->d_U));              // residual.jl:13 / jacobian.jl:11
  // multi-GPU, 2-D row-owner kernel: the ghost exchange runs on a second stream while the clusters that touch no
  // ghost node are assembled; the few clusters at the slab faces follow once it has landed
  if (h->distributed && h->nneigh > 0 && h->opt_overlap_halo && nls_assembly_mode(h) == NLS_ASM_GATHER &&
      h->gplan.ncl_interior > 0 && (want_r || want_k) && h->nel > 0) {
    NLS_CUDA(h, cudaEventRecord(h->ev_halo0, h->stream));
    NLS_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_halo0, 0));
    h->halo_stream = h->comm_stream;
    const int rc = nls_halo_exchange(h, h->d_U);
    h->halo_stream = nullptr;
    if (rc) return rc;
    NLS_CUDA(h, cudaEventRecord(h->ev_halo1, h->comm_stream));
    NLS_TRY(nls_launch_gather_2d(h, want_r, want_k, 0, h->gplan.ncl_interior));
    NLS_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo1, 0));
    NLS_TRY(nls_launch_gather_2d(h, want_r, want_k, h->gplan.ncl_interior, h->gplan.ncl));
    if (want_k) h->vals_version++;
    if (want_r) NLS_TRY(nls_launch_body_force(h, h->d_R));
    if (want_r) NLS_TRY(nls_launch_neumann(h, h->d_R));
    return NLS_OK;
  }
  NLS_TRY(nls_halo_exchange(h, h->d_U));
  if (!nls_assembly_overwrites(h)) {   // the gather kernel writes every row itself
    if (want_r) NLS_CUDA(h, cudaMemsetAsync(h->d_R, 0, sizeof(double) * h->nrows, h->stream));   // residual.jl:14
    if (want_k) NLS_CUDA(h, cudaMemsetAsync(h->d_vals, 0, sizeof(double) * h->nnz, h->stream));  // jacobian.jl:13
  }
  NLS_TRY(nls_launch_assembly(h, want_r, want_k));
  if (want_k) h->vals_version++;
  if (want_r) NLS_TRY(nls_launch_body_force(h, h->d_R));       // residual.jl:19-23 (DistributedForce)
  if (want_r) NLS_TRY(nls_launch_neumann(h, h->d_R));          // residual.jl:24-25
  return NLS_OK;
}

// Newton on the acceleration (newmark.jl:75-107): U, Ud follow from the unknown Udd (update_state!), the static
// operators act on that U, then the inertia terms: R += M Udd (residual.jl:53), K = (dt^2 beta) K + M (jacobian.jl:46-50).
static int assemble_any(nls_context *h, bool want_r, bool want_k) {
  if (!h->dyn_on) return assemble(h, want_r, want_k);
  NLS_TRY(nls_halo_exchange(h, h->d_Udd));
  NLS_TRY(nls_launch_newmark_state(h, h->dyn_cu, h->dyn_cv));
  NLS_TRY(assemble(h, want_r, want_k));
  if (want_r) NLS_TRY(nls_launch_apply_mass(h, h->d_Udd, h->d_R, true));
  if (want_k) NLS_TRY(nls_launch_scale_add_mass(h, h->dyn_cu, true));
  return NLS_OK;
}

static int host_assemble(nls_context *h, double *U, double *R, double *vals) {
  NLS_TRY(check_ready(h, "assemble"));
  if (!U) NLS_FAIL(h, NLS_ERR_ARG, "assemble: null U");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(assemble(h, R != nullptr, vals != nullptr));
  NLS_CUDA(h, cudaMemcpyAsync(U, h->d_U, sizeof(double) * h->ndof_local(), cudaMemcpyDeviceToHost, h->stream));
  if (R) NLS_CUDA(h, cudaMemcpyAsync(R, h->d_R, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  if (vals) NLS_CUDA(h, cudaMemcpyAsync(vals, h->d_vals, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_residual(nls_handle h, double *U, double *R) {
  NLS_ENTER(h);
  if (!R) NLS_FAIL(h, NLS_ERR_ARG, "nls_residual: null R");
  return host_assemble(h, U, R, nullptr);
}
extern "C" int32_t nls_jacobian(nls_handle h, double *U, double *vals) {
  NLS_ENTER(h);
  if (!vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_jacobian: null vals");
  return host_assemble(h, U, nullptr, vals);
}
extern "C" int32_t nls_residual_jacobian(nls_handle h, double *U, double *R, double *vals) {
  NLS_ENTER(h);
  if (!R || !vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_residual_jacobian: null output");
  return host_assemble(h, U, R, vals);
}

// compute_internalforce(material, fem, U, Fint, ctx; mode=:set) / compute_dinternalforce(...; mode=:set) -- the element
// loops alone (defaults.jl:29-46, 68-82): U is NOT modified, no Dirichlet values are applied, no external force enters.
// For stateful materials the internal-force pass updates the trial state like the reference (defaults.jl:40-42).
extern "C" int32_t nls_internal_force(nls_handle h, const double *U, double *Fint) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_internal_force"));
  if (!U || !Fint) NLS_FAIL(h, NLS_ERR_ARG, "nls_internal_force: null argument");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(assemble(h, true, false, true));
  NLS_CUDA(h, cudaMemcpyAsync(Fint, h->d_R, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_internal_force_jacobian(nls_handle h, const double *U, double *vals) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_internal_force_jacobian"));
  if (!U || !vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_internal_force_jacobian: null argument");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(assemble(h, false, true, true));
  NLS_CUDA(h, cudaMemcpyAsync(vals, h->d_vals, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

// ---- Krylov hooks -----------------------------------------------------------------------
extern "C" int32_t nls_set_values(nls_handle h, const double *vals) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_values: call nls_build_pattern first");
  if (!vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_values: null vals");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_vals, vals, sizeof(double) * h->nnz, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->vals_version++;
  return NLS_OK;
}

extern "C" int32_t nls_spmv(nls_handle h, const double *x, double *y) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_spmv: call nls_build_pattern first");
  if (!x || !y) NLS_FAIL(h, NLS_ERR_ARG, "nls_spmv: null vector");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_p, x, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(nls_halo_exchange(h, h->d_p));
  NLS_TRY(nls_launch_spmv(h, h->d_p, h->d_Ap, false, false));
  NLS_CUDA(h, cudaMemcpyAsync(y, h->d_Ap, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_dot(nls_handle h, int64_t n, const double *x, const double *y, double *out) {
  NLS_ENTER(h);
  if (!h->have_mesh || n < 0 || n > h->ndof_local() || !x || !y || !out) NLS_FAIL(h, NLS_ERR_ARG, "nls_dot: bad arguments (n must be <= local DoFs)");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_tmp, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_b, y, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  return nls_dev_dot(h, n, h->d_tmp, h->d_b, out);
}

extern "C" int32_t nls_axpy(nls_handle h, int64_t n, double a, const double *x, double *y) {
  NLS_ENTER(h);
  if (!h->have_mesh || n < 0 || n > h->ndof_local() || !x || !y) NLS_FAIL(h, NLS_ERR_ARG, "nls_axpy: bad arguments (n must be <= local DoFs)");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_tmp, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_b, y, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(nls_launch_axpy(h, n, a, h->d_tmp, h->d_b));
  NLS_CUDA(h, cudaMemcpyAsync(y, h->d_b, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_linear_solve(nls_handle h, double rtol, int32_t maxits, const double *b, double *x,
                                    int32_t *its, double *relres, int32_t *status_out) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_linear_solve: call nls_build_pattern first");
  if (!b || !x || maxits < 0) NLS_FAIL(h, NLS_ERR_ARG, "nls_linear_solve: bad arguments");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_b, b, sizeof(double) * h->nrows, cudaMemcpyHostToDevice, h->stream));
  int li = 0, st = 0; double rr = 0.0;
  NLS_TRY(nls_pcg_solve(h, rtol, maxits, 0, &li, &rr, &st));
  NLS_CUDA(h, cudaMemcpyAsync(x, h->d_dU, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  if (its) *its = li;
  if (relres) *relres = rr;
  if (status_out) *status_out = st;
  return NLS_OK;
}

// symmetric indefinite variant (arc-length continuation past limit points): Jacobi-preconditioned MINRES
extern "C" int32_t nls_linear_solve_indefinite(nls_handle h, double rtol, int32_t maxits, const double *b, double *x,
                                               int32_t *its, double *relres, int32_t *status_out) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_linear_solve_indefinite: call nls_build_pattern first");
  if (!b || !x || maxits < 0) NLS_FAIL(h, NLS_ERR_ARG, "nls_linear_solve_indefinite: bad arguments");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_b, b, sizeof(double) * h->nrows, cudaMemcpyHostToDevice, h->stream));
  int li = 0, st = 0; double rr = 0.0;
  NLS_TRY(nls_minres_solve(h, rtol, maxits, &li, &rr, &st));
  NLS_CUDA(h, cudaMemcpyAsync(x, h->d_dU, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  if (its) *its = li;
  if (relres) *relres = rr;
  if (status_out) *status_out = st;
  return NLS_OK;
}

// ---- Newton ---------------------------------------------------------------------------
static float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }

// One linear solve of K dU = -R on the free block, dU = 0 on constrained DoFs (newton_fem.jl:158)
static int newton_linear(nls_context *h, double lin_rtol, int lin_maxits, int *lin_its, int *status) {
  const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((h->nrows + 255) / 256, (int64_t)h->sm_count * 8));
  k_negate_into<<<g, 256, 0, h->stream>>>(h->nrows, h->d_R, h->d_b);
  h->launches++;
  // small systems: the reference's `\` is exact; so is this (banded LU, kernels_direct.cu)
  if (h->opt_direct && nls_direct_available(h)) { if (lin_its) *lin_its = 1; return nls_direct_solve(h, status); }
  double rr;
  return nls_pcg_solve(h, lin_rtol, lin_maxits, 0, lin_its, &rr, status);
}

// solve!(solver::NewtonSolver, U0) on the device-resident U (h->d_U holds U0 on entry)
static int newton_loop(nls_context *h, const nls_newton_opts *o, double *hist, nls_newton_result *res) {
  NLS_TRY(check_ready(h, "nls_newton_solve"));
  if (!o || !res || o->maxits < 0) NLS_FAIL(h, NLS_ERR_ARG, "nls_newton_solve: bad options");
  std::memset(res, 0, sizeof(*res));
  int rc = NLS_OK, nh = 0, its = 0, lin_total = 0;
  NLS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  NLS_CUDA(h, cudaEventRecord(h->ev2, h->stream));
  NLS_TRY(assemble_any(h, true, true));                                    // :137-138
  NLS_CUDA(h, cudaEventRecord(h->ev3, h->stream));
  double norm_r = nls_dev_masked_norm2(h, h->d_R, true, &rc);              // :140-141
  if (rc) return rc;
  res->ms_assembly += elapsed(h->ev2, h->ev3);
  const double norm_r0 = norm_r > 2.220446049250313e-16 ? norm_r : 1.0;    // :143
  int reason = NLS_REASON_NOT_YET_CONVERGED;
  if (hist) hist[nh++] = norm_r;
  bool finished = false;
  while (its < o->maxits) {                                                // :145
    if (norm_r / norm_r0 < o->rtol) { reason = NLS_REASON_CONVERGED_RELATIVE; break; }   // :146-150
    if (norm_r < o->atol) { reason = NLS_REASON_CONVERGED_ABSOLUTE; break; }             // :151-155
    int li = 0, st = 0;
    NLS_CUDA(h, cudaEventRecord(h->ev2, h->stream));
    NLS_TRY(newton_linear(h, o->lin_rtol, o->lin_maxits, &li, &st));       // :158
    NLS_CUDA(h, cudaEventRecord(h->ev3, h->stream));
    lin_total += li;
    if (st != 0) { reason = NLS_REASON_DIVERGED_LINEAR_SOLVE; finished = true; }         // :159-167
    double nd = 0.0;
    if (!finished) {
      nd = nls_dev_masked_norm2(h, h->d_dU, false, &rc);                   // norm(dU), :168
      if (rc) return rc;
    } else NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    res->ms_linear += elapsed(h->ev2, h->ev3);
    if (finished) break;
    if (nd > o->dtol) { reason = NLS_REASON_DIVERGED_STEP; finished = true; break; }     // :168-172
    NLS_TRY(nls_launch_axpy(h, h->nrows, 1.0, h->d_dU, h->dyn_on ? h->d_Udd : h->d_U));   // :173
    NLS_CUDA(h, cudaEventRecord(h->ev2, h->stream));
    NLS_TRY(assemble_any(h, true, !o->type_modified));                     // :174-179
    NLS_CUDA(h, cudaEventRecord(h->ev3, h->stream));
    norm_r = nls_dev_masked_norm2(h, h->d_R, true, &rc);                   // :180
    if (rc) return rc;
    res->ms_assembly += elapsed(h->ev2, h->ev3);
    its++;                                                                 // :181
    if (hist) hist[nh++] = norm_r;
  }
  if (!finished && its >= o->maxits) reason = NLS_REASON_DIVERGED_MAXITS;  // :184-187
  NLS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  NLS_TRY(nls_p2p_failed(h));
  res->ms_total = elapsed(h->ev0, h->ev1);
  res->reason = reason; res->iterations = its; res->lin_its_total = lin_total; res->nhist = nh;
  res->norm_r = norm_r; res->norm_r0 = norm_r0;
  return NLS_OK;
}

extern "C" int32_t nls_newton_solve(nls_handle h, const nls_newton_opts *opts, const double *U0, double *U, double *R,
                                    double *norm_hist, nls_newton_result *res) {
  NLS_ENTER(h);
  if (!U0 || !U) NLS_FAIL(h, NLS_ERR_ARG, "nls_newton_solve: null U0/U");
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_newton_solve: model incomplete");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U0, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));  // :135
  NLS_TRY(newton_loop(h, opts, norm_hist, res));
  NLS_TRY(nls_halo_exchange(h, h->d_U));
  NLS_CUDA(h, cudaMemcpyAsync(U, h->d_U, sizeof(double) * h->ndof_local(), cudaMemcpyDeviceToHost, h->stream));
  if (R) NLS_CUDA(h, cudaMemcpyAsync(R, h->d_R, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

// ---- dynamics: consistent mass, dynamic operators, Newmark step ---------------------
static int ensure_dyn(nls_context *h) {
  if (h->d_Udd) return NLS_OK;
  const size_t nd = (size_t)h->ndof_local();
  NLS_TRY(dev_alloc(h, &h->d_Ut, nd)); NLS_TRY(dev_alloc(h, &h->d_Udt, nd));
  NLS_TRY(dev_alloc(h, &h->d_Udd, nd)); NLS_TRY(dev_alloc(h, &h->d_Ud, nd));
  NLS_TRY(dev_alloc(h, &h->d_con_zero, nd));
  NLS_CUDA(h, cudaMemsetAsync(h->d_Ut, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_Udt, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_Udd, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_Ud, 0, nd * sizeof(double), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(h->d_con_zero, 0, nd, h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_set_mass(nls_handle h, double rho, double area) {
  NLS_ENTER(h);
  if (!h->have_pattern || !h->have_disc) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_mass: call nls_build_pattern first");
  if (!(rho > 0.0) || !(area > 0.0)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_mass: density and area must be positive");
  NLS_TRY(ensure_dyn(h));
  NLS_TRY(dev_alloc(h, &h->d_mass, (size_t)h->nnzb));
  NLS_TRY(nls_launch_mass_build(h, rho, h->dim == 1 ? area : 1.0));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->mass_rho = rho; h->mass_area = area; h->have_mass = true;
  return NLS_OK;
}

#define NEED_MASS(h, who) do { if (!(h)->have_mass) NLS_FAIL(h, NLS_ERR_STATE, who ": call nls_set_mass first"); } while (0)

extern "C" int32_t nls_get_mass(nls_handle h, double *vals) {
  NLS_ENTER(h);
  NEED_MASS(h, "nls_get_mass");
  if (!vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_get_mass: null vals");
  NLS_TRY(nls_launch_scale_add_mass(h, 0.0, false));
  NLS_CUDA(h, cudaMemcpyAsync(vals, h->d_vals, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_apply_mass(nls_handle h, const double *x, double *y) {
  NLS_ENTER(h);
  NEED_MASS(h, "nls_apply_mass");
  if (!x || !y) NLS_FAIL(h, NLS_ERR_ARG, "nls_apply_mass: null vector");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_tmp, x, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_b, y, sizeof(double) * h->nrows, cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(nls_halo_exchange(h, h->d_tmp));
  NLS_TRY(nls_launch_apply_mass(h, h->d_tmp, h->d_b, true));
  NLS_CUDA(h, cudaMemcpyAsync(y, h->d_b, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_dynamic_residual(nls_handle h, double *U, const double *Udd, double *R) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_dynamic_residual"));
  NEED_MASS(h, "nls_dynamic_residual");
  if (!U || !Udd || !R) NLS_FAIL(h, NLS_ERR_ARG, "nls_dynamic_residual: null argument");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_Udd, Udd, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(assemble(h, true, false));
  NLS_TRY(nls_halo_exchange(h, h->d_Udd));
  NLS_TRY(nls_launch_apply_mass(h, h->d_Udd, h->d_R, true));
  NLS_CUDA(h, cudaMemcpyAsync(U, h->d_U, sizeof(double) * h->ndof_local(), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(R, h->d_R, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

extern "C" int32_t nls_dynamic_jacobian(nls_handle h, double *U, double scale, double *vals) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_dynamic_jacobian"));
  NEED_MASS(h, "nls_dynamic_jacobian");
  if (!U || !vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_dynamic_jacobian: null argument");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_TRY(assemble(h, false, true));
  NLS_TRY(nls_launch_scale_add_mass(h, scale, true));
  NLS_CUDA(h, cudaMemcpyAsync(U, h->d_U, sizeof(double) * h->ndof_local(), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(vals, h->d_vals, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

// M x = b on ALL DoFs (newmark.jl:136-140 solves with the full mass matrix, no constraints): Jacobi-PCG
extern "C" int32_t nls_mass_solve(nls_handle h, double rtol, int32_t maxits, const double *b, double *x, int32_t *its, int32_t *status_out) {
  NLS_ENTER(h);
  NEED_MASS(h, "nls_mass_solve");
  if (!b || !x || maxits < 0) NLS_FAIL(h, NLS_ERR_ARG, "nls_mass_solve: bad arguments");
  NLS_TRY(nls_launch_scale_add_mass(h, 0.0, false));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_b, b, sizeof(double) * h->nrows, cudaMemcpyHostToDevice, h->stream));
  std::swap(h->d_con, h->d_con_zero);
  int li = 0, st = 0; double rr = 0.0;
  const int rc = nls_pcg_solve(h, rtol, maxits, 0, &li, &rr, &st);
  std::swap(h->d_con, h->d_con_zero);
  if (rc) return rc;
  NLS_CUDA(h, cudaMemcpyAsync(x, h->d_dU, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  if (its) *its = li;
  if (status_out) *status_out = st;
  return NLS_OK;
}

// One Newmark step: solve!(ts.solver, Udd_prev) with the closures of newmark.jl:75-107 installed, then
// update_state! (newmark.jl:114-117). Inputs: predictors Ut, Udt (update_predictors!, newmark.jl:109-112) and the
// initial guess Udd0; outputs: the converged acceleration and U, Ud = update_state!(Udd) (without the Dirichlet
// overwrite: the reference applies it to a temporary copy, newmark.jl:78,103).
extern "C" int32_t nls_newmark_solve(nls_handle h, const nls_newton_opts *opts, double beta, double gamma, double dt,
                                     const double *Ut, const double *Udt, const double *Udd0,
                                     double *Udd, double *U, double *Ud, double *R, double *norm_hist, nls_newton_result *res) {
  NLS_ENTER(h);
  NEED_MASS(h, "nls_newmark_solve");
  if (!Ut || !Udt || !Udd0 || !Udd || !U || !Ud) NLS_FAIL(h, NLS_ERR_ARG, "nls_newmark_solve: null vector");
  const size_t nb = sizeof(double) * h->ndof_local();
  NLS_CUDA(h, cudaMemcpyAsync(h->d_Ut, Ut, nb, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_Udt, Udt, nb, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_Udd, Udd0, nb, cudaMemcpyHostToDevice, h->stream));
  h->dyn_on = true; h->dyn_cu = dt * dt * beta; h->dyn_cv = dt * gamma;
  const int rc = newton_loop(h, opts, norm_hist, res);
  h->dyn_on = false;
  if (rc) return rc;
  NLS_TRY(nls_halo_exchange(h, h->d_Udd));
  NLS_TRY(nls_launch_newmark_state(h, h->dyn_cu, h->dyn_cv));
  NLS_CUDA(h, cudaMemcpyAsync(Udd, h->d_Udd, nb, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(U, h->d_U, nb, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(Ud, h->d_Ud, nb, cudaMemcpyDeviceToHost, h->stream));
  if (R) NLS_CUDA(h, cudaMemcpyAsync(R, h->d_R, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

// ---- device-resident variants --------------------------------------------------------------
extern "C" int32_t nls_dev_upload_u(nls_handle h, const double *U) {
  NLS_ENTER(h);
  if (!h->have_mesh || !U) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_upload_u: no mesh or null U");
  NLS_CUDA(h, cudaMemcpyAsync(h->d_U, U, sizeof(double) * h->ndof_local(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_dev_download_u(nls_handle h, double *U) {
  NLS_ENTER(h);
  if (!h->have_mesh || !U) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_download_u: no mesh or null U");
  NLS_CUDA(h, cudaMemcpyAsync(U, h->d_U, sizeof(double) * h->ndof_local(), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_dev_download_r(nls_handle h, double *R) {
  NLS_ENTER(h);
  if (!h->have_pattern || !R) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_download_r: no pattern or null R");
  NLS_CUDA(h, cudaMemcpyAsync(R, h->d_R, sizeof(double) * h->nrows, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_dev_download_vals(nls_handle h, double *vals) {
  NLS_ENTER(h);
  if (!h->have_pattern || !vals) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_download_vals: no pattern or null vals");
  NLS_CUDA(h, cudaMemcpyAsync(vals, h->d_vals, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_dev_download_rows(nls_handle h, int64_t row0, int64_t row1, double *R, int64_t *rowptr, int32_t *colind, double *vals) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_dev_download_rows: call nls_build_pattern first");
  if (row0 < 0 || row1 < row0 || row1 > h->nrows) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_download_rows: row range outside the owned rows");
  std::vector<int64_t> rp((size_t)(row1 - row0 + 1));
  NLS_CUDA(h, cudaMemcpyAsync(rp.data(), h->d_rowptr + row0, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  const int64_t e0 = rp.front(), e1 = rp.back();
  if (rowptr) std::copy(rp.begin(), rp.end(), rowptr);
  if (R && row1 > row0) NLS_CUDA(h, cudaMemcpyAsync(R, h->d_R + row0, sizeof(double) * (row1 - row0), cudaMemcpyDeviceToHost, h->stream));
  if (colind && e1 > e0) NLS_CUDA(h, cudaMemcpyAsync(colind, h->d_colind + e0, sizeof(int32_t) * (e1 - e0), cudaMemcpyDeviceToHost, h->stream));
  if (vals && e1 > e0) NLS_CUDA(h, cudaMemcpyAsync(vals, h->d_vals + e0, sizeof(double) * (e1 - e0), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_dev_assemble(nls_handle h, int32_t want_residual, int32_t want_jacobian) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_dev_assemble"));
  return assemble(h, want_residual != 0, want_jacobian != 0);
}
extern "C" int32_t nls_dev_spmv(nls_handle h) {
  NLS_ENTER(h);
  if (!h->have_pattern) NLS_FAIL(h, NLS_ERR_STATE, "nls_dev_spmv: call nls_build_pattern first");
  return nls_launch_spmv(h, h->d_U, h->d_Ap, false, false);
}
extern "C" int32_t nls_dev_pcg_iterations(nls_handle h, int32_t niter) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_dev_pcg_iterations"));
  if (niter < 1) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_pcg_iterations: niter must be >= 1");
  const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((h->nrows + 255) / 256, (int64_t)h->sm_count * 8));
  k_negate_into<<<g, 256, 0, h->stream>>>(h->nrows, h->d_R, h->d_b);
  h->launches++;
  int li, st; double rr;
  return nls_pcg_solve(h, 0.0, niter, niter, &li, &rr, &st);
}
extern "C" int32_t nls_dev_newton_step(nls_handle h, double lin_rtol, int32_t lin_maxits, int32_t *lin_its, double *norm_r) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_dev_newton_step"));
  int rc = NLS_OK, li = 0, st = 0;
  NLS_TRY(assemble(h, true, true));
  NLS_TRY(newton_linear(h, lin_rtol, lin_maxits, &li, &st));
  NLS_TRY(nls_launch_axpy(h, h->nrows, 1.0, h->d_dU, h->d_U));
  NLS_TRY(assemble(h, true, false));
  const double nr = nls_dev_masked_norm2(h, h->d_R, true, &rc);
  if (rc) return rc;
  if (lin_its) *lin_its = li;
  if (norm_r) *norm_r = nr;
  return NLS_OK;
}
extern "C" int32_t nls_dev_newton_solve(nls_handle h, const nls_newton_opts *opts, double *norm_hist, nls_newton_result *res) {
  NLS_ENTER(h);
  return newton_loop(h, opts, norm_hist, res);
}
extern "C" int32_t nls_sync(nls_handle h) {
  NLS_ENTER(h);
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return nls_p2p_check(h);
}
extern "C" int32_t nls_timer_start(nls_handle h) {
  NLS_ENTER(h);
  NLS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  return NLS_OK;
}
extern "C" int32_t nls_timer_stop(nls_handle h, float *ms) {
  NLS_ENTER(h);
  NLS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  NLS_CUDA(h, cudaEventSynchronize(h->ev1));
  if (ms) *ms = elapsed(h->ev0, h->ev1);
  return NLS_OK;
}
extern "C" int32_t nls_flush_l2(nls_handle h) {
  NLS_ENTER(h);
  if (!h->d_flush) {
    h->flush_bytes = (size_t)256 << 20;  // > 126 MB L2
    NLS_CUDA(h, cudaMalloc(&h->d_flush, h->flush_bytes));
  }
  NLS_CUDA(h, cudaMemsetAsync(h->d_flush, 0x5a, h->flush_bytes, h->stream));
  return NLS_OK;
}
extern "C" int64_t nls_kernel_launches(nls_handle h) { return h ? h->launches : 0; }

extern "C" int32_t nls_assembly_scheme(nls_handle h) {
  if (!h || !h->have_pattern) return -1;
  cudaSetDevice(h->device);
  return nls_assembly_mode(h);
}

extern "C" int32_t nls_debug_counters(nls_handle h, int64_t *out8) {
  NLS_ENTER(h);
  if (!out8) NLS_FAIL(h, NLS_ERR_ARG, "nls_debug_counters: null output");
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  if (h->d_prof) {
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    NLS_CUDA(h, cudaMemcpy(out8, h->d_prof, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  return NLS_OK;
}

// the 3-D line kernel's per-warp phase-B cycles follow the 8 shared counters: 24 values in all
// which preconditioner the next Krylov solve of this model uses: 0 Jacobi, 1 block-Jacobi, 2 multigrid
extern "C" int32_t nls_krylov_preconditioner(nls_handle h) {
  if (!h) return -1;
  bool use = false;                                  // collective when distributed (one small allreduce)
  if (nls_mg_agreed(h, &use) != NLS_OK) return -1;
  return use ? (h->opt_precond == 1 ? 1 : 2) : 0;
}

// Allocates what the first linear solve would otherwise allocate inside its timed region (the multigrid hierarchy of a
// grid-topology 3-D mesh). Collective after nls_comm_init. Optional: the solvers do it themselves on first use.
extern "C" int32_t nls_solver_prepare(nls_handle h) {
  NLS_ENTER(h);
  NLS_TRY(check_ready(h, "nls_solver_prepare"));
  bool use = false;
  NLS_TRY(nls_mg_agreed(h, &use));
  return use ? nls_mg_prepare(h, true) : NLS_OK;
}

extern "C" int32_t nls_debug_counters_ext(nls_handle h, int64_t *out24) {
  NLS_ENTER(h);
  if (!out24) NLS_FAIL(h, NLS_ERR_ARG, "nls_debug_counters_ext: null output");
  for (int i = 0; i < 24; ++i) out24[i] = 0;
  if (h->d_prof) {
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    NLS_CUDA(h, cudaMemcpy(out24, h->d_prof, 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  return NLS_OK;
}

extern "C" int32_t nls_set_option(nls_handle h, const char *key, int32_t value) {
  NLS_ENTER(h);
  if (!key) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_option: null key");
  if (!std::strcmp(key, "assembly")) { h->opt_assembly = value; return NLS_OK; }
  // preconditioner of the Krylov solve: -1 auto, 0 Jacobi, 1 block-Jacobi, 2 multigrid V-cycle (1, 2: grid-topology 3-D meshes)
  if (!std::strcmp(key, "precond")) { h->opt_precond = value; return NLS_OK; }
  if (!std::strcmp(key, "direct")) { h->opt_direct = value; return NLS_OK; }        // 0: Newton's linear solves are always Krylov
  if (!std::strcmp(key, "mg_degree")) { h->mg.degree = value < 1 ? 1 : value; return NLS_OK; }
  if (!std::strcmp(key, "mg_coarse_degree")) { h->mg.coarse_degree = value < 1 ? 1 : value; return NLS_OK; }
  if (!std::strcmp(key, "mg_ratio_x10")) { h->mg.ratio = value < 11 ? 1.1 : value / 10.0; return NLS_OK; }
  if (!std::strcmp(key, "mg_graph")) { h->mg.use_graph = value != 0; h->mg.graph_off = false; return NLS_OK; }
  if (!std::strcmp(key, "mg_max_levels")) { h->mg.max_levels = value < 2 ? 2 : value; nls_mg_invalidate(h); return NLS_OK; }
  if (!std::strcmp(key, "mg_coarse_ratio")) { h->mg.coarse_ratio = value < 2 ? 2.0 : (double)value; return NLS_OK; }
  if (!std::strcmp(key, "mg_fp32")) { h->mg.fp32 = value; h->mg.version = -1; return NLS_OK; }
  if (!std::strcmp(key, "mg_power_its")) { h->mg.power_its = value < 1 ? 1 : value; h->mg.version = -1; return NLS_OK; }
  if (!std::strcmp(key, "spmv")) { h->opt_spmv = value; return NLS_OK; }
  if (!std::strcmp(key, "twophase_3d")) { h->opt_twophase_3d = value; return NLS_OK; }
  if (!std::strcmp(key, "lap3d")) { h->opt_lap3d = value; return NLS_OK; }
  if (!std::strcmp(key, "pcg_graph")) { h->opt_pcg_graph = value; return NLS_OK; }
  if (!std::strcmp(key, "pcg_small")) { h->opt_pcg_small = value; return NLS_OK; }
  if (!std::strcmp(key, "sym_spmv")) { h->opt_sym_spmv = value; return NLS_OK; }
  if (!std::strcmp(key, "overlap_halo")) { h->opt_overlap_halo = value; return NLS_OK; }
  if (!std::strcmp(key, "scratch_mb")) {   // takes effect at the next nls_build_pattern
    h->opt_scratch_mb = value;
    if (h->have_pattern) { NLS_TRY(nls_twophase_build(h)); }
    return NLS_OK;
  }
  if (!std::strcmp(key, "gather_ctas")) { h->opt_gather_ctas = value; return NLS_OK; }
  if (!std::strcmp(key, "profile_phases")) {
    if (!h->d_prof) NLS_CUDA(h, cudaMalloc((void **)&h->d_prof, 24 * sizeof(uint64_t)));
    NLS_CUDA(h, cudaMemset(h->d_prof, 0, 24 * sizeof(uint64_t)));
    h->opt_profile_phases = value;
    return NLS_OK;
  }
  if (!std::strcmp(key, "spmv_lanes")) { h->opt_spmv_lanes = value; return NLS_OK; }
  NLS_FAIL(h, NLS_ERR_ARG, "nls_set_option: unknown key");
}

// ---- multi-GPU ----------------------------------------------------------------------------
extern "C" int32_t nls_set_halo(nls_handle h, int64_t n_owned_nodes, int32_t nneigh, const int32_t *neigh_ranks,
                                const int64_t *send_ptr, const int32_t *send_nodes,
                                const int64_t *recv_ptr, const int32_t *recv_nodes) {
  NLS_ENTER(h);
  if (!h->have_mesh) NLS_FAIL(h, NLS_ERR_STATE, "nls_set_halo: call nls_set_mesh first");
  if (n_owned_nodes < 0 || n_owned_nodes > h->nn || nneigh < 0) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: bad sizes");
  if (nneigh > 0 && (!neigh_ranks || !send_ptr || !recv_ptr)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: null lists");
  free_pattern(h);
  h->n_owned = n_owned_nodes;
  h->nneigh = nneigh;
  h->neigh.assign(neigh_ranks, neigh_ranks + nneigh);
  h->send_ptr.assign(1, 0); h->recv_ptr.assign(1, 0);
  if (nneigh > 0) { h->send_ptr.assign(send_ptr, send_ptr + nneigh + 1); h->recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1); }
  const int64_t ns = h->send_ptr.back(), nr = h->recv_ptr.back();
  if ((ns > 0 && !send_nodes) || (nr > 0 && !recv_nodes)) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: null node list with a positive count");
  for (int64_t k = 0; k < ns; ++k) if (send_nodes[k] < 0 || send_nodes[k] >= h->n_owned) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: send node is not an owned node");
  for (int64_t k = 0; k < nr; ++k) if (recv_nodes[k] < h->n_owned || recv_nodes[k] >= h->nn) NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: recv node is not a ghost node");
  NLS_TRY(dev_alloc(h, &h->d_send_nodes, (size_t)ns)); NLS_TRY(dev_alloc(h, &h->d_recv_nodes, (size_t)nr));
  NLS_TRY(dev_alloc(h, &h->d_sendbuf, (size_t)(ns * h->dim))); NLS_TRY(dev_alloc(h, &h->d_recvbuf, (size_t)(nr * h->dim)));
  if (ns) NLS_CUDA(h, cudaMemcpyAsync(h->d_send_nodes, send_nodes, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, h->stream));
  if (nr) NLS_CUDA(h, cudaMemcpyAsync(h->d_recv_nodes, recv_nodes, sizeof(int32_t) * nr, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return nls_p2p_setup(h);      // collective: every rank of the communicator sets its halo
}

extern "C" int32_t nls_comm_unique_id(const char *nccl_lib_path, uint8_t *id128) {
  if (!id128) { g_noctx_error = "nls_comm_unique_id: null id"; return NLS_ERR_ARG; }
  std::string e = g_nccl.load(nccl_lib_path);
  if (!e.empty()) { g_noctx_error = e; return NLS_ERR_NCCL; }
  NcclDynId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) { g_noctx_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc); return NLS_ERR_NCCL; }
  std::memcpy(id128, id.internal, NCCL_DYN_ID_BYTES);
  return NLS_OK;
}

extern "C" int32_t nls_comm_init(nls_handle h, const char *nccl_lib_path, int32_t rank, int32_t nranks, const uint8_t *id128) {
  NLS_ENTER(h);
  if (!id128 || rank < 0 || rank >= nranks) NLS_FAIL(h, NLS_ERR_ARG, "nls_comm_init: bad arguments");
  std::string e = g_nccl.load(nccl_lib_path);
  if (!e.empty()) NLS_FAIL(h, NLS_ERR_NCCL, e);
  h->nccl = &g_nccl;
  NcclDynId id;
  std::memcpy(id.internal, id128, NCCL_DYN_ID_BYTES);
  void *comm = nullptr;
  int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (rc != 0) NLS_FAIL(h, NLS_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
  h->comm = comm; h->rank = rank; h->nranks = nranks; h->distributed = true;
  return NLS_OK;
}

// Enqueues one reduction over all ranks on the library stream (no host synchronisation): a device-side barrier that
// lines the ranks' streams up, e.g. right before a timed region.
extern "C" int32_t nls_dev_barrier(nls_handle h) {
  NLS_ENTER(h);
  if (!h->distributed) return NLS_OK;
  NLS_CUDA(h, cudaMemsetAsync(h->d_scal + 4, 0, 2 * sizeof(double), h->stream));
  return nls_allreduce_sum(h, h->d_scal + 4, 1);
}

extern "C" int32_t nls_comm_path(nls_handle h) {
  if (!h || !h->distributed) return 0;
  return h->p2p.on ? 2 : 1;
}

extern "C" int32_t nls_dev_comm_probe(nls_handle h, int32_t iters, float *us_halo, float *us_allreduce) {
  NLS_ENTER(h);
  if (!h->have_mesh || iters < 1) NLS_FAIL(h, NLS_ERR_ARG, "nls_dev_comm_probe: no mesh or iters < 1");
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {          // first round warms the connections up
    NLS_CUDA(h, cudaEventRecord(h->ev2, h->stream));
    for (int i = 0; i < iters; ++i) NLS_TRY(nls_halo_exchange(h, h->d_U));
    NLS_CUDA(h, cudaEventRecord(h->ev3, h->stream));
    NLS_CUDA(h, cudaEventSynchronize(h->ev3));
    ms = elapsed(h->ev2, h->ev3);
  }
  if (us_halo) *us_halo = ms * 1e3f / iters;
  NLS_CUDA(h, cudaMemsetAsync(h->d_scal, 0, 4 * sizeof(double), h->stream));
  for (int rep = 0; rep < 2; ++rep) {
    NLS_CUDA(h, cudaEventRecord(h->ev2, h->stream));
    for (int i = 0; i < iters; ++i) NLS_TRY(nls_allreduce_sum(h, h->d_scal, 2));
    NLS_CUDA(h, cudaEventRecord(h->ev3, h->stream));
    NLS_CUDA(h, cudaEventSynchronize(h->ev3));
    ms = elapsed(h->ev2, h->ev3);
  }
  if (us_allreduce) *us_allreduce = ms * 1e3f / iters;
  return nls_p2p_check(h);
}

extern "C" int32_t nls_dev_halo_exchange_u(nls_handle h) {
  NLS_ENTER(h);
  NLS_TRY(nls_halo_exchange(h, h->d_U));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}
