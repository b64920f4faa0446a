// kernels_dynamics.cu -- consistent mass and the Newmark state update on the device.
//
// Replaces (paths relative to the reference source tree):
//   _mass_el_fn! / assemblemass!       src/gallery/mass.jl:2-17     -> k_mass_1d / k_mass_q1 (once per model) + k_scale_add_mass
//   _applymass_el_fn! / applymass!     src/gallery/mass.jl:19-37    -> k_apply_mass (node-level SpMV over the same pattern)
//   update_state!                      src/timesteppers/newmark.jl:114-117 -> k_newmark_state
// The mass couples equal components only (M_{ai,bk} = delta_ik m_ab), so it is stored at node level: one double
// per block of the node-level pattern (nnzb), built once when the density is set; applying it and adding it to
// the Jacobian are pure HBM-bound passes.
#include "nls_internal.h"

// 1-D: m_ab = sum_q ((((N_a N_b) A) rho) J) w_q, J = length(e)/2 -- the product order of mass.jl:4.
template <int M>
__global__ void k_mass_1d(int64_t nel, int64_t n_owned, const int32_t *conn, const double *coords, double rho, double area,
                          const __grid_constant__ Tab1D T, const int64_t *browptr, const uint16_t *pos, double *mass) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nel) return;
  int nd[M];
#pragma unroll
  for (int a = 0; a < M; ++a) nd[a] = conn[e * M + a];
  const double J = (coords[nd[M - 1]] - coords[nd[0]]) / 2.0;
#pragma unroll
  for (int a = 0; a < M; ++a) {
    if (nd[a] >= n_owned) continue;
    const int64_t base = browptr[nd[a]];
#pragma unroll
    for (int b = 0; b < M; ++b) {
      double s = 0.0;
      for (int q = 0; q < T.nq; ++q) {
        const double v = ((((T.N[q][a] * T.N[q][b]) * area) * rho) * J) * T.w[q];
        s = (q == 0) ? v : s + v;
      }
      atomicAdd(mass + base + pos[(e * M + a) * M + b], s);
    }
  }
}

// Q1 (builder-defined, same structure): m_ab = sum_q ((N_a N_b) rho) (w_q det(dX/dxi)).
template <int DIM>
__global__ void k_mass_q1(int64_t nel, int64_t n_owned, const int32_t *conn, const double *coords, double rho,
                          const __grid_constant__ TabQ1 T, const int64_t *browptr, const uint16_t *pos, double *mass) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nel) return;
  constexpr int NEN = 1 << DIM;
  int nd[NEN];
  double X[NEN][DIM], wd[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    nd[a] = conn[e * NEN + a];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[a][d] = coords[(int64_t)nd[a] * DIM + d];
  }
#pragma unroll
  for (int q = 0; q < NEN; ++q) {
    double J[DIM * DIM], Ji[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) J[i] = 0.0;
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int D = 0; D < DIM; ++D) J[d * DIM + D] += X[a][d] * T.dN[q][a][D];
    double det;
    if (DIM == 2) { det = J[0] * J[3] - J[1] * J[2]; }
    else {
      det = J[0] * (J[4] * J[8] - J[5] * J[7]) + J[1] * (J[5] * J[6] - J[3] * J[8]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
    }
    (void)Ji;
    wd[q] = T.w[q] * det;
  }
  for (int a = 0; a < NEN; ++a) {
    if (nd[a] >= n_owned) continue;
    const int64_t base = browptr[nd[a]];
    for (int b = 0; b < NEN; ++b) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < NEN; ++q) s += ((T.N[q][a] * T.N[q][b]) * rho) * wd[q];
      atomicAdd(mass + base + pos[(e * NEN + a) * NEN + b], s);
    }
  }
}

// y[a,i] (+)= sum_j m[a,j] x[bcol[j], i]: one thread per (owned node, component)
__global__ void __launch_bounds__(256)
k_apply_mass(int64_t n_owned, int dim, const int64_t *browptr, const int32_t *bcol, const double *mass, const double *x,
             double *y, int accumulate) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_owned * dim) return;
  const int64_t a = r / dim;
  const int i = (int)(r - a * dim);
  double s = 0.0;
  for (int64_t j = browptr[a]; j < browptr[a + 1]; ++j) s = fma(mass[j], x[(int64_t)bcol[j] * dim + i], s);
  y[r] = accumulate ? y[r] + s : s;
}

// vals <- scale * vals + M (jacobian.jl:46-50), or vals <- M when scale_existing == 0: one thread per node-level block
__global__ void __launch_bounds__(256)
k_scale_add_mass(int64_t n_owned, int dim, const int64_t *browptr, const double *mass, double scale, int scale_existing,
                 double *vals) {
  const int64_t a = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per node
  if (a >= n_owned) return;
  const int64_t b0 = browptr[a];
  const int deg = (int)(browptr[a + 1] - b0);
  double *row = vals + (int64_t)dim * dim * b0;
  const int n = dim * dim * deg;     // the node's rows are contiguous: [i][j][k]
  for (int t = threadIdx.x & 31; t < n; t += 32) {
    const int i = t / (dim * deg), rem = t - i * dim * deg, j = rem / dim, k = rem - j * dim;
    const double m = (i == k) ? mass[b0 + j] : 0.0;
    row[t] = scale_existing ? fma(scale, row[t], m) : m;
  }
}

// update_state! (newmark.jl:114-117): U = Ut + dt^2 beta Udd ; Ud = Udt + dt gamma Udd
__global__ void __launch_bounds__(256)
k_newmark_state(int64_t n, const double *Ut, const double *Udt, const double *Udd, double c_u, double c_v, double *U, double *Ud) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = Udd[i];
    U[i] = fma(c_u, a, Ut[i]);
    Ud[i] = fma(c_v, a, Udt[i]);
  }
}

#define DYN_KCHECK(h) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string("kernel launch: ") + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

int nls_launch_mass_build(nls_context *h, double rho, double area) {
  NLS_CUDA(h, cudaMemsetAsync(h->d_mass, 0, sizeof(double) * h->nnzb, h->stream));
  if (h->nel == 0) return NLS_OK;
  const unsigned g = (unsigned)((h->nel + 127) / 128);
  if (h->dim == 1) {
    switch (h->nen) {
#define MK(M) case M: k_mass_1d<M><<<g, 128, 0, h->stream>>>(h->nel, h->n_owned, h->d_conn, h->d_coords, rho, area, h->tab1, h->d_browptr, h->d_pos, h->d_mass); break;
      MK(2) MK(3) MK(4) MK(5) MK(6) MK(7) MK(8)
#undef MK
      default: NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "1-D elements support p = 1..7");
    }
  } else if (h->dim == 2) {
    k_mass_q1<2><<<g, 128, 0, h->stream>>>(h->nel, h->n_owned, h->d_conn, h->d_coords, rho, h->tabq, h->d_browptr, h->d_pos, h->d_mass);
  } else {
    k_mass_q1<3><<<g, 128, 0, h->stream>>>(h->nel, h->n_owned, h->d_conn, h->d_coords, rho, h->tabq, h->d_browptr, h->d_pos, h->d_mass);
  }
  DYN_KCHECK(h);
  return NLS_OK;
}

int nls_launch_apply_mass(nls_context *h, const double *x, double *y, bool accumulate) {
  if (h->nrows == 0) return NLS_OK;
  k_apply_mass<<<(unsigned)((h->nrows + 255) / 256), 256, 0, h->stream>>>(h->n_owned, h->dim, h->d_browptr, h->d_bcol, h->d_mass, x, y,
                                                                             accumulate ? 1 : 0);
  DYN_KCHECK(h);
  return NLS_OK;
}

int nls_launch_scale_add_mass(nls_context *h, double scale, bool scale_existing) {
  if (h->n_owned == 0) return NLS_OK;
  k_scale_add_mass<<<(unsigned)((h->n_owned + 7) / 8), 256, 0, h->stream>>>(h->n_owned, h->dim, h->d_browptr, h->d_mass, scale,
                                                                             scale_existing ? 1 : 0, h->d_vals);
  DYN_KCHECK(h);
  h->vals_version++;
  return NLS_OK;
}

int nls_launch_newmark_state(nls_context *h, double c_u, double c_v) {
  const int64_t n = h->ndof_local();
  const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8));
  k_newmark_state<<<g, 256, 0, h->stream>>>(n, h->d_Ut, h->d_Udt, h->d_Udd, c_u, c_v, h->d_U, h->d_Ud);
  DYN_KCHECK(h);
  return NLS_OK;
}
