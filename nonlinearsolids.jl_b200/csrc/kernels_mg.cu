// kernels_mg.cu -- geometric multigrid preconditioner of the Krylov solve on 3-D meshes with grid topology.
//
// Replaces, together with kernels_krylov.cu:  dofs(K) \ dofs(-R)  (src/nonlinearsolvers/newton_fem.jl:158). The reference
// factorises K; at 10^7 DoF the B200 path iterates, and Jacobi-PCG needs O(n) iterations per solve (1 600 at C3). Here the
// preconditioner is one V-cycle on a hierarchy of Galerkin operators K_l+1 = P^T K_l P:
//   * levels: the node grid coarsened by two per axis (coarse nodes = even fine nodes, plus the last node of an axis with an
//     even node count), P = trilinear interpolation (weights 1, 1/2 per axis), the same for the three displacement components;
//   * level 0 is the assembled CSR Jacobian itself (block positions from the line tables of kernels_lines3d.cu, Dirichlet
//     rows and columns masked out = the free block dofs(K)); coarser operators are stored as 27 dense 3 x 3 blocks per node;
//   * smoother: Chebyshev polynomial in D^-1 K (D = the 3 x 3 node blocks), largest eigenvalue by a few power iterations
//     whenever the values change; the coarsest grid (<= 6 nodes per axis) gets a long Chebyshev run instead of a factorisation;
//   * the Galerkin products are gathers: one thread per (coarse node, stencil slot) sums its <= 343 weighted fine blocks in a
//     fixed order: deterministic, no atomics.
// Multi-GPU: every rank runs the V-cycle on its own slab with the couplings to ghost nodes dropped (a non-overlapping
// Schwarz layer on top of the multigrid); the Krylov loop around it exchanges halos and reduces as before.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "nls_internal.h"

#ifndef NLS_TRY
#define NLS_TRY(call) do { int rc_ = (call); if (rc_ != NLS_OK) return rc_; } while (0)
#endif
#define MG_KCHECK(h) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string("multigrid kernel launch: ") + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

struct MgL3Line { int64_t lb; int32_t nl; int8_t rank[9]; int8_t pad[3]; };   // = L3Line of kernels_lines3d.cu

struct MgFine {          // level 0: the CSR Jacobian of a grid-topology mesh, owned planes [z0, z0 + nz)
  int nx, ny, nz, z0, nzp;
  const int32_t *pbase;
  const uint8_t *pown;
  const MgL3Line *info;
  const double *vals;
  const uint8_t *con;
};

__host__ __device__ __forceinline__ int mg_coarse_n(int n) { return n <= 2 ? n : ((n & 1) ? (n + 1) / 2 : n / 2 + 1); }
// fine index of coarse node X, and the interpolation weight of fine node x on it (0 if none)
__device__ __forceinline__ int mg_fx(int X, int n) { return n <= 2 ? X : min(2 * X, n - 1); }
__device__ __forceinline__ double mg_w(int x, int X, int n) {
  const int f = mg_fx(X, n);
  if (x == f) return 1.0;
  if (n <= 2) return 0.0;
  const bool coarse = !(x & 1) || x == n - 1;
  return (!coarse && (x - f == 1 || f - x == 1)) ? 0.5 : 0.0;
}

__device__ __forceinline__ bool mg_inv3(const double *M, double *Mi) {
  const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  const double sc = fabs(M[0] * M[4] * M[8]);
  if (!(fabs(det) > 1e-14 * sc) || !(sc > 0.0)) { for (int k = 0; k < 9; ++k) Mi[k] = 0.0; return false; }
  const double id = 1.0 / det;
  Mi[0] = c00 * id; Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id; Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  Mi[3] = c01 * id; Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id; Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  Mi[6] = c02 * id; Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id; Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  return true;
}

// block (i, j) of the free block of the fine CSR matrix, i = (x, y, z), j = i + (dx, dy, dz); z counts owned planes.
// Returns false (B untouched) when j does not exist or is a ghost node.
__device__ __forceinline__ bool mg_fine_block(const MgFine &F, int x, int y, int z, int dx, int dy, int dz, double (&B)[9]) {
  const int x2 = x + dx, y2 = y + dy, z2 = z + dz;
  if (x2 < 0 || x2 >= F.nx || y2 < 0 || y2 >= F.ny || z2 < 0 || z2 >= F.nz) return false;
  const int zp = F.z0 + z;
  const MgL3Line *li = F.info + ((int64_t)zp * F.ny + y);
  const int rank = li->rank[(dz + 1) * 3 + (dy + 1)];
  if (rank < 0) return false;
  const int nl = li->nl, w = (x == 0 || x == F.nx - 1) ? 2 : 3;
  const int64_t b0 = li->lb + (int64_t)nl * (x == 0 ? 0 : 3 * x - 1);
  const int pos = rank * w + dx + (x == 0 ? 0 : 1);
  const double *p = F.vals + 9 * b0 + 3 * pos;
  const int64_t stride = 3 * (int64_t)nl * w;
  const int64_t ni = (int64_t)F.pbase[zp] + (int64_t)y * F.nx + x, nj = (int64_t)F.pbase[zp + dz] + (int64_t)y2 * F.nx + x2;
  const uint8_t *ci = F.con + 3 * ni, *cj = F.con + 3 * nj;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) B[r * 3 + c] = (ci[r] | cj[c]) ? 0.0 : p[r * stride + c];
  return true;
}

// ---- level 0: inverse of the masked 3 x 3 node blocks (block-Jacobi); constrained DoFs get zero rows and columns ----
__global__ void __launch_bounds__(256) k_mg_fine_dinv(const MgFine F, double *__restrict__ dinv) {
  const int64_t nn = (int64_t)F.nx * F.ny * F.nz;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nn; t += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(t % F.nx), y = (int)((t / F.nx) % F.ny), z = (int)(t / ((int64_t)F.nx * F.ny));
    double B[9], Bi[9];
    mg_fine_block(F, x, y, z, 0, 0, 0, B);
    const int64_t ni = (int64_t)F.pbase[F.z0 + z] + (int64_t)y * F.nx + x;
    const uint8_t *ci = F.con + 3 * ni;
    for (int r = 0; r < 3; ++r) if (ci[r]) B[r * 3 + r] = 1.0;
    mg_inv3(B, Bi);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) dinv[9 * ni + r * 3 + c] = (ci[r] | ci[c]) ? 0.0 : Bi[r * 3 + c];
  }
}

// ---- Galerkin product, one thread per (coarse node, stencil slot) ----
// FINE: the fine operator is the CSR matrix (level 0), else a 27-slot stencil matrix on a (fnx, fny, fnz) grid
template <bool FINE>
__global__ void __launch_bounds__(128) k_mg_rap(const MgFine F, const double *__restrict__ Kf, int fnx, int fny, int fnz,
                                                 int cnx, int cny, int cnz, double *__restrict__ Kc) {
  const int64_t total = (int64_t)cnx * cny * cnz * 27;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(t % 27);
    const int64_t I = t / 27;
    const int X = (int)(I % cnx), Y = (int)((I / cnx) % cny), Z = (int)(I / ((int64_t)cnx * cny));
    const int X2 = X + s % 3 - 1, Y2 = Y + (s / 3) % 3 - 1, Z2 = Z + s / 9 - 1;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (X2 >= 0 && X2 < cnx && Y2 >= 0 && Y2 < cny && Z2 >= 0 && Z2 < cnz) {
      const int fxI = mg_fx(X, fnx), fyI = mg_fx(Y, fny), fzI = mg_fx(Z, fnz);
      const int fxJ = mg_fx(X2, fnx), fyJ = mg_fx(Y2, fny), fzJ = mg_fx(Z2, fnz);
      for (int z = max(fzI - 1, 0); z <= min(fzI + 1, fnz - 1); ++z) {
        const double wz = mg_w(z, Z, fnz);
        if (wz == 0.0) continue;
        for (int z2 = max(max(fzJ - 1, 0), z - 1); z2 <= min(min(fzJ + 1, fnz - 1), z + 1); ++z2) {
          const double wz2 = wz * mg_w(z2, Z2, fnz);
          if (wz2 == 0.0) continue;
          for (int y = max(fyI - 1, 0); y <= min(fyI + 1, fny - 1); ++y) {
            const double wy = wz2 * mg_w(y, Y, fny);
            if (wy == 0.0) continue;
            for (int y2 = max(max(fyJ - 1, 0), y - 1); y2 <= min(min(fyJ + 1, fny - 1), y + 1); ++y2) {
              const double wy2 = wy * mg_w(y2, Y2, fny);
              if (wy2 == 0.0) continue;
              for (int x = max(fxI - 1, 0); x <= min(fxI + 1, fnx - 1); ++x) {
                const double wx = wy2 * mg_w(x, X, fnx);
                if (wx == 0.0) continue;
                for (int x2 = max(max(fxJ - 1, 0), x - 1); x2 <= min(min(fxJ + 1, fnx - 1), x + 1); ++x2) {
                  const double w = wx * mg_w(x2, X2, fnx);
                  if (w == 0.0) continue;
                  if (FINE) {
                    double B[9];
                    if (!mg_fine_block(F, x, y, z, x2 - x, y2 - y, z2 - z, B)) continue;
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc[k] = fma(w, B[k], acc[k]);
                  } else {
                    const int sf = (z2 - z + 1) * 9 + (y2 - y + 1) * 3 + (x2 - x + 1);
                    const double *B = Kf + ((((int64_t)z * fny + y) * fnx + x) * 27 + sf) * 9;
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc[k] = fma(w, B[k], acc[k]);
                  }
                }
              }
            }
          }
        }
      }
    }
    double *o = Kc + t * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = acc[k];
  }
}

// The same product for level 1 (fine operator = the CSR Jacobian, 3/4 of the whole set-up), one WARP per coarse node: lane s
// owns stencil slot s. For every fine node i under the coarse node (<= 27) the 27 lanes first fetch the 27 blocks of i's CSR
// block row -- one contiguous 1.9 KB row, each block once -- into shared memory; then every lane adds the <= 8 of them whose
// column node interpolates from its coarse neighbour. Same terms and the same order as k_mg_rap<true>.
__global__ void __launch_bounds__(128) k_mg_rap_fine_warp(const MgFine F, int cnx, int cny, int cnz, double *__restrict__ Kc) {
  __shared__ double sB[4][27][9];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int fnx = F.nx, fny = F.ny, fnz = F.nz;
  const int64_t ncoarse = (int64_t)cnx * cny * cnz;
  const int dxs = lane % 3 - 1, dys = (lane / 3) % 3 - 1, dzs = lane / 9 - 1;      // lane < 27: both the block it fetches and its slot
  for (int64_t I = blockIdx.x * 4ll + wib; I < ncoarse; I += (int64_t)gridDim.x * 4) {
    const int X = (int)(I % cnx), Y = (int)((I / cnx) % cny), Z = (int)(I / ((int64_t)cnx * cny));
    const int X2 = X + dxs, Y2 = Y + dys, Z2 = Z + dzs;
    const bool jok = lane < 27 && X2 >= 0 && X2 < cnx && Y2 >= 0 && Y2 < cny && Z2 >= 0 && Z2 < cnz;
    const int fxI = mg_fx(X, fnx), fyI = mg_fx(Y, fny), fzI = mg_fx(Z, fnz);
    const int fxJ = jok ? mg_fx(X2, fnx) : 0, fyJ = jok ? mg_fx(Y2, fny) : 0, fzJ = jok ? mg_fx(Z2, fnz) : 0;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int z = max(fzI - 1, 0); z <= min(fzI + 1, fnz - 1); ++z) {
      const double wz = mg_w(z, Z, fnz);
      if (wz == 0.0) continue;
      for (int y = max(fyI - 1, 0); y <= min(fyI + 1, fny - 1); ++y) {
        const double wy = wz * mg_w(y, Y, fny);
        if (wy == 0.0) continue;
        for (int x = max(fxI - 1, 0); x <= min(fxI + 1, fnx - 1); ++x) {
          const double wi = wy * mg_w(x, X, fnx);
          if (wi == 0.0) continue;
          __syncwarp();
          if (lane < 27) {
            double B[9];
            if (!mg_fine_block(F, x, y, z, dxs, dys, dzs, B)) {
#pragma unroll
              for (int k = 0; k < 9; ++k) B[k] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) sB[wib][lane][k] = B[k];
          }
          __syncwarp();
          if (jok) {
            for (int z2 = max(max(fzJ - 1, 0), z - 1); z2 <= min(min(fzJ + 1, fnz - 1), z + 1); ++z2) {
              const double wz2 = wi * mg_w(z2, Z2, fnz);
              if (wz2 == 0.0) continue;
              for (int y2 = max(max(fyJ - 1, 0), y - 1); y2 <= min(min(fyJ + 1, fny - 1), y + 1); ++y2) {
                const double wy2 = wz2 * mg_w(y2, Y2, fny);
                if (wy2 == 0.0) continue;
                for (int x2 = max(max(fxJ - 1, 0), x - 1); x2 <= min(min(fxJ + 1, fnx - 1), x + 1); ++x2) {
                  const double w = wy2 * mg_w(x2, X2, fnx);
                  if (w == 0.0) continue;
                  const double *B = sB[wib][(z2 - z + 1) * 9 + (y2 - y + 1) * 3 + (x2 - x + 1)];
#pragma unroll
                  for (int k = 0; k < 9; ++k) acc[k] = fma(w, B[k], acc[k]);
                }
              }
            }
          }
        }
      }
    }
    if (lane < 27) {
      double *o = Kc + (I * 27 + lane) * 9;
#pragma unroll
      for (int k = 0; k < 9; ++k) o[k] = acc[k];
    }
  }
}

// coarse levels: inverse of the diagonal blocks (slot 13); a singular block (a coarse node whose fine nodes are all
// constrained) gets zero
__global__ void __launch_bounds__(256) k_mg_dinv(int64_t nn, const double *__restrict__ K, double *__restrict__ dinv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double Bi[9];
    mg_inv3(K + (i * 27 + 13) * 9, Bi);
    for (int k = 0; k < 9; ++k) dinv[9 * i + k] = Bi[k];
  }
}

// coarse levels: y = K x, three lanes per node (one per scalar row), 27 blocks each
__global__ void __launch_bounds__(256) k_mg_spmv(int nx, int ny, int nz, const double *__restrict__ K, const double *__restrict__ x,
                                                  double *__restrict__ y) {
  const int64_t nrows = (int64_t)nx * ny * nz * 3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / 3;
    const int r = (int)(t - 3 * i);
    const int X = (int)(i % nx), Y = (int)((i / nx) % ny), Z = (int)(i / ((int64_t)nx * ny));
    const double *Ki = K + i * 243 + r * 3;
    double s = 0.0;
    for (int dz = -1; dz <= 1; ++dz) {
      if (Z + dz < 0 || Z + dz >= nz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        if (Y + dy < 0 || Y + dy >= ny) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          if (X + dx < 0 || X + dx >= nx) continue;
          const int sl = (dz + 1) * 9 + (dy + 1) * 3 + dx + 1;
          const int64_t j = i + ((int64_t)dz * ny + dy) * nx + dx;
          const double *b = Ki + sl * 9, *xj = x + 3 * j;
          s = fma(b[0], xj[0], fma(b[1], xj[1], fma(b[2], xj[2], s)));
        }
      }
    }
    y[t] = s;
  }
}

// one Chebyshev step on nodes [0, nn): d = a d + c D^-1 (b - y), x += d  (y = K x of the current x, or absent while x = 0).
// Vectors are indexed by node id; `dinv` too.
__global__ void __launch_bounds__(256) k_mg_cheb(int64_t nn, const double *__restrict__ dinv, const double *__restrict__ b,
                                                  const double *__restrict__ y, double *__restrict__ d, double *__restrict__ x,
                                                  double a, double c, int mode) {   // mode 1: x = 0 on entry (x = d), 2: new run (d restarts), 0: inside a run
  const bool first = mode == 1, restart = mode != 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double r0 = b[3 * i], r1 = b[3 * i + 1], r2 = b[3 * i + 2];
    if (y) { r0 -= y[3 * i]; r1 -= y[3 * i + 1]; r2 -= y[3 * i + 2]; }
    const double *D = dinv + 9 * i;
    const double z0 = D[0] * r0 + D[1] * r1 + D[2] * r2, z1 = D[3] * r0 + D[4] * r1 + D[5] * r2, z2 = D[6] * r0 + D[7] * r1 + D[8] * r2;
    double d0 = c * z0, d1 = c * z1, d2 = c * z2;
    if (!restart) { d0 = fma(a, d[3 * i], d0); d1 = fma(a, d[3 * i + 1], d1); d2 = fma(a, d[3 * i + 2], d2); }
    d[3 * i] = d0; d[3 * i + 1] = d1; d[3 * i + 2] = d2;
    if (first) { x[3 * i] = d0; x[3 * i + 1] = d1; x[3 * i + 2] = d2; }
    else { x[3 * i] += d0; x[3 * i + 1] += d1; x[3 * i + 2] += d2; }
  }
}

// z = D^-1 r on the owned nodes (block-Jacobi alone, and the power iteration)
__global__ void __launch_bounds__(256) k_mg_apply_dinv(int64_t nn, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    const double r0 = r[3 * i], r1 = r[3 * i + 1], r2 = r[3 * i + 2];
    const double *D = dinv + 9 * i;
    z[3 * i] = D[0] * r0 + D[1] * r1 + D[2] * r2; z[3 * i + 1] = D[3] * r0 + D[4] * r1 + D[5] * r2; z[3 * i + 2] = D[6] * r0 + D[7] * r1 + D[8] * r2;
  }
}

// restriction bc = P^T (b - y): one thread per coarse DoF gathers its <= 27 fine nodes. `pbase` maps the fine grid to node
// ids on level 0 (owned planes z0 ..); null = lexicographic.
__global__ void __launch_bounds__(256) k_mg_restrict(int fnx, int fny, int fnz, int cnx, int cny, int cnz, const int32_t *__restrict__ pbase, int z0,
                                                      const double *__restrict__ b, const double *__restrict__ y, double *__restrict__ bc) {
  const int64_t total = (int64_t)cnx * cny * cnz * 3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = t / 3;
    const int r = (int)(t - 3 * I);
    const int X = (int)(I % cnx), Y = (int)((I / cnx) % cny), Z = (int)(I / ((int64_t)cnx * cny));
    const int fx = mg_fx(X, fnx), fy = mg_fx(Y, fny), fz = mg_fx(Z, fnz);
    double s = 0.0;
    for (int z = max(fz - 1, 0); z <= min(fz + 1, fnz - 1); ++z) {
      const double wz = mg_w(z, Z, fnz);
      if (wz == 0.0) continue;
      for (int yy = max(fy - 1, 0); yy <= min(fy + 1, fny - 1); ++yy) {
        const double wy = wz * mg_w(yy, Y, fny);
        if (wy == 0.0) continue;
        for (int x = max(fx - 1, 0); x <= min(fx + 1, fnx - 1); ++x) {
          const double w = wy * mg_w(x, X, fnx);
          if (w == 0.0) continue;
          const int64_t i = pbase ? (int64_t)pbase[z0 + z] + (int64_t)yy * fnx + x : ((int64_t)z * fny + yy) * fnx + x;
          s = fma(w, b[3 * i + r] - y[3 * i + r], s);
        }
      }
    }
    bc[t] = s;
  }
}

// prolongation x += P xc: one thread per fine DoF gathers its <= 8 coarse parents; constrained DoFs (level 0) stay zero
__global__ void __launch_bounds__(256) k_mg_prolong(int fnx, int fny, int fnz, int cnx, int cny, int cnz, const int32_t *__restrict__ pbase, int z0,
                                                     const uint8_t *__restrict__ con, const double *__restrict__ xc, double *__restrict__ x) {
  const int64_t total = (int64_t)fnx * fny * fnz * 3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t il = t / 3;
    const int r = (int)(t - 3 * il);
    const int xx = (int)(il % fnx), yy = (int)((il / fnx) % fny), zz = (int)(il / ((int64_t)fnx * fny));
    const int64_t i = pbase ? (int64_t)pbase[z0 + zz] + (int64_t)yy * fnx + xx : il;
    if (con && con[3 * i + r]) continue;
    // parents per axis: the coarse node at x (weight 1), or the two at x - 1 and x + 1 (weight 1/2 each)
    int px[2], py[2], pz[2], npx, npy, npz;
    auto parents = [](int v, int n, int *p) -> int {
      if (n <= 2) { p[0] = v; return 1; }
      if (!(v & 1)) { p[0] = v >> 1; return 1; }
      if (v == n - 1) { p[0] = mg_coarse_n(n) - 1; return 1; }
      p[0] = (v - 1) >> 1; p[1] = (v + 1 == n - 1) ? mg_coarse_n(n) - 1 : (v + 1) >> 1;
      return 2;
    };
    npx = parents(xx, fnx, px); npy = parents(yy, fny, py); npz = parents(zz, fnz, pz);
    const double w = 1.0 / (double)(npx * npy * npz);
    double s = 0.0;
    for (int c = 0; c < npz; ++c)
      for (int bq = 0; bq < npy; ++bq)
        for (int a = 0; a < npx; ++a) s += xc[3 * ((((int64_t)pz[c] * cny) + py[bq]) * cnx + px[a]) + r];
    x[3 * i + r] += w * s;
  }
}

__global__ void __launch_bounds__(256) k_mg_xpby(int64_t n, const double *__restrict__ z, double beta, double *__restrict__ p) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}
__global__ void __launch_bounds__(256) k_mg_mask_copy(int64_t n, const double *__restrict__ b, const uint8_t *__restrict__ con, double *__restrict__ r) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] = con[i] ? 0.0 : b[i];
}
__global__ void __launch_bounds__(256) k_mg_fill_hash(int64_t n, unsigned seed, double *__restrict__ v) {   // fixed pseudo-random start vector
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long s = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
    v[i] = (double)(s & 0xffff) / 65536.0 - 0.5;
  }
}
__global__ void __launch_bounds__(256) k_mg_scale(int64_t n, double a, double *__restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] *= a;
}

// ---- host side -----------------------------------------------------------------------------------------------------------
static unsigned mg_grid(nls_context *h, int64_t n, int threads = 256) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)h->sm_count * 16;
  return (unsigned)std::max<int64_t>(1, std::min(g, cap));
}

void nls_mg_free(nls_context *h) {
  MgPlan &m = h->mg;
  for (int l = 0; l < MG_MAX_LEVELS; ++l) {
    MgLevel &L = m.lev[l];
    if (L.K) cudaFree(L.K);
    if (L.dinv) cudaFree(L.dinv);
    for (double *v : {L.x, L.b, L.y, L.d}) if (v) cudaFree(v);
  }
  if (m.vals32) cudaFree(m.vals32);
  m = MgPlan();
}

// the hierarchy of a grid-topology mesh (called when the line plan is built or on first use); no values yet
int nls_mg_build(nls_context *h) {
  nls_mg_free(h);
  MgPlan &m = h->mg;
  const Lines3Plan &p = h->lplan;
  if (!p.ok || h->dim != 3 || p.h_pown.empty()) return NLS_OK;
  int z0 = -1, z1 = -1;
  for (int z = 0; z < p.nzp; ++z) if (p.h_pown[z]) { if (z0 < 0) z0 = z; z1 = z + 1; }
  if (z0 < 0) return NLS_OK;
  for (int z = z0; z < z1; ++z) if (!p.h_pown[z]) return NLS_OK;       // owned planes must be contiguous
  if ((int64_t)p.nx * p.ny * (z1 - z0) != h->n_owned) return NLS_OK;
  m.z0 = z0;
  int nx = p.nx, ny = p.ny, nz = z1 - z0, l = 0;
  for (;;) {
    MgLevel &L = m.lev[l];
    L.nx = nx; L.ny = ny; L.nz = nz; L.nn = (int64_t)nx * ny * nz;
    const size_t nd = (size_t)(l == 0 ? h->ndof_local() : 3 * L.nn);
    if (l > 0) NLS_CUDA(h, cudaMalloc((void **)&L.K, sizeof(double) * 243 * (size_t)L.nn));
    NLS_CUDA(h, cudaMalloc((void **)&L.dinv, sizeof(double) * 9 * (size_t)(l == 0 ? h->nn : L.nn)));
    if (l == 0) NLS_CUDA(h, cudaMemsetAsync(L.dinv, 0, sizeof(double) * 9 * (size_t)h->nn, h->stream));
    for (double **v : {&L.x, &L.b, &L.y, &L.d}) {
      if (l == 0 && (v == &L.b)) continue;                  // level 0: b is the caller's residual
      NLS_CUDA(h, cudaMalloc((void **)v, sizeof(double) * nd));
      NLS_CUDA(h, cudaMemsetAsync(*v, 0, sizeof(double) * nd, h->stream));   // ghost entries stay zero for ever
    }
    ++l;
    const int cx = mg_coarse_n(nx), cy = mg_coarse_n(ny), cz = mg_coarse_n(nz);
    if (l >= MG_MAX_LEVELS || std::max(nx, std::max(ny, nz)) <= 6 || (cx == nx && cy == ny && cz == nz)) break;
    nx = cx; ny = cy; nz = cz;
  }
  m.nlev = l;
  m.ok = m.nlev >= 2;
  m.version = -1;
  return NLS_OK;
}

static MgFine mg_fine_args(nls_context *h) {
  const Lines3Plan &p = h->lplan;
  const MgLevel &L = h->mg.lev[0];
  MgFine F;
  F.nx = L.nx; F.ny = L.ny; F.nz = L.nz; F.z0 = h->mg.z0; F.nzp = p.nzp;
  F.pbase = p.d_pbase; F.pown = p.d_pown; F.info = reinterpret_cast<const MgL3Line *>(p.d_info);
  F.vals = h->d_vals; F.con = h->d_con;
  return F;
}

static int mg_apply_K(nls_context *h, int l, const double *x, double *y) {
  MgLevel &L = h->mg.lev[l];
  if (l == 0) return (h->mg.fp32 && h->mg.vals32) ? nls_launch_spmv_f32(h, h->mg.vals32, x, y) : nls_launch_spmv(h, x, y, true, false);
  k_mg_spmv<<<mg_grid(h, 3 * L.nn), 256, 0, h->stream>>>(L.nx, L.ny, L.nz, L.K, x, y);
  MG_KCHECK(h);
  return NLS_OK;
}

// largest eigenvalue of D^-1 K on level l by power iteration (start vector: a fixed pseudo-random pattern via the smoother
// vectors; the estimate is inflated by 10 %, the usual safeguard of a Chebyshev smoother)
static int mg_lambda_max(nls_context *h, int l, double *out) {
  MgLevel &L = h->mg.lev[l];
  const int64_t nn = l == 0 ? h->n_owned : L.nn, n = 3 * nn;
  k_mg_fill_hash<<<mg_grid(h, n), 256, 0, h->stream>>>(n, (unsigned)l, L.x); MG_KCHECK(h);
  if (l == 0) { k_mg_mask_copy<<<mg_grid(h, n), 256, 0, h->stream>>>(n, L.x, h->d_con, L.x); MG_KCHECK(h); }
  double lam = 1.0;
  for (int it = 0; it < h->mg.power_its; ++it) {
    double nrm2;
    { // local norm (the preconditioner is rank-local: no allreduce)
      NLS_TRY(nls_dev_dot_local(h, n, L.x, L.x, &nrm2));
    }
    if (!(nrm2 > 0.0)) { lam = 1.0; break; }
    k_mg_scale<<<mg_grid(h, n), 256, 0, h->stream>>>(n, 1.0 / sqrt(nrm2), L.x); MG_KCHECK(h);
    NLS_TRY(mg_apply_K(h, l, L.x, L.y));
    k_mg_apply_dinv<<<mg_grid(h, nn), 256, 0, h->stream>>>(nn, L.dinv, L.y, L.d); MG_KCHECK(h);
    double xd;
    NLS_TRY(nls_dev_dot_local(h, n, L.x, L.d, &xd));
    lam = xd;                                    // Rayleigh quotient of the normalised iterate
    NLS_CUDA(h, cudaMemcpyAsync(L.x, L.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
  }
  NLS_CUDA(h, cudaMemsetAsync(L.x, 0, sizeof(double) * n, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(L.d, 0, sizeof(double) * n, h->stream));
  *out = 1.1 * fabs(lam);
  return NLS_OK;
}

// Galerkin operators, block inverses and eigenvalue bounds for the current CSR values
int nls_mg_setup(nls_context *h) {
  MgPlan &m = h->mg;
  if (!m.ok) NLS_FAIL(h, NLS_ERR_STATE, "multigrid: no hierarchy for this mesh");
  if (m.version == h->vals_version) return NLS_OK;
  const MgFine F = mg_fine_args(h);
  k_mg_fine_dinv<<<mg_grid(h, m.lev[0].nn), 256, 0, h->stream>>>(F, m.lev[0].dinv); MG_KCHECK(h);
  for (int l = 1; l < m.nlev; ++l) {
    MgLevel &Lf = m.lev[l - 1], &Lc = m.lev[l];
    if (l == 1 && m.rap_warp) k_mg_rap_fine_warp<<<mg_grid(h, Lc.nn * 32, 128), 128, 0, h->stream>>>(F, Lc.nx, Lc.ny, Lc.nz, Lc.K);
    else if (l == 1) k_mg_rap<true><<<mg_grid(h, Lc.nn * 27, 128), 128, 0, h->stream>>>(F, nullptr, Lf.nx, Lf.ny, Lf.nz, Lc.nx, Lc.ny, Lc.nz, Lc.K);
    else k_mg_rap<false><<<mg_grid(h, Lc.nn * 27, 128), 128, 0, h->stream>>>(F, Lf.K, Lf.nx, Lf.ny, Lf.nz, Lc.nx, Lc.ny, Lc.nz, Lc.K);
    MG_KCHECK(h);
    k_mg_dinv<<<mg_grid(h, Lc.nn), 256, 0, h->stream>>>(Lc.nn, Lc.K, Lc.dinv); MG_KCHECK(h);
  }
  if (m.fp32) {
    if (!m.vals32 && cudaMalloc((void **)&m.vals32, sizeof(float) * (size_t)h->nnz) != cudaSuccess) { cudaGetLastError(); m.vals32 = nullptr; }
    if (m.vals32) NLS_TRY(nls_launch_vals_to_f32(h, m.vals32));
  }
  // eigenvalue bounds: the Jacobian of consecutive Newton iterations changes by a few per cent at most, well inside the
  // 10 % safety margin: a full power iteration every eighth set of values, in between the bounds are kept
  if (m.power_version < 0 || h->vals_version - m.power_version >= 8 || h->vals_version < m.power_version) {
    for (int l = 0; l < m.nlev; ++l) NLS_TRY(mg_lambda_max(h, l, &m.lev[l].lmax));
    m.power_version = h->vals_version;
  }
  m.version = h->vals_version;
  return NLS_OK;
}

// Chebyshev smoothing of K_l x = b on [lmax / ratio, lmax]; zero = x starts from zero
static int mg_cheb(nls_context *h, int l, const double *b, double *x, int degree, double ratio, bool zero) {
  MgLevel &L = h->mg.lev[l];
  const int64_t nn = l == 0 ? h->n_owned : L.nn;
  const double lmax = L.lmax, lmin = lmax / ratio, theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  for (int k = 0; k < degree; ++k) {
    const bool first = zero && k == 0;
    if (!first) NLS_TRY(mg_apply_K(h, l, x, L.y));
    double a, c;
    if (k == 0) { a = 0.0; c = 1.0 / theta; }
    else { const double rn = 1.0 / (2.0 * sigma - rho); a = rn * rho; c = 2.0 * rn / delta; rho = rn; }
    // the first step of a run always restarts the direction d (a = 0); `first` also means x = 0 (no product, x = d)
    k_mg_cheb<<<mg_grid(h, nn), 256, 0, h->stream>>>(nn, L.dinv, b, first ? nullptr : L.y, L.d, x, a, c, first ? 1 : (k == 0 ? 2 : 0));
    MG_KCHECK(h);
  }
  return NLS_OK;
}

static int mg_vcycle(nls_context *h, int l, const double *b, double *x) {
  MgPlan &m = h->mg;
  MgLevel &L = m.lev[l];
  if (l == m.nlev - 1) return mg_cheb(h, l, b, x, m.coarse_degree, m.coarse_ratio, true);
  NLS_TRY(mg_cheb(h, l, b, x, m.degree, m.ratio, true));
  NLS_TRY(mg_apply_K(h, l, x, L.y));
  MgLevel &C = m.lev[l + 1];
  k_mg_restrict<<<mg_grid(h, 3 * C.nn), 256, 0, h->stream>>>(L.nx, L.ny, L.nz, C.nx, C.ny, C.nz, l == 0 ? h->lplan.d_pbase : nullptr, m.z0, b, L.y, C.b);
  MG_KCHECK(h);
  NLS_TRY(mg_vcycle(h, l + 1, C.b, C.x));
  k_mg_prolong<<<mg_grid(h, 3 * L.nn), 256, 0, h->stream>>>(L.nx, L.ny, L.nz, C.nx, C.ny, C.nz, l == 0 ? h->lplan.d_pbase : nullptr, m.z0,
                                                            l == 0 ? h->d_con : nullptr, C.x, x);
  MG_KCHECK(h);
  return mg_cheb(h, l, b, x, m.degree, m.ratio, false);
}

// z = M^-1 r (r masked, ghost entries of z stay zero)
int nls_mg_apply(nls_context *h, const double *r, double *z) {
  if (h->mg.mode == 1) {        // block-Jacobi alone
    k_mg_apply_dinv<<<mg_grid(h, h->n_owned), 256, 0, h->stream>>>(h->n_owned, h->mg.lev[0].dinv, r, z);
    MG_KCHECK(h);
    return NLS_OK;
  }
  return mg_vcycle(h, 0, r, z);
}

// ---- preconditioned conjugate gradients around it (host-driven: an iteration is tens of launches, the three reductions
//      per iteration cost nothing beside them) ----
int nls_pcg_solve_mg(nls_context *h, double rtol, int maxits, int *its, double *relres, int *status) {
  const int64_t n = h->nrows;
  NLS_TRY(nls_mg_setup(h));
  MgLevel &L0 = h->mg.lev[0];
  double *x = h->d_dU, *r = h->d_r, *p = h->d_p, *Ap = h->d_Ap, *z = L0.x;
  k_mg_mask_copy<<<mg_grid(h, n), 256, 0, h->stream>>>(n, h->d_b, h->d_con, r); MG_KCHECK(h);
  NLS_CUDA(h, cudaMemsetAsync(x, 0, sizeof(double) * n, h->stream));
  double bb, rz, rr;
  NLS_TRY(nls_dev_dot(h, n, r, r, &bb));
  int it = 0, st = 1;
  rr = bb;
  if (bb == 0.0) st = 0;
  else {
    NLS_TRY(nls_mg_apply(h, r, z));
    NLS_TRY(nls_dev_dot(h, n, r, z, &rz));
    NLS_CUDA(h, cudaMemcpyAsync(p, z, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
    const double tol2 = rtol * rtol;
    while (it < maxits) {
      NLS_TRY(nls_halo_exchange(h, p));
      NLS_TRY(nls_launch_spmv(h, p, Ap, true, false));
      double pAp;
      NLS_TRY(nls_dev_dot(h, n, p, Ap, &pAp));
      if (!(pAp > 0.0) || !(rz == rz)) { st = 2; break; }
      const double alpha = rz / pAp;
      NLS_TRY(nls_launch_axpy(h, n, alpha, p, x));
      NLS_TRY(nls_launch_axpy(h, n, -alpha, Ap, r));
      ++it;
      NLS_TRY(nls_dev_dot(h, n, r, r, &rr));
      if (!(rr == rr)) { st = 2; break; }
      if (rr <= tol2 * bb) { st = 0; break; }
      NLS_TRY(nls_mg_apply(h, r, z));
      double rzn;
      NLS_TRY(nls_dev_dot(h, n, r, z, &rzn));
      const double beta = rzn / rz;
      rz = rzn;
      k_mg_xpby<<<mg_grid(h, n), 256, 0, h->stream>>>(n, z, beta, p); MG_KCHECK(h);
    }
  }
  if (its) *its = it;
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  if (status) *status = st;
  return NLS_OK;
}
