// kernels_mg.cu -- geometric multigrid preconditioner of the Krylov solve on 2-D and 3-D meshes with grid topology.
//
// Replaces, together with kernels_krylov.cu:  dofs(K) \ dofs(-R)  (src/nonlinearsolvers/newton_fem.jl:158). The reference
// factorises K; at 10^7 DoF the B200 path iterates, and Jacobi-PCG needs O(n) iterations per solve (1 600 at C3). Here the
// preconditioner is one V-cycle on a hierarchy of Galerkin operators K_l+1 = P^T K_l P:
//   * levels: the node grid coarsened by two per axis (coarse nodes = even fine nodes, plus the last node of an axis with an
//     even node count), P = trilinear interpolation (weights 1, 1/2 per axis), the same for the three displacement components;
//   * level 0 is the assembled CSR Jacobian itself (block positions from the line tables of kernels_lines3d.cu, Dirichlet
//     rows and columns masked out = the free block dofs(K)); coarser operators are stored as 27 dense 3 x 3 blocks per node;
//   * 2-D models run the same code: the node rows play the z-planes (ny = 1), every node carries a decoupled dummy third DoF on
//     the stencil levels (identity on the diagonal, zero right-hand side), level 0 keeps its two-component vectors;
//   * smoother: Chebyshev polynomial in D^-1 K (D = the 3 x 3 node blocks), largest eigenvalue by a few power iterations;
//     the coarsest grid (<= 6 nodes per axis) gets a long Chebyshev run instead of a factorisation;
//   * the Galerkin products are gathers: every coarse block is the sum of its <= 343 weighted fine blocks in a fixed order:
//     deterministic, no atomics.
// Multi-GPU (z-slabs): levels 0 and 1 are distributed like the mesh -- every rank owns the coarse planes whose fine plane it
// owns, one ghost plane either side, filled by neighbour send/recv (vectors before every product, operator rows once per
// set-up) -- so smoothing, restriction and prolongation are the single-GPU operations. From level 2 on (1/64 of the nodes) the
// grids are global and replicated: operator rows and right-hand sides are summed over the ranks (every rank contributes its
// planes, zeros elsewhere) and every rank runs the same coarse cycle. The iteration count does not depend on the rank count.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "nls_internal.h"
#include "nccl_dyn.h"

#ifndef NLS_TRY
#define NLS_TRY(call) do { int rc_ = (call); if (rc_ != NLS_OK) return rc_; } while (0)
#endif
static const bool mg_trace_on = std::getenv("NLS_B200_MG_TRACE") != nullptr;
#define MG_TRACE(h, ...) do { if (mg_trace_on) { cudaStreamSynchronize((h)->stream); fprintf(stderr, "[mg %d] ", (h)->rank); fprintf(stderr, __VA_ARGS__); fprintf(stderr, " (%s)\n", cudaGetErrorString(cudaGetLastError())); fflush(stderr); } } while (0)
#define MG_KCHECK(h) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string("multigrid kernel launch: ") + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

struct MgL3Line { int64_t lb; int32_t nl; int8_t rank[9]; int8_t pad[3]; };   // = L3Line of kernels_lines3d.cu

struct MgFine {          // level 0: the CSR Jacobian of a grid-topology mesh
  int dim;               // 3, or 2: the node rows of a 2-D mesh play the planes (ny = 1) and every node gets a decoupled dummy third DoF
  int nx, ny, nzo;       // owned planes
  int z0;                // line-table index of the first owned plane
  int p0, NZ;            // global index of the first owned plane, global plane count
  const int32_t *pbase;
  const MgL3Line *info;
  const double *vals;
  const uint8_t *con;
  const double *gbelow, *gabove;   // stencil-format rows [ny * nx][27][9] of the neighbours' boundary planes (global p0 - 1, p0 + nzo)
};

// a grid level stored as a box of z-planes: planes [zbase, zbase + nzb) of a global grid with NZ planes; the kernel works on
// the planes [zown0, zown1) (global indices), the rest of the box is read only. Replicated levels: the box is the whole grid.
struct MgBox { int nx, ny, nzb, zbase, NZ, zown0, zown1; };

__host__ __device__ __forceinline__ int mg_coarse_n(int n) { return n <= 2 ? n : ((n & 1) ? (n + 1) / 2 : n / 2 + 1); }
// fine index of coarse node X, and the interpolation weight of fine node x on it (0 if none); n = fine node count of the axis
__host__ __device__ __forceinline__ int mg_fx(int X, int n) { return n <= 2 ? X : (2 * X < n - 1 ? 2 * X : n - 1); }
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

// block (i, j) of the free block of the fine matrix, i = (x, y, zg), j = i + (dx, dy, dz); zg = GLOBAL plane index. Rows of
// owned planes come from the CSR values, rows of the two neighbouring planes from the exchanged stencil copies.
// Returns false (B untouched) when j does not exist.
__device__ __forceinline__ bool mg_fine_block(const MgFine &F, int x, int y, int zg, int dx, int dy, int dz, double (&B)[9]) {
  const int x2 = x + dx, y2 = y + dy, z2 = zg + dz;
  if (x2 < 0 || x2 >= F.nx || y2 < 0 || y2 >= F.ny || z2 < 0 || z2 >= F.NZ) return false;
  const int zl = zg - F.p0;
  if (zl < 0 || zl >= F.nzo) {
    const double *g = zl < 0 ? F.gbelow : F.gabove;
    if (!g || (zl != -1 && zl != F.nzo)) return false;
    const double *p = g + (((int64_t)y * F.nx + x) * 27 + (dz + 1) * 9 + (dy + 1) * 3 + dx + 1) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) B[k] = p[k];
    return true;
  }
  const int zp = F.z0 + zl;
  const MgL3Line *li = F.info + ((int64_t)zp * F.ny + y);
  const int rank = li->rank[(dz + 1) * 3 + (dy + 1)];
  if (rank < 0) return false;
  const int nl = li->nl, w = (x == 0 || x == F.nx - 1) ? 2 : 3;
  const int64_t b0 = li->lb + (int64_t)nl * (x == 0 ? 0 : 3 * x - 1);
  const int pos = rank * w + dx + (x == 0 ? 0 : 1);
  const int D = F.dim;
  const double *p = F.vals + (int64_t)D * D * b0 + D * pos;
  const int64_t stride = (int64_t)D * nl * w;
  const int64_t ni = (int64_t)F.pbase[zp] + (int64_t)y * F.nx + x, nj = (int64_t)F.pbase[zp + dz] + (int64_t)y2 * F.nx + x2;
  const uint8_t *ci = F.con + D * ni, *cj = F.con + D * nj;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (r < D && c < D) B[r * 3 + c] = (ci[r] | cj[c]) ? 0.0 : p[r * stride + c];
      else B[r * 3 + c] = (r == c && dx == 0 && dy == 0 && dz == 0) ? 1.0 : 0.0;      // the dummy DoF of a 2-D node: identity
    }
  return true;
}

// ---- level 0: inverse of the masked 3 x 3 node blocks (block-Jacobi); constrained DoFs get zero rows and columns ----
__global__ void __launch_bounds__(256) k_mg_fine_dinv(const MgFine F, double *__restrict__ dinv) {
  const int64_t nn = (int64_t)F.nx * F.ny * F.nzo;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nn; t += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(t % F.nx), y = (int)((t / F.nx) % F.ny), z = (int)(t / ((int64_t)F.nx * F.ny));
    double B[9], Bi[9];
    mg_fine_block(F, x, y, F.p0 + z, 0, 0, 0, B);
    const int64_t ni = (int64_t)F.pbase[F.z0 + z] + (int64_t)y * F.nx + x;
    const uint8_t *ci = F.con + F.dim * ni;
    bool cr[3] = {false, false, false};
    for (int r = 0; r < F.dim; ++r) cr[r] = ci[r] != 0;
    for (int r = 0; r < 3; ++r) if (cr[r]) B[r * 3 + r] = 1.0;
    mg_inv3(B, Bi);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) dinv[9 * ni + r * 3 + c] = (cr[r] | cr[c]) ? 0.0 : Bi[r * 3 + c];
  }
}

// level 0: the rows of owned plane zl in stencil format (for the neighbour rank's Galerkin product)
__global__ void __launch_bounds__(256) k_mg_export_rows(const MgFine F, int zl, double *__restrict__ out) {
  const int64_t total = (int64_t)F.nx * F.ny * 27;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(t % 27);
    const int64_t node = t / 27;
    const int x = (int)(node % F.nx), y = (int)(node / F.nx);
    double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    mg_fine_block(F, x, y, F.p0 + zl, s % 3 - 1, (s / 3) % 3 - 1, s / 9 - 1, B);
    for (int k = 0; k < 9; ++k) out[t * 9 + k] = B[k];
  }
}

// ---- Galerkin product level 1 <- level 0, one WARP per owned coarse node: lane s owns stencil slot s. For every fine node i
// under the coarse node (<= 27) the 27 lanes first fetch the 27 blocks of i's block row -- one contiguous 1.9 KB CSR row, each
// block once -- into shared memory; then every lane adds the <= 8 of them whose column node interpolates from its coarse
// neighbour. Rows of the coarse box's owned planes only (ghost rows arrive by exchange).
__global__ void __launch_bounds__(128) k_mg_rap_fine_warp(const MgFine F, const MgBox C, double *__restrict__ Kc) {
  __shared__ double sB[4][27][9];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int fnx = F.nx, fny = F.ny, fnz = F.NZ, cnx = C.nx, cny = C.ny;
  const int64_t ncoarse = (int64_t)cnx * cny * (C.zown1 - C.zown0);
  const int dxs = lane % 3 - 1, dys = (lane / 3) % 3 - 1, dzs = lane / 9 - 1;      // lane < 27: both the block it fetches and its slot
  for (int64_t I = blockIdx.x * 4ll + wib; I < ncoarse; I += (int64_t)gridDim.x * 4) {
    const int X = (int)(I % cnx), Y = (int)((I / cnx) % cny), Z = C.zown0 + (int)(I / ((int64_t)cnx * cny));
    const int X2 = X + dxs, Y2 = Y + dys, Z2 = Z + dzs;
    const bool jok = lane < 27 && X2 >= 0 && X2 < cnx && Y2 >= 0 && Y2 < cny && Z2 >= 0 && Z2 < C.NZ;
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
      double *o = Kc + (((((int64_t)(Z - C.zbase)) * cny + Y) * cnx + X) * 27 + lane) * 9;
#pragma unroll
      for (int k = 0; k < 9; ++k) o[k] = acc[k];
    }
  }
}

// ---- Galerkin product between stencil levels, one thread per (coarse node of the planes [zown0, zown1), stencil slot). The
// fine level is a box (rows of its ghost planes present); the coarse rows go to their place in the coarse box.
__global__ void __launch_bounds__(128) k_mg_rap(const MgBox Fb, const double *__restrict__ Kf, const MgBox C, double *__restrict__ Kc) {
  const int fnx = Fb.nx, fny = Fb.ny, fnz = Fb.NZ, cnx = C.nx, cny = C.ny;
  const int64_t total = (int64_t)cnx * cny * (C.zown1 - C.zown0) * 27;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(t % 27);
    const int64_t I = t / 27;
    const int X = (int)(I % cnx), Y = (int)((I / cnx) % cny), Z = C.zown0 + (int)(I / ((int64_t)cnx * cny));
    const int X2 = X + s % 3 - 1, Y2 = Y + (s / 3) % 3 - 1, Z2 = Z + s / 9 - 1;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (X2 >= 0 && X2 < cnx && Y2 >= 0 && Y2 < cny && Z2 >= 0 && Z2 < C.NZ) {
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
                  const int sf = (z2 - z + 1) * 9 + (y2 - y + 1) * 3 + (x2 - x + 1);
                  const double *B = Kf + ((((int64_t)(z - Fb.zbase) * fny + y) * fnx + x) * 27 + sf) * 9;
#pragma unroll
                  for (int k = 0; k < 9; ++k) acc[k] = fma(w, B[k], acc[k]);
                }
              }
            }
          }
        }
      }
    }
    double *o = Kc + (((((int64_t)(Z - C.zbase)) * cny + Y) * cnx + X) * 27 + s) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = acc[k];
  }
}

// stencil levels: inverse of the diagonal blocks (slot 13) of nodes [0, nn); a singular block (a coarse node whose fine nodes
// are all constrained) gets zero
__global__ void __launch_bounds__(256) k_mg_dinv(int64_t nn, const double *__restrict__ K, double *__restrict__ dinv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double Bi[9];
    mg_inv3(K + (i * 27 + 13) * 9, Bi);
    for (int k = 0; k < 9; ++k) dinv[9 * i + k] = Bi[k];
  }
}

// stencil levels: y = K x on the planes [zown0, zown1) of a box, three lanes per node (one per scalar row), 27 blocks each
__global__ void __launch_bounds__(256) k_mg_spmv(const MgBox B, const double *__restrict__ K, const double *__restrict__ x, double *__restrict__ y) {
  const int nx = B.nx, ny = B.ny;
  const int64_t first = (int64_t)(B.zown0 - B.zbase) * ny * nx * 3, nrows = (int64_t)(B.zown1 - B.zown0) * ny * nx * 3;
  for (int64_t tt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; tt < nrows; tt += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = first + tt, i = t / 3;
    const int r = (int)(t - 3 * i);
    const int X = (int)(i % nx), Y = (int)((i / nx) % ny), Z = (int)(i / ((int64_t)nx * ny));      // Z: plane inside the box
    const double *Ki = K + i * 243 + r * 3;
    double s = 0.0;
    for (int dz = -1; dz <= 1; ++dz) {
      if (Z + dz < 0 || Z + dz >= B.nzb) continue;
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
// mode 1: x = 0 on entry (x = d), 2: new run (d restarts), 0: inside a run. NC = components per node of the vectors (2: the
// level-0 vectors of a 2-D model; D^-1 is block diagonal there, its leading 2 x 2 block is the inverse of the node block).
template <int NC>
__global__ void __launch_bounds__(256) k_mg_cheb(int64_t nn, const double *__restrict__ dinv, const double *__restrict__ b,
                                                  const double *__restrict__ y, double *__restrict__ d, double *__restrict__ x,
                                                  double a, double c, int mode) {
  const bool first = mode == 1, restart = mode != 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double r[NC], z[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { r[k] = b[NC * i + k]; if (y) r[k] -= y[NC * i + k]; }
    const double *D = dinv + 9 * i;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      z[k] = 0.0;
#pragma unroll
      for (int j = 0; j < NC; ++j) z[k] = fma(D[3 * k + j], r[j], z[k]);
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      double dk = c * z[k];
      if (!restart) dk = fma(a, d[NC * i + k], dk);
      d[NC * i + k] = dk;
      if (first) x[NC * i + k] = dk; else x[NC * i + k] += dk;
    }
  }
}

// z = D^-1 r on nodes [0, nn) (block-Jacobi alone, and the power iteration)
template <int NC>
__global__ void __launch_bounds__(256) k_mg_apply_dinv(int64_t nn, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    const double *D = dinv + 9 * i;
    double rr[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) rr[k] = r[NC * i + k];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NC; ++j) s = fma(D[3 * k + j], rr[j], s);
      z[NC * i + k] = s;
    }
  }
}

// r = b - y on n entries
__global__ void __launch_bounds__(256) k_mg_residual(int64_t n, const double *__restrict__ b, const double *__restrict__ y, double *__restrict__ r) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] = b[i] - y[i];
}

// restriction bc = P^T r on the coarse planes [zown0, zown1): one thread per coarse DoF gathers its <= 27 fine nodes (ghost
// planes of r filled beforehand). Level 0: r is indexed by node id through the line tables (pbase; z0 = table index of the
// rank's first owned plane, global plane p0).
__global__ void __launch_bounds__(256) k_mg_restrict(const MgBox Fb, const int32_t *__restrict__ pbase, int z0, int p0, int fnc,
                                                      const MgBox C, const double *__restrict__ r, double *__restrict__ bc) {
  const int fnx = Fb.nx, fny = Fb.ny, fnz = Fb.NZ, cnx = C.nx, cny = C.ny;
  const int64_t total = (int64_t)cnx * cny * (C.zown1 - C.zown0) * 3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = t / 3;
    const int c = (int)(t - 3 * I);
    const int X = (int)(I % cnx), Y = (int)((I / cnx) % cny), Z = C.zown0 + (int)(I / ((int64_t)cnx * cny));
    const int fx = mg_fx(X, fnx), fy = mg_fx(Y, fny), fz = mg_fx(Z, fnz);
    double s = 0.0;
    for (int z = max(fz - 1, 0); z <= min(fz + 1, fnz - 1) && c < fnc; ++z) {      // c >= fnc: the dummy DoF of a 2-D model
      const double wz = mg_w(z, Z, fnz);
      if (wz == 0.0) continue;
      for (int yy = max(fy - 1, 0); yy <= min(fy + 1, fny - 1); ++yy) {
        const double wy = wz * mg_w(yy, Y, fny);
        if (wy == 0.0) continue;
        for (int x = max(fx - 1, 0); x <= min(fx + 1, fnx - 1); ++x) {
          const double w = wy * mg_w(x, X, fnx);
          if (w == 0.0) continue;
          const int64_t i = pbase ? (int64_t)pbase[z0 + z - p0] + (int64_t)yy * fnx + x : ((int64_t)(z - Fb.zbase) * fny + yy) * fnx + x;
          s = fma(w, r[(int64_t)fnc * i + c], s);
        }
      }
    }
    bc[3 * ((((int64_t)(Z - C.zbase)) * cny + Y) * cnx + X) + c] = s;
  }
}

// prolongation x += P xc on the fine planes [zown0, zown1): one thread per fine DoF gathers its <= 8 coarse parents (ghost
// planes of xc filled beforehand); constrained DoFs (level 0) stay zero
__global__ void __launch_bounds__(256) k_mg_prolong(const MgBox Fb, const int32_t *__restrict__ pbase, int z0, int p0, int fnc,
                                                     const uint8_t *__restrict__ con, const MgBox C, const double *__restrict__ xc, double *__restrict__ x) {
  const int fnx = Fb.nx, fny = Fb.ny, fnz = Fb.NZ, cnx = C.nx, cny = C.ny;
  const int64_t total = (int64_t)fnx * fny * (Fb.zown1 - Fb.zown0) * fnc;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t il = t / fnc;
    const int r = (int)(t - fnc * il);
    const int xx = (int)(il % fnx), yy = (int)((il / fnx) % fny), zz = Fb.zown0 + (int)(il / ((int64_t)fnx * fny));     // zz: global plane
    const int64_t i = pbase ? (int64_t)pbase[z0 + zz - p0] + (int64_t)yy * fnx + xx : ((int64_t)(zz - Fb.zbase) * fny + yy) * fnx + xx;
    if (con && con[(int64_t)fnc * i + r]) continue;
    // parents per axis: the coarse node at v (weight 1), or the two at v - 1 and v + 1 (weight 1/2 each)
    int px[2], py[2], pz[2];
    auto parents = [](int v, int n, int *p) -> int {
      if (n <= 2) { p[0] = v; return 1; }
      if (!(v & 1)) { p[0] = v >> 1; return 1; }
      if (v == n - 1) { p[0] = mg_coarse_n(n) - 1; return 1; }
      p[0] = (v - 1) >> 1; p[1] = (v + 1 == n - 1) ? mg_coarse_n(n) - 1 : (v + 1) >> 1;
      return 2;
    };
    const int npx = parents(xx, fnx, px), npy = parents(yy, fny, py), npz = parents(zz, fnz, pz);
    const double w = 1.0 / (double)(npx * npy * npz);
    double s = 0.0;
    for (int c = 0; c < npz; ++c)
      for (int bq = 0; bq < npy; ++bq)
        for (int a = 0; a < npx; ++a) s += xc[3 * ((((int64_t)(pz[c] - C.zbase) * cny) + py[bq]) * cnx + px[a]) + r];
    x[(int64_t)fnc * i + r] += w * s;
  }
}

__global__ void __launch_bounds__(256) k_mg_xpby(int64_t n, const double *__restrict__ z, double beta, double *__restrict__ p) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}
__global__ void __launch_bounds__(256) k_mg_mask_copy(int64_t n, const double *__restrict__ b, const uint8_t *__restrict__ con, double *__restrict__ r) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] = con[i] ? 0.0 : b[i];
}
__global__ void __launch_bounds__(256) k_mg_con_to_f64(int64_t n, const uint8_t *__restrict__ con, double *__restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = con[i] ? 1.0 : 0.0;
}
__global__ void __launch_bounds__(256) k_mg_f64_to_con(int64_t n, const double *__restrict__ v, uint8_t *__restrict__ con) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) con[i] = v[i] != 0.0;
}
__global__ void __launch_bounds__(256) k_mg_fill_hash(int64_t n, unsigned seed, int64_t gofs, int zero_third, double *__restrict__ v) {   // fixed pseudo-random start vector
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long s = (unsigned long long)(i + gofs) * 0x9E3779B97F4A7C15ull + seed;
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
    v[i] = (zero_third && i % 3 == 2) ? 0.0 : (double)(s & 0xffff) / 65536.0 - 0.5;      // 2-D models: the dummy DoF stays out of the eigenvalue estimate
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

static void mg_drop_graph(nls_context *h) {
  MgPlan &m = h->mg;
  if (m.vgraph) { cudaGraphExecDestroy(m.vgraph); m.vgraph = nullptr; }
  m.vgraph_version = -1;
}

static void mg_free_levels(nls_context *h) {
  MgPlan &m = h->mg;
  mg_drop_graph(h);               // it holds the addresses of the level buffers
  for (int l = 0; l < MG_MAX_LEVELS; ++l) {
    MgLevel &L = m.lev[l];
    if (L.K) cudaFree(L.K);
    if (L.dinv) cudaFree(L.dinv);
    for (double *v : {L.x, L.b, L.y, L.d}) if (v) cudaFree(v);
    L = MgLevel();
  }
  if (m.vals32) cudaFree(m.vals32);
  for (double *v : {m.gbelow, m.gabove, m.gsend, m.rf}) if (v) cudaFree(v);
  if (m.conf) cudaFree(m.conf);
  if (m.d_vote) cudaFree(m.d_vote);
  m.d_vote = nullptr;
  m.vals32 = nullptr; m.gbelow = m.gabove = m.gsend = m.rf = nullptr; m.conf = nullptr;
  m.laid_out = false; m.version = -1; m.power_version = -1000; m.nlev = 0;
}

void nls_mg_free(nls_context *h) {
  mg_free_levels(h);
  MgPlan &m = h->mg;
  m.ok = false;
}

// the communicator changed (or appeared): the levels are laid out again at the next set-up
void nls_mg_invalidate(nls_context *h) { mg_free_levels(h); }

// can this mesh have a hierarchy? (called when the line plan is built; the levels themselves are laid out at the first
// set-up, when the communicator -- if any -- exists)
int nls_mg_build(nls_context *h) {
  nls_mg_free(h);
  MgPlan &m = h->mg;
  const Lines3Plan &p = h->lplan;
  if (!p.tables || (h->dim != 3 && h->dim != 2) || p.h_pown.empty()) return NLS_OK;
  int z0 = -1, z1 = -1;
  for (int z = 0; z < p.nzp; ++z) if (p.h_pown[z]) { if (z0 < 0) z0 = z; z1 = z + 1; }
  if (z0 < 0) return NLS_OK;
  for (int z = z0; z < z1; ++z) if (!p.h_pown[z]) return NLS_OK;       // owned planes must be contiguous
  if ((int64_t)p.nx * p.ny * (z1 - z0) != h->n_owned) return NLS_OK;
  if (z0 > 1 || p.nzp - z1 > 1) return NLS_OK;                          // at most one ghost plane either side
  if (std::max(p.nx, std::max(p.ny, z1 - z0)) <= 6 && z0 == 0 && p.nzp == z1) return NLS_OK;   // too small for a second level
  m.z0 = z0; m.nzo = z1 - z0;
  m.ghost_below = z0 == 1; m.ghost_above = p.nzp - z1 == 1;
  m.ok = true;
  return NLS_OK;
}

// single GPU: lay the levels out right away (multi-GPU: at the first set-up, when the communicator exists, or through
// nls_solver_prepare, which is collective)
static int mg_layout(nls_context *h);
static int mg_exchange(nls_context *h, int l, double *a, int npd);
static int mg_allsum(nls_context *h, double *a, size_t n);
int nls_mg_prepare(nls_context *h, bool collective) {
  MgPlan &m = h->mg;
  if (!m.ok || (m.laid_out && m.dist == (h->distributed && h->nranks > 1))) return NLS_OK;
  const bool ghosts = m.ghost_below || m.ghost_above;
  if (!h->distributed && ghosts) return NLS_OK;
  if (h->distributed && !collective) return NLS_OK;
  NLS_TRY(mg_layout(h));
  if (m.fp32 && !m.vals32 && cudaMalloc((void **)&m.vals32, sizeof(float) * (size_t)h->nnz) != cudaSuccess) { cudaGetLastError(); m.vals32 = nullptr; }
  if (m.dist) {      // NCCL opens its point-to-point channels at the first send/recv between two ranks (~0.3 s): do it here
    NLS_TRY(mg_exchange(h, 1, m.lev[1].x, 3));
    if (m.nlev > 2) NLS_TRY(mg_allsum(h, m.lev[2].b, 3 * (size_t)m.lev[2].nn));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return NLS_OK;
}

static MgBox mg_box(const MgLevel &L) { MgBox b; b.nx = L.nx; b.ny = L.ny; b.nzb = L.nzb; b.zbase = L.zbase; b.NZ = L.NZ; b.zown0 = L.zown0; b.zown1 = L.zown1; return b; }

// levels, ownership of the coarse planes, buffers. Collective when distributed.
static int mg_layout(nls_context *h) {
  MgPlan &m = h->mg;
  const Lines3Plan &p = h->lplan;
  mg_free_levels(h);
  m.dist = h->distributed && h->nranks > 1;
  m.rank_below = m.rank_above = -1;
  int p0 = 0, NZ = m.nzo;
  if (m.dist) {
    // plane counts of all ranks (slabs are stacked in rank order)
    std::vector<double> cnt((size_t)h->nranks, 0.0);
    double *d_cnt = nullptr;
    NLS_CUDA(h, cudaMalloc((void **)&d_cnt, sizeof(double) * h->nranks));
    cnt[h->rank] = (double)m.nzo;
    NLS_CUDA(h, cudaMemcpyAsync(d_cnt, cnt.data(), sizeof(double) * h->nranks, cudaMemcpyHostToDevice, h->stream));
    if (h->nccl->AllReduce(d_cnt, d_cnt, (size_t)h->nranks, NCCL_DYN_FLOAT64, NCCL_DYN_SUM, h->comm, h->stream) != 0)
      NLS_FAIL(h, NLS_ERR_NCCL, "multigrid: ncclAllReduce (plane counts) failed");
    NLS_CUDA(h, cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(double) * h->nranks, cudaMemcpyDeviceToHost, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_cnt);
    NZ = 0;
    for (int r = 0; r < h->nranks; ++r) { if (r == h->rank) p0 = NZ; NZ += (int)cnt[r]; }
    if (m.ghost_below) m.rank_below = h->rank - 1;
    if (m.ghost_above) m.rank_above = h->rank + 1;
    if (m.ghost_below != (p0 > 0) || m.ghost_above != (p0 + m.nzo < NZ))
      NLS_FAIL(h, NLS_ERR_STATE, "multigrid: the slabs are not stacked in rank order");
  } else if (m.ghost_below || m.ghost_above) {
    NLS_FAIL(h, NLS_ERR_STATE, "multigrid: ghost planes without a communicator");
  }
  m.p0 = p0; m.NZ0 = NZ;
  MG_TRACE(h, "layout: dist %d p0 %d nzo %d NZ %d below %d above %d", (int)m.dist, p0, m.nzo, NZ, m.rank_below, m.rank_above);
  int nx = p.nx, ny = p.ny, nz = NZ, l = 0;
  int own0 = p0, own1 = p0 + m.nzo;       // owned global planes on the current level
  for (;;) {
    MgLevel &L = m.lev[l];
    L.nx = nx; L.ny = ny; L.NZ = nz;
    L.rown0 = own0; L.rown1 = own1;
    if (l <= 1) {                          // distributed levels: owned planes + ghost planes
      L.zown0 = own0; L.zown1 = own1;
      L.zbase = own0 - ((own0 > 0 && m.dist) ? 1 : 0);
      const int ztop = own1 + ((own1 < nz && m.dist) ? 1 : 0);
      L.nzb = ztop - L.zbase;
    } else {                               // replicated levels: the whole grid on every rank
      L.zown0 = 0; L.zown1 = nz; L.zbase = 0; L.nzb = nz;
    }
    L.nn = (int64_t)nx * ny * L.nzb;
    const size_t nd = (size_t)(l == 0 ? h->ndof_local() : 3 * L.nn);
    if (l > 0) NLS_CUDA(h, cudaMalloc((void **)&L.K, sizeof(double) * 243 * (size_t)L.nn));
    NLS_CUDA(h, cudaMalloc((void **)&L.dinv, sizeof(double) * 9 * (size_t)(l == 0 ? h->nn : L.nn)));
    NLS_CUDA(h, cudaMemsetAsync(L.dinv, 0, sizeof(double) * 9 * (size_t)(l == 0 ? h->nn : L.nn), h->stream));
    for (double **v : {&L.x, &L.b, &L.y, &L.d}) {
      if (l == 0 && (v == &L.b)) continue;                  // level 0: b is the caller's residual
      NLS_CUDA(h, cudaMalloc((void **)v, sizeof(double) * nd));
      NLS_CUDA(h, cudaMemsetAsync(*v, 0, sizeof(double) * nd, h->stream));
    }
    ++l;
    const int cx = mg_coarse_n(nx), cy = mg_coarse_n(ny), cz = mg_coarse_n(nz);
    if (l >= std::min(MG_MAX_LEVELS, m.max_levels) || std::max(nx, std::max(ny, nz)) <= 6 || (cx == nx && cy == ny && cz == nz)) break;
    // Slender domains (the weak-scaled column: 150 x 150 x 150 N nodes): once the thinnest axis is down to 6 nodes, further
    // levels have trilinear elements longer than the body is thick; they lock in bending and the coarse correction of the
    // bending modes is too small (48 x 48 x 384 nodes: 20 Krylov iterations per solve against 9 for the cube; 13 with the
    // hierarchy cut here). The cut level is small enough for the Chebyshev coarse solve, whose degree follows its length.
    {
      int thin = 1 << 30;
      for (int n : {nx, ny, nz}) if (n > 2) thin = std::min(thin, n);
      if (l >= 2 && thin <= 6 && (int64_t)nx * ny * nz <= 4096) break;      // (a hierarchy has at least two levels)
    }
    // coarse planes of this rank: those whose fine plane it owns
    int c0 = 0, c1 = 0;
    while (c0 < cz && mg_fx(c0, nz) < own0) ++c0;
    c1 = c0;
    while (c1 < cz && mg_fx(c1, nz) < own1) ++c1;
    if (m.dist && l == 1 && c1 == c0) NLS_FAIL(h, NLS_ERR_STATE, "multigrid: a rank owns no coarse plane (slab too thin)");
    own0 = c0; own1 = c1;
    nx = cx; ny = cy; nz = cz;
  }
  MG_TRACE(h, "layout: %d levels", l);
  m.nlev = l;
  if (m.nlev < 2) { m.ok = false; NLS_FAIL(h, NLS_ERR_STATE, "multigrid: the grid is too small for a hierarchy"); }
  const size_t plane_rows = (size_t)p.nx * p.ny * 243;
  if (m.dist) {
    if (m.ghost_below) NLS_CUDA(h, cudaMalloc((void **)&m.gbelow, sizeof(double) * plane_rows));
    if (m.ghost_above) NLS_CUDA(h, cudaMalloc((void **)&m.gabove, sizeof(double) * plane_rows));
    NLS_CUDA(h, cudaMalloc((void **)&m.gsend, sizeof(double) * 2 * plane_rows));
  }
  NLS_CUDA(h, cudaMalloc((void **)&m.rf, sizeof(double) * (size_t)h->ndof_local()));
  NLS_CUDA(h, cudaMemsetAsync(m.rf, 0, sizeof(double) * (size_t)h->ndof_local(), h->stream));
  NLS_CUDA(h, cudaMalloc((void **)&m.conf, (size_t)h->ndof_local()));
  m.laid_out = true;
  m.version = -1;
  return NLS_OK;
}

static MgFine mg_fine_args(nls_context *h) {
  const Lines3Plan &p = h->lplan;
  const MgPlan &m = h->mg;
  MgFine F;
  F.dim = h->dim;
  F.nx = p.nx; F.ny = p.ny; F.nzo = m.nzo; F.z0 = m.z0; F.p0 = m.p0; F.NZ = m.NZ0;
  F.pbase = p.d_pbase; F.info = reinterpret_cast<const MgL3Line *>(p.d_info);
  F.vals = h->d_vals; F.con = m.conf;
  F.gbelow = m.gbelow; F.gabove = m.gabove;
  return F;
}

// ghost planes of a distributed stencil-level array (npd doubles per node) <- the neighbours' boundary planes
static int mg_exchange(nls_context *h, int l, double *a, int npd) {
  MgPlan &m = h->mg;
  if (!m.dist || l != 1) return NLS_OK;
  MgLevel &L = m.lev[l];
  const size_t pl = (size_t)L.nx * L.ny * npd;
  const int g0 = L.zown0 - L.zbase, nown = L.zown1 - L.zown0;
  NcclApi *nc = h->nccl;
  if (nc->GroupStart() != 0) NLS_FAIL(h, NLS_ERR_NCCL, "ncclGroupStart failed");
  int e = 0;
  if (m.rank_below >= 0) {
    e |= nc->Send(a + (size_t)g0 * pl, pl, NCCL_DYN_FLOAT64, m.rank_below, h->comm, h->stream);
    e |= nc->Recv(a, pl, NCCL_DYN_FLOAT64, m.rank_below, h->comm, h->stream);
  }
  if (m.rank_above >= 0) {
    e |= nc->Send(a + (size_t)(g0 + nown - 1) * pl, pl, NCCL_DYN_FLOAT64, m.rank_above, h->comm, h->stream);
    e |= nc->Recv(a + (size_t)(g0 + nown) * pl, pl, NCCL_DYN_FLOAT64, m.rank_above, h->comm, h->stream);
  }
  const int ge = nc->GroupEnd();
  if (e || ge) NLS_FAIL(h, NLS_ERR_NCCL, "multigrid: plane exchange (ncclSend/Recv) failed");
  return NLS_OK;
}

// sum over the ranks of a replicated-level array (every rank wrote its planes, zeros elsewhere)
static int mg_allsum(nls_context *h, double *a, size_t n) {
  if (!h->mg.dist) return NLS_OK;
  if (h->nccl->AllReduce(a, a, n, NCCL_DYN_FLOAT64, NCCL_DYN_SUM, h->comm, h->stream) != 0) NLS_FAIL(h, NLS_ERR_NCCL, "multigrid: ncclAllReduce failed");
  return NLS_OK;
}

// y = K_l x on the owned rows; x's ghost entries are brought in first on the distributed levels
static int mg_apply_K(nls_context *h, int l, double *x, double *y) {
  MgPlan &m = h->mg;
  MgLevel &L = m.lev[l];
  if (l == 0) {
    if (m.dist) NLS_TRY(nls_halo_exchange(h, x));
    return (m.fp32 && m.vals32) ? nls_launch_spmv_f32(h, m.vals32, x, y) : nls_launch_spmv(h, x, y, true, false);
  }
  NLS_TRY(mg_exchange(h, l, x, 3));
  k_mg_spmv<<<mg_grid(h, 3 * (int64_t)L.nx * L.ny * (L.zown1 - L.zown0)), 256, 0, h->stream>>>(mg_box(L), L.K, x, y);
  MG_KCHECK(h);
  return NLS_OK;
}

// components per node of a level's vectors: the model's on level 0, three on the stencil levels (2-D: two + the dummy)
static inline int mg_nc(nls_context *h, int l) { return l == 0 ? h->dim : 3; }

// owned part of a level's vectors: offset (doubles) and node count
static void mg_owned(nls_context *h, int l, int64_t *ofs, int64_t *nn) {
  const MgLevel &L = h->mg.lev[l];
  if (l == 0) { *ofs = 0; *nn = h->n_owned; }
  else { *ofs = 3 * (int64_t)(L.zown0 - L.zbase) * L.nx * L.ny; *nn = (int64_t)(L.zown1 - L.zown0) * L.nx * L.ny; }
}

static int mg_dot(nls_context *h, int l, int64_t n, const double *x, const double *y, double *out) {
  return l <= 1 ? nls_dev_dot(h, n, x, y, out) : nls_dev_dot_local(h, n, x, y, out);     // replicated levels: same value on every rank
}

// largest eigenvalue of D^-1 K on level l by power iteration (fixed pseudo-random start vector; the estimate is inflated
// by 10 %, the usual safeguard of a Chebyshev smoother)
static int mg_lambda_max(nls_context *h, int l, double *out) {
  MgPlan &m = h->mg;
  MgLevel &L = m.lev[l];
  int64_t ofs, nn;
  mg_owned(h, l, &ofs, &nn);
  const int nc = mg_nc(h, l);
  const int64_t n = nc * nn;
  const int64_t gofs = l == 0 ? nc * (int64_t)m.p0 * L.nx * L.ny : 3 * (int64_t)L.zown0 * L.nx * L.ny;
  double *x = L.x + ofs, *y = L.y + ofs, *d = L.d + ofs;
  k_mg_fill_hash<<<mg_grid(h, n), 256, 0, h->stream>>>(n, (unsigned)l, gofs, (h->dim == 2 && l >= 1) ? 1 : 0, x); MG_KCHECK(h);
  if (l == 0) { k_mg_mask_copy<<<mg_grid(h, n), 256, 0, h->stream>>>(n, x, h->d_con, x); MG_KCHECK(h); }
  double lam = 1.0;
  for (int it = 0; it < m.power_its; ++it) {
    double nrm2;
    NLS_TRY(mg_dot(h, l, n, x, x, &nrm2));
    if (!(nrm2 > 0.0)) { lam = 1.0; break; }
    k_mg_scale<<<mg_grid(h, n), 256, 0, h->stream>>>(n, 1.0 / sqrt(nrm2), x); MG_KCHECK(h);
    NLS_TRY(mg_apply_K(h, l, L.x, L.y));
    if (nc == 2) k_mg_apply_dinv<2><<<mg_grid(h, nn), 256, 0, h->stream>>>(nn, L.dinv + 3 * ofs, y, d);
    else k_mg_apply_dinv<3><<<mg_grid(h, nn), 256, 0, h->stream>>>(nn, L.dinv + 3 * ofs, y, d);
    MG_KCHECK(h);
    double xd;
    NLS_TRY(mg_dot(h, l, n, x, d, &xd));
    lam = xd;                                    // Rayleigh quotient of the normalised iterate
    NLS_CUDA(h, cudaMemcpyAsync(x, d, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
  }
  NLS_CUDA(h, cudaMemsetAsync(x, 0, sizeof(double) * n, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(d, 0, sizeof(double) * n, h->stream));
  *out = 1.1 * fabs(lam);
  return NLS_OK;
}

// Galerkin operators, block inverses and eigenvalue bounds for the current CSR values. Collective when distributed.
int nls_mg_setup(nls_context *h) {
  MgPlan &m = h->mg;
  if (!m.ok) NLS_FAIL(h, NLS_ERR_STATE, "multigrid: no hierarchy for this mesh");
  if (!m.laid_out || m.dist != (h->distributed && h->nranks > 1)) NLS_TRY(mg_layout(h));
  if (m.version == h->vals_version) return NLS_OK;
  {   // the constrained mask of every local node: ghost nodes are constrained by their owners only
    const int64_t nl = h->ndof_local();
    k_mg_con_to_f64<<<mg_grid(h, nl), 256, 0, h->stream>>>(nl, h->d_con, m.rf); MG_KCHECK(h);
    if (m.dist) NLS_TRY(nls_halo_exchange(h, m.rf));
    k_mg_f64_to_con<<<mg_grid(h, nl), 256, 0, h->stream>>>(nl, m.rf, m.conf); MG_KCHECK(h);
    NLS_CUDA(h, cudaMemsetAsync(m.rf, 0, sizeof(double) * (size_t)nl, h->stream));
  }
  MG_TRACE(h, "setup: masks done");
  MgFine F = mg_fine_args(h);
  k_mg_fine_dinv<<<mg_grid(h, (int64_t)F.nx * F.ny * F.nzo), 256, 0, h->stream>>>(F, m.lev[0].dinv); MG_KCHECK(h);
  if (m.dist) {      // boundary row planes of the CSR matrix, in stencil format, to the neighbours
    const size_t pr = (size_t)F.nx * F.ny * 243;
    NcclApi *nc = h->nccl;
    if (m.rank_below >= 0) { k_mg_export_rows<<<mg_grid(h, (int64_t)F.nx * F.ny * 27), 256, 0, h->stream>>>(F, 0, m.gsend); MG_KCHECK(h); }
    if (m.rank_above >= 0) { k_mg_export_rows<<<mg_grid(h, (int64_t)F.nx * F.ny * 27), 256, 0, h->stream>>>(F, F.nzo - 1, m.gsend + pr); MG_KCHECK(h); }
    if (nc->GroupStart() != 0) NLS_FAIL(h, NLS_ERR_NCCL, "ncclGroupStart failed");
    int e = 0;
    if (m.rank_below >= 0) { e |= nc->Send(m.gsend, pr, NCCL_DYN_FLOAT64, m.rank_below, h->comm, h->stream); e |= nc->Recv(m.gbelow, pr, NCCL_DYN_FLOAT64, m.rank_below, h->comm, h->stream); }
    if (m.rank_above >= 0) { e |= nc->Send(m.gsend + pr, pr, NCCL_DYN_FLOAT64, m.rank_above, h->comm, h->stream); e |= nc->Recv(m.gabove, pr, NCCL_DYN_FLOAT64, m.rank_above, h->comm, h->stream); }
    const int ge = nc->GroupEnd();
    if (e || ge) NLS_FAIL(h, NLS_ERR_NCCL, "multigrid: row exchange (ncclSend/Recv) failed");
  }
  MG_TRACE(h, "setup: rows exchanged");
  for (int l = 1; l < m.nlev; ++l) {
    MgLevel &Lf = m.lev[l - 1], &Lc = m.lev[l];
    MgBox C = mg_box(Lc);
    MG_TRACE(h, "setup: level %d grid %d x %d x %d own [%d,%d) box base %d planes %d", l, Lc.nx, Lc.ny, Lc.NZ, Lc.zown0, Lc.zown1, Lc.zbase, Lc.nzb);
    // levels >= 2 are replicated: every rank computes the rows whose fine planes it owns (all of them from level 3 on, where
    // the fine operator is already whole on every rank), zeros elsewhere, and the ranks sum
    const bool partial = l == 2 && m.dist;
    if (partial) { C.zown0 = Lc.rown0; C.zown1 = Lc.rown1; NLS_CUDA(h, cudaMemsetAsync(Lc.K, 0, sizeof(double) * 243 * (size_t)Lc.nn, h->stream)); }
    const int64_t nown = (int64_t)Lc.nx * Lc.ny * (C.zown1 - C.zown0);
    if (nown > 0) {
      if (l == 1) k_mg_rap_fine_warp<<<mg_grid(h, nown * 32, 128), 128, 0, h->stream>>>(F, C, Lc.K);
      else k_mg_rap<<<mg_grid(h, nown * 27, 128), 128, 0, h->stream>>>(mg_box(Lf), Lf.K, C, Lc.K);
      MG_KCHECK(h);
    }
    if (l == 1) NLS_TRY(mg_exchange(h, 1, Lc.K, 243));
    else if (partial) NLS_TRY(mg_allsum(h, Lc.K, 243 * (size_t)Lc.nn));
    k_mg_dinv<<<mg_grid(h, Lc.nn), 256, 0, h->stream>>>(Lc.nn, Lc.K, Lc.dinv); MG_KCHECK(h);
  }
  if (m.fp32) {
    if (!m.vals32 && cudaMalloc((void **)&m.vals32, sizeof(float) * (size_t)h->nnz) != cudaSuccess) { cudaGetLastError(); m.vals32 = nullptr; }
    if (m.vals32) NLS_TRY(nls_launch_vals_to_f32(h, m.vals32));
  }
  // eigenvalue bounds: the Jacobian of consecutive Newton iterations changes by a few per cent at most, well inside the
  // 10 % safety margin: a full power iteration every eighth set of values, in between the bounds are kept
  if (m.power_version < 0 || h->vals_version - m.power_version >= 8 || h->vals_version < m.power_version) {
    for (int l = 0; l < m.nlev; ++l) { NLS_TRY(mg_lambda_max(h, l, &m.lev[l].lmax)); MG_TRACE(h, "setup: lambda_max level %d = %g", l, m.lev[l].lmax); }
    m.power_version = h->vals_version;
  }
  m.version = h->vals_version;
  return NLS_OK;
}

// Chebyshev smoothing of K_l x = b on [lmax / ratio, lmax]; zero = x starts from zero. x, b: whole level arrays.
static int mg_cheb(nls_context *h, int l, const double *b, double *x, int degree, double ratio, bool zero) {
  MgLevel &L = h->mg.lev[l];
  int64_t ofs, nn;
  mg_owned(h, l, &ofs, &nn);
  const double lmax = L.lmax, lmin = lmax / ratio, theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  for (int k = 0; k < degree; ++k) {
    const bool first = zero && k == 0;
    if (!first) NLS_TRY(mg_apply_K(h, l, x, L.y));
    double a, c;
    if (k == 0) { a = 0.0; c = 1.0 / theta; }
    else { const double rn = 1.0 / (2.0 * sigma - rho); a = rn * rho; c = 2.0 * rn / delta; rho = rn; }
    const int mode = first ? 1 : (k == 0 ? 2 : 0);
    if (mg_nc(h, l) == 2) k_mg_cheb<2><<<mg_grid(h, nn), 256, 0, h->stream>>>(nn, L.dinv + 3 * ofs, b + ofs, first ? nullptr : L.y + ofs, L.d + ofs, x + ofs, a, c, mode);
    else k_mg_cheb<3><<<mg_grid(h, nn), 256, 0, h->stream>>>(nn, L.dinv + 3 * ofs, b + ofs, first ? nullptr : L.y + ofs, L.d + ofs, x + ofs, a, c, mode);
    MG_KCHECK(h);
  }
  return NLS_OK;
}

static int mg_vcycle(nls_context *h, int l, const double *b, double *x) {
  MgPlan &m = h->mg;
  MgLevel &L = m.lev[l];
  if (l == m.nlev - 1) {
    // the coarsest grid is "solved" by Chebyshev iteration: condition number ~ length^2, degree ~ length (6 nodes: the defaults)
    const double f = std::max(1.0, std::max(L.nx, std::max(L.ny, L.NZ)) / 6.0);
    return mg_cheb(h, l, b, x, (int)std::ceil(m.coarse_degree * f), m.coarse_ratio * f * f, true);
  }
  NLS_TRY(mg_cheb(h, l, b, x, m.degree, m.ratio, true));
  NLS_TRY(mg_apply_K(h, l, x, L.y));
  MgLevel &C = m.lev[l + 1];
  int64_t ofs, nn;
  mg_owned(h, l, &ofs, &nn);
  // residual of the owned rows, its ghost planes, restriction to the coarse planes this rank owns
  double *r = l == 0 ? m.rf : L.y;
  const int nc = mg_nc(h, l);
  k_mg_residual<<<mg_grid(h, nc * nn), 256, 0, h->stream>>>(nc * nn, b + ofs, L.y + ofs, r + ofs); MG_KCHECK(h);
  if (l == 0) { if (m.dist) NLS_TRY(nls_halo_exchange(h, r)); }
  else NLS_TRY(mg_exchange(h, l, r, 3));
  MgBox Cb = mg_box(C);
  const bool partial = l + 1 == 2 && m.dist;      // first replicated level: this rank's planes, then the sum over the ranks
  if (partial) { Cb.zown0 = C.rown0; Cb.zown1 = C.rown1; NLS_CUDA(h, cudaMemsetAsync(C.b, 0, sizeof(double) * 3 * (size_t)C.nn, h->stream)); }
  const int64_t ncown = (int64_t)C.nx * C.ny * (Cb.zown1 - Cb.zown0);
  if (ncown > 0) {
    k_mg_restrict<<<mg_grid(h, 3 * ncown), 256, 0, h->stream>>>(mg_box(L), l == 0 ? h->lplan.d_pbase : nullptr, m.z0, m.p0, nc, Cb, r, C.b);
    MG_KCHECK(h);
  }
  if (partial) NLS_TRY(mg_allsum(h, C.b, 3 * (size_t)C.nn));
  NLS_TRY(mg_vcycle(h, l + 1, C.b, C.x));
  NLS_TRY(mg_exchange(h, l + 1, C.x, 3));
  k_mg_prolong<<<mg_grid(h, nc * nn), 256, 0, h->stream>>>(mg_box(L), l == 0 ? h->lplan.d_pbase : nullptr, m.z0, m.p0, nc,
                                                          l == 0 ? h->d_con : nullptr, mg_box(C), C.x, x);
  MG_KCHECK(h);
  return mg_cheb(h, l, b, x, m.degree, m.ratio, false);
}

// Does this solve take the multigrid path? Single GPU: the local answer. Multi-GPU: the path is a sequence of collectives,
// so either every rank takes it or none does -- one allreduce per solve turns the ranks' votes (hierarchy available, option,
// size rule) into a common decision; a rank whose slab has no hierarchy vetoes it for all. COLLECTIVE when distributed.
int nls_mg_agreed(nls_context *h, bool *use) {
  MgPlan &m = h->mg;
  const int64_t nodes = h->distributed ? h->n_owned * (int64_t)h->nranks : h->nn;
  const bool mine = m.ok && (h->distributed || !(m.ghost_below || m.ghost_above)) &&
                    (h->opt_precond >= 1 || (h->opt_precond < 0 && nodes >= 10000));
  if (!h->distributed || h->nranks < 2) { *use = mine; return NLS_OK; }
  if (!m.d_vote) NLS_CUDA(h, cudaMalloc((void **)&m.d_vote, sizeof(double)));
  double v = mine ? 1.0 : 0.0;
  NLS_CUDA(h, cudaMemcpyAsync(m.d_vote, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (h->nccl->AllReduce(m.d_vote, m.d_vote, 1, NCCL_DYN_FLOAT64, NCCL_DYN_SUM, h->comm, h->stream) != 0)
    NLS_FAIL(h, NLS_ERR_NCCL, "multigrid: ncclAllReduce (agreement) failed");
  NLS_CUDA(h, cudaMemcpyAsync(&v, m.d_vote, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  *use = v == (double)h->nranks;
  return NLS_OK;
}

// z = M^-1 r (r masked; z's ghost entries are scratch)
int nls_mg_apply(nls_context *h, const double *r, double *z) {
  if (h->mg.mode == 1) {        // block-Jacobi alone
    if (h->dim == 2) k_mg_apply_dinv<2><<<mg_grid(h, h->n_owned), 256, 0, h->stream>>>(h->n_owned, h->mg.lev[0].dinv, r, z);
    else k_mg_apply_dinv<3><<<mg_grid(h, h->n_owned), 256, 0, h->stream>>>(h->n_owned, h->mg.lev[0].dinv, r, z);
    MG_KCHECK(h);
    return NLS_OK;
  }
  MgPlan &m = h->mg;
  if (!m.use_graph || m.dist || m.graph_off) return mg_vcycle(h, 0, r, z);
  const double key[6] = {(double)m.degree, m.ratio, (double)m.coarse_degree, m.coarse_ratio, (double)(m.fp32 && m.vals32), (double)m.nlev};
  if (!m.vgraph || m.vgraph_version != m.version || m.vgraph_r != r || m.vgraph_z != z || std::memcmp(key, m.vgraph_key, sizeof(key)) != 0) {
    mg_drop_graph(h);
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); m.graph_off = true; return mg_vcycle(h, 0, r, z); }
    const int64_t l0 = h->launches;
    const int rc = mg_vcycle(h, 0, r, z);
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    m.vgraph_launches = h->launches - l0;
    h->launches = l0;                       // nothing ran yet
    if (rc != NLS_OK || e != cudaSuccess || !g || cudaGraphInstantiate(&m.vgraph, g, 0) != cudaSuccess) {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      m.vgraph = nullptr; m.graph_off = true;
      return rc != NLS_OK ? rc : mg_vcycle(h, 0, r, z);
    }
    cudaGraphDestroy(g);
    m.vgraph_version = m.version; m.vgraph_r = r; m.vgraph_z = z;
    std::memcpy(m.vgraph_key, key, sizeof(key));
  }
  NLS_CUDA(h, cudaGraphLaunch(m.vgraph, h->stream));
  h->launches += m.vgraph_launches;
  return NLS_OK;
}

// ---- preconditioned conjugate gradients around it (host-driven: an iteration is tens of launches, the three reductions
//      per iteration cost nothing beside them) ----
int nls_pcg_solve_mg(nls_context *h, double rtol, int maxits, int *its, double *relres, int *status) {
  const int64_t n = h->nrows;
  static const bool timing = std::getenv("NLS_B200_MG_TIMING") != nullptr;
  auto now = [&]() { cudaStreamSynchronize(h->stream); return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = timing ? now() : 0.0;
  NLS_TRY(nls_mg_setup(h));
  const double t1 = timing ? now() : 0.0;
  double t_prec = 0.0, t_spmv = 0.0;
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
      const double ta = timing ? now() : 0.0;
      NLS_TRY(nls_halo_exchange(h, p));
      NLS_TRY(nls_launch_spmv(h, p, Ap, true, false));
      if (timing) t_spmv += now() - ta;
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
      const double tb = timing ? now() : 0.0;
      NLS_TRY(nls_mg_apply(h, r, z));
      if (timing) t_prec += now() - tb;
      double rzn;
      NLS_TRY(nls_dev_dot(h, n, r, z, &rzn));
      const double beta = rzn / rz;
      rz = rzn;
      k_mg_xpby<<<mg_grid(h, n), 256, 0, h->stream>>>(n, z, beta, p); MG_KCHECK(h);
    }
  }
  if (timing) fprintf(stderr, "[mg rank %d] set-up %.2f ms, %d iterations %.2f ms: V-cycles %.2f ms, Krylov products %.2f ms\n", h->rank, t1 - t0, it, now() - t1, t_prec, t_spmv);
  if (its) *its = it;
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  if (status) *status = st;
  return NLS_OK;
}
