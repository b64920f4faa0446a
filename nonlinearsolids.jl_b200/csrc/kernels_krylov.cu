// kernels_krylov.cu -- the linear-solve and vector steps of the Newton loop as sm_100a kernels.
//
// Replaces (paths relative to the reference source tree):
//   dofs(K) \ dofs(-R)            src/nonlinearsolvers/newton_fem.jl:158   -> Jacobi-PCG (SpMV + fused dot/axpy)
//   norm(dofs(fem, R))            src/nonlinearsolvers/newton_fem.jl:141,180 -> masked 2-norm reduction
//   norm(dU), U .+= dU            src/nonlinearsolvers/newton_fem.jl:168,173
//   dofs()/expand()               src/fem/fem.jl:93-108 -> constrained mask: identity rows/cols, x_c = 0
//
// The PCG loop is device-resident: scalars (alpha, beta, rz, convergence flag) live in a
// control block in HBM; every kernel reads the flag and becomes a no-op once converged, so
// the host may queue iterations ahead without synchronising. Reductions are deterministic:
// per-block partials, summed in block order by the last block to finish.
#include "nls_internal.h"
#include "nccl_dyn.h"

#define RED_BLOCKS_MAX 2048

struct RedScratch {
  double *partial;     // [3][RED_BLOCKS_MAX]
  unsigned *counter;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-reduce NV values, store per-block partials; the last block to arrive sums all
// partials in block order and writes out[0..NV). Returns true in thread 0 of that block.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partial, unsigned *counter, double *out) {
  __shared__ double sh[NV][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = warp_sum(v[i]);
    if (lane == 0) sh[i][wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = lane < nw ? sh[i][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) partial[i * RED_BLOCKS_MAX + blockIdx.x] = s;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // fixed-order sum: thread j accumulates partials j, j+blockDim, ... then a block tree
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      acc[i] += reinterpret_cast<volatile double *>(partial)[i * RED_BLOCKS_MAX + b];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = warp_sum(acc[i]);
    if (lane == 0) sh[i][wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = lane < nw ? sh[i][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) out[i] = s;
    }
  }
  if (threadIdx.x == 0) { *counter = 0; __threadfence(); }
  return threadIdx.x == 0;
}

// ------------------------------------------------------------------------------------
// CSR SpMV, T lanes per row (T | 32), grid-stride over row groups.
//   y[row] = mask ? 0 : sum_k vals[k] x[colind[k]]
// With DOT: also reduces sum_row x[row]*y[row] into ctl->red[0] (p.Ap of PCG).
// ------------------------------------------------------------------------------------
template <int T, bool DOT>
__global__ void __launch_bounds__(256)
k_spmv(int64_t nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
       const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
       const uint8_t *__restrict__ con, PcgCtl *ctl, RedScratch rs) {
  if (DOT) {
    if (ctl->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->rz = ctl->red[1];  // rz <- rz_new of the previous update
  }
  const int lane = threadIdx.x % T;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / T;
  double dot[1] = {0.0};
  for (int64_t base = 0; base < nrows; base += stride) {
    // every thread runs the same trip count, so whole warps stay converged for the shuffles
    const int64_t row = base + gid;
    const bool valid = row < nrows;
    double s = 0.0;
    if (valid) {
      const int64_t b = rowptr[row], e = rowptr[row + 1];
      for (int64_t k = b + lane; k < e; k += T) s += vals[k] * x[colind[k]];
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (valid && lane == 0) {
      if (con && con[row]) s = 0.0;
      y[row] = s;
      if (DOT) dot[0] += x[row] * s;
    }
  }
  if (DOT) grid_reduce<1>(dot, rs.partial, rs.counter, &ctl->red[0]);
}

// ------------------------------------------------------------------------------------
// Block-aware SpMV on the same CSR values: the scalar pattern is the dim x dim expansion of the
// node-level pattern (browptr/bcol), so the dim rows of a node share one column list. T lanes
// per node walk its blocks: one int32 block column per dim*dim values instead of dim*dim int32
// scalar columns (index traffic 4/dim^2 B per nonzero instead of 4 B), x gathered as a dim-vector.
// ------------------------------------------------------------------------------------
template <int DIM, int T, bool DOT, typename VT = double>
__global__ void __launch_bounds__(256)
k_bspmv(int64_t n_owned, const int64_t *__restrict__ browptr, const int32_t *__restrict__ bcol,
        const VT *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
        const uint8_t *__restrict__ con, PcgCtl *ctl, RedScratch rs) {
  if (DOT) {
    if (ctl->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->rz = ctl->red[1];
  }
  const int lane = threadIdx.x % T;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / T;
  double dot[1] = {0.0};
  for (int64_t base = 0; base < n_owned; base += stride) {
    const int64_t node = base + gid;
    const bool valid = node < n_owned;
    double acc[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) acc[i] = 0.0;
    if (valid) {
      const int64_t b0 = browptr[node];
      const int deg = (int)(browptr[node + 1] - b0);
      const VT *v0 = vals + (int64_t)DIM * DIM * b0;
      for (int j = lane; j < deg; j += T) {
        const int c = bcol[b0 + j];
        if (DIM == 2 && sizeof(VT) == 8) {
          const double2 xv = reinterpret_cast<const double2 *>(x)[c];
          const double2 a0 = *reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(v0) + 2 * j);
          const double2 a1 = *reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(v0) + 2 * deg + 2 * j);
          acc[0] = fma(a0.x, xv.x, fma(a0.y, xv.y, acc[0]));
          acc[1] = fma(a1.x, xv.x, fma(a1.y, xv.y, acc[1]));
        } else {
          double xv[DIM];
#pragma unroll
          for (int k = 0; k < DIM; ++k) xv[k] = x[(int64_t)DIM * c + k];
#pragma unroll
          for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int k = 0; k < DIM; ++k) acc[i] = fma((double)v0[(int64_t)i * DIM * deg + DIM * j + k], xv[k], acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (valid && lane == 0) {
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        const int64_t row = DIM * node + i;
        double sv = acc[i];
        if (con && con[row]) sv = 0.0;
        y[row] = sv;
        if (DOT) dot[0] += x[row] * sv;
      }
    }
  }
  if (DOT) grid_reduce<1>(dot, rs.partial, rs.counter, &ctl->red[0]);
}

// ------------------------------------------------------------------------------------
// Symmetric SpMV for the Krylov solvers (PCG / MINRES assume K = K^T): only the upper blocks of every block row
// (column node >= row node, local numbering) are kept, packed contiguously per row as dim x dim row-major blocks.
// Row a multiplies its own upper blocks; for its lower neighbours b < a it reads block (b, a) from row b's packed
// list through a precomputed slot and applies it transposed. Each stored block is fetched from HBM once by its
// own row and a second time, shortly afterwards, out of L2 by the mirrored row (the window between the two is one
// node plane of upper blocks: 23 MB at C3), so the DRAM traffic per product is halved: 8 -> ~4.4 B per nonzero.
// No atomics: every row is still produced by one thread group, in a fixed order.
// MEASURED SLOWER than the full block-aware product (C3 PCG iteration 1816 vs 1386 us, C2 71 vs 67.5 us): the
// mirrored 72-byte blocks come out of L2 as scattered sector requests behind a dependent slot lookup, which costs
// more than streaming twice the bytes from HBM at 90 % of peak. Kept behind nls_set_option("sym_spmv", 1), off by
// default.
// ------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256)
k_sym_pack(int64_t n_owned, const int64_t *__restrict__ browptr, const int32_t *__restrict__ dpos,
           const int64_t *__restrict__ symptr, const double *__restrict__ vals, double *__restrict__ sv) {
  // one warp per node: the packed blocks of a row are contiguous, so the stores coalesce
  const int64_t node = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= n_owned) return;
  const int64_t b0 = browptr[node];
  const int deg = (int)(browptr[node + 1] - b0), dp = dpos[node], nup = deg - dp;
  const double *v0 = vals + (int64_t)DIM * DIM * b0;
  double *o = sv + (int64_t)DIM * DIM * symptr[node];
  for (int t = threadIdx.x & 31; t < nup * DIM * DIM; t += 32) {
    const int j = t / (DIM * DIM), ik = t - j * DIM * DIM, i = ik / DIM, k = ik - i * DIM;
    o[t] = v0[(int64_t)i * DIM * deg + DIM * (dp + j) + k];
  }
}

// slot of block (a, b), a < b, in the packed storage, seen from row b (its lower entry jb): one thread per block
__global__ void __launch_bounds__(256)
k_sym_index(int64_t n_owned, const int64_t *__restrict__ browptr, const int32_t *__restrict__ bcol, const int32_t *__restrict__ dpos,
            const int64_t *__restrict__ symptr, uint32_t *__restrict__ lowslot) {
  const int64_t node = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= n_owned) return;
  const int64_t b0 = browptr[node];
  const int dp = dpos[node];
  const int64_t lp = b0 - symptr[node];
  for (int jb = threadIdx.x & 31; jb < dp; jb += 32) {
    const int a = bcol[b0 + jb];
    int64_t lo = browptr[a], hi = browptr[a + 1] - 1;
    const int64_t a0 = lo;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (bcol[mid] < (int32_t)node) lo = mid + 1; else hi = mid; }
    lowslot[lp + jb] = (uint32_t)(symptr[a] + (lo - a0) - dpos[a]);
  }
}

// T lanes per node; a lane takes whole blocks: first the lower neighbours (block read at its slot, applied
// transposed), then the node's own packed upper blocks (contiguous).
template <int DIM, int T, bool DOT>
__global__ void __launch_bounds__(256)
k_symspmv(int64_t n_owned, const int64_t *__restrict__ browptr, const int32_t *__restrict__ bcol, const int32_t *__restrict__ dpos,
          const int64_t *__restrict__ symptr, const uint32_t *__restrict__ lowslot, const double *__restrict__ sv,
          const double *__restrict__ x, double *__restrict__ y, const uint8_t *__restrict__ con, PcgCtl *ctl, RedScratch rs) {
  if (DOT) {
    if (ctl->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->rz = ctl->red[1];
  }
  constexpr int D2 = DIM * DIM;
  const int lane = threadIdx.x % T;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / T;
  double dot[1] = {0.0};
  for (int64_t base = 0; base < n_owned; base += stride) {
    const int64_t node = base + gid;
    const bool valid = node < n_owned;
    double acc[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) acc[i] = 0.0;
    if (valid) {
      const int64_t b0 = browptr[node], sp = symptr[node];
      const int deg = (int)(browptr[node + 1] - b0), dp = dpos[node];
      const uint32_t *ls = lowslot + (b0 - sp);
      for (int j = lane; j < dp; j += T) {                 // lower neighbours: K_(c,node)^T x_c
        const int c = bcol[b0 + j];
        const double *kp = sv + (int64_t)D2 * (int64_t)ls[j];
        double xv[DIM], K[D2];
#pragma unroll
        for (int k = 0; k < DIM; ++k) xv[k] = x[(int64_t)DIM * c + k];
#pragma unroll
        for (int t = 0; t < D2; ++t) K[t] = kp[t];
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int k = 0; k < DIM; ++k) acc[i] = fma(K[k * DIM + i], xv[k], acc[i]);
      }
      const double *up = sv + (int64_t)D2 * sp;
      for (int j = lane; j < deg - dp; j += T) {           // own upper blocks
        const int c = bcol[b0 + dp + j];
        const double *kp = up + D2 * j;
        double xv[DIM], K[D2];
#pragma unroll
        for (int k = 0; k < DIM; ++k) xv[k] = x[(int64_t)DIM * c + k];
#pragma unroll
        for (int t = 0; t < D2; ++t) K[t] = kp[t];
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int k = 0; k < DIM; ++k) acc[i] = fma(K[i * DIM + k], xv[k], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (valid && lane == 0) {
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        const int64_t row = DIM * node + i;
        double s = acc[i];
        if (con && con[row]) s = 0.0;
        y[row] = s;
        if (DOT) dot[0] += x[row] * s;
      }
    }
  }
  if (DOT) grid_reduce<1>(dot, rs.partial, rs.counter, &ctl->red[0]);
}

// diagonal -> Jacobi preconditioner; constrained rows get 0 (their residual is 0 anyway)
__global__ void k_extract_dinv(int64_t nrows, const int64_t *rowptr, const int32_t *diag, const double *vals,
                               const uint8_t *con, double *dinv) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  dinv[r] = con[r] ? 0.0 : 1.0 / vals[rowptr[r] + diag[r]];
}

// PCG start: x = 0, r = masked b, p = z = dinv r; reduces rz and bb.
__global__ void __launch_bounds__(256)
k_pcg_init(int64_t nrows, const double *__restrict__ b, const uint8_t *__restrict__ con,
           const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
           double *__restrict__ p, PcgCtl *ctl, RedScratch rs) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = con[i] ? 0.0 : b[i];
    const double zi = dinv[i] * ri;
    x[i] = 0.0; r[i] = ri; p[i] = zi;
    v[0] += ri * zi; v[1] += ri * ri;
  }
  grid_reduce<2>(v, rs.partial, rs.counter, &ctl->red[1]);  // red[1] = rz, red[2] = bb
}
__global__ void k_pcg_after_init(PcgCtl *ctl, double tol2, int maxits) {
  ctl->bb = ctl->red[2]; ctl->rr = ctl->red[2]; ctl->rz = ctl->red[1];
  ctl->tol2 = tol2; ctl->iters = 0; ctl->maxits = maxits;
  ctl->done = (ctl->red[2] == 0.0) ? 1 : 0;
}

// x += alpha p; r -= alpha Ap; reduces rz_new = r.(dinv r) and rr = r.r
__global__ void __launch_bounds__(256)
k_pcg_update(int64_t nrows, const double *__restrict__ p, const double *__restrict__ Ap,
             const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
             PcgCtl *ctl, RedScratch rs) {
  if (ctl->done) return;
  const double pAp = ctl->red[0];
  if (!(pAp > 0.0)) {  // breakdown (not SPD, NaN): every thread sees the same value
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->done = 3;
    return;
  }
  const double alpha = ctl->rz / pAp;
  double v[2] = {0.0, 0.0};
  // two rows per thread and load (16-byte accesses); the odd last row, if any, goes to one thread
  const int64_t n2 = nrows >> 1;
  const double2 *p2 = reinterpret_cast<const double2 *>(p), *Ap2 = reinterpret_cast<const double2 *>(Ap), *d2 = reinterpret_cast<const double2 *>(dinv);
  double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 pp = p2[i], ap = Ap2[i], dd = d2[i];
    double2 xx = x2[i], rr = r2[i];
    xx.x += alpha * pp.x; xx.y += alpha * pp.y;
    rr.x -= alpha * ap.x; rr.y -= alpha * ap.y;
    x2[i] = xx; r2[i] = rr;
    v[0] += rr.x * (dd.x * rr.x); v[1] += rr.x * rr.x;
    v[0] += rr.y * (dd.y * rr.y); v[1] += rr.y * rr.y;
  }
  if ((nrows & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = nrows - 1;
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    v[0] += ri * (dinv[i] * ri); v[1] += ri * ri;
  }
  if (grid_reduce<2>(v, rs.partial, rs.counter, &ctl->red[1])) ctl->iters += 1;
}

// convergence test, then p = dinv r + beta p
__global__ void __launch_bounds__(256)
k_pcg_pupdate(int64_t nrows, const double *__restrict__ r, const double *__restrict__ dinv,
              double *__restrict__ p, PcgCtl *ctl) {
  if (ctl->done) return;
  const double rr = ctl->red[2];
  if (!(rr == rr)) { if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->done = 3; ctl->rr = rr; } return; }
  if (rr <= ctl->tol2 * ctl->bb) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->done = 1; ctl->rr = rr; }
    return;
  }
  const double beta = ctl->red[1] / ctl->rz;
  const int64_t n2 = nrows >> 1;
  const double2 *r2 = reinterpret_cast<const double2 *>(r), *d2 = reinterpret_cast<const double2 *>(dinv);
  double2 *p2 = reinterpret_cast<double2 *>(p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 rr = r2[i], dd = d2[i];
    double2 pp = p2[i];
    pp.x = dd.x * rr.x + beta * pp.x; pp.y = dd.y * rr.y + beta * pp.y;
    p2[i] = pp;
  }
  if ((nrows & 1) && blockIdx.x == 0 && threadIdx.x == 0) { const int64_t i = nrows - 1; p[i] = dinv[i] * r[i] + beta * p[i]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->rr = rr;
}

// generic masked dot: out = sum_{i<n, !con[i]} x_i y_i
__global__ void __launch_bounds__(256)
k_dot(int64_t n, const double *__restrict__ x, const double *__restrict__ y, const uint8_t *__restrict__ con,
      double *out, RedScratch rs) {
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!con || !con[i]) v[0] += x[i] * y[i];
  grid_reduce<1>(v, rs.partial, rs.counter, out);
}

__global__ void __launch_bounds__(256)
k_axpy(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void __launch_bounds__(256)
k_negate(int64_t n, const double *__restrict__ x, double *__restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = -x[i];
}

// halo pack/unpack: node lists -> dim doubles per node
__global__ void k_pack(int64_t nnodes, int dim, const int32_t *nodes, const double *v, double *buf) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nnodes * dim) return;
  buf[t] = v[(int64_t)nodes[t / dim] * dim + t % dim];
}
__global__ void k_unpack(int64_t nnodes, int dim, const int32_t *nodes, const double *buf, double *v) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nnodes * dim) return;
  v[(int64_t)nodes[t / dim] * dim + t % dim] = buf[t];
}

// ------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------
#define NLS_KCHECK(h) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string("kernel launch: ") + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

static inline RedScratch red_of(nls_context *h) {
  RedScratch rs;
  rs.partial = h->d_partial;
  rs.counter = reinterpret_cast<unsigned *>(h->d_partial + 3 * RED_BLOCKS_MAX);
  return rs;
}
static inline unsigned vec_grid(nls_context *h, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (g > cap) g = cap;
  if (g > RED_BLOCKS_MAX) g = RED_BLOCKS_MAX;
  if (g < 1) g = 1;
  return (unsigned)g;
}

template <int T>
static int spmv_t(nls_context *h, const double *x, double *y, const uint8_t *con, bool dot) {
  int64_t g = (h->nrows * T + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (g > cap) g = cap;
  if (g > RED_BLOCKS_MAX) g = RED_BLOCKS_MAX;
  if (g < 1) g = 1;
  if (dot) k_spmv<T, true><<<(unsigned)g, 256, 0, h->stream>>>(h->nrows, h->d_rowptr, h->d_colind, h->d_vals, x, y, con, h->d_ctl, red_of(h));
  else     k_spmv<T, false><<<(unsigned)g, 256, 0, h->stream>>>(h->nrows, h->d_rowptr, h->d_colind, h->d_vals, x, y, con, h->d_ctl, red_of(h));
  NLS_KCHECK(h);
  return NLS_OK;
}

// Grid cap of the grid-stride products: exactly the blocks that are resident at once. With the former cap of 8 blocks per SM
// and 40 registers per thread only 6 fit, and the kernel ran 1.33 waves -- the last third of the blocks at a third of the
// memory parallelism of a latency-bound kernel (profiles/r2_ncu_spmv_f32_c3.txt: long_scoreboard 52.9, 57.8 % warps active).
template <typename Kern>
static int64_t resident_blocks(nls_context *h, Kern kernel) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 256, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 6; }
  return (int64_t)h->sm_count * nb;
}

template <int DIM, int T>
static int bspmv_t(nls_context *h, const double *x, double *y, const uint8_t *con, bool dot) {
  int64_t g = (h->n_owned * T + 255) / 256;
  static const int64_t cap_dot = resident_blocks(h, k_bspmv<DIM, T, true>), cap_plain = resident_blocks(h, k_bspmv<DIM, T, false>);
  const int64_t cap = dot ? cap_dot : cap_plain;
  if (g > cap) g = cap;
  if (g > RED_BLOCKS_MAX) g = RED_BLOCKS_MAX;
  if (g < 1) g = 1;
  if (dot) k_bspmv<DIM, T, true><<<(unsigned)g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_bcol, h->d_vals, x, y, con, h->d_ctl, red_of(h));
  else     k_bspmv<DIM, T, false><<<(unsigned)g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_bcol, h->d_vals, x, y, con, h->d_ctl, red_of(h));
  NLS_KCHECK(h);
  return NLS_OK;
}

template <int DIM>
static int symspmv_t(nls_context *h, const double *x, double *y, const uint8_t *con, bool dot) {
  constexpr int T = 4;
  int64_t g = (h->n_owned * T + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (g > cap) g = cap;
  if (g > RED_BLOCKS_MAX) g = RED_BLOCKS_MAX;
  if (g < 1) g = 1;
  if (dot) k_symspmv<DIM, T, true><<<(unsigned)g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_bcol, h->d_dpos, h->d_symptr, h->d_lowslot, h->d_symvals, x, y, con, h->d_ctl, red_of(h));
  else     k_symspmv<DIM, T, false><<<(unsigned)g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_bcol, h->d_dpos, h->d_symptr, h->d_lowslot, h->d_symvals, x, y, con, h->d_ctl, red_of(h));
  NLS_KCHECK(h);
  return NLS_OK;
}

// Packs the upper blocks of the current CSR values for the symmetric product (once per set of values). Returns with
// h->sym_active = false when the option is off or the storage does not fit; the solvers then use the full product.
int nls_sym_prepare(nls_context *h) {
  h->sym_active = false;
  if (!h->opt_sym_spmv || h->dim < 2 || !h->have_sym_index || h->n_owned == 0) return NLS_OK;
  if (!h->d_symvals) {
    if (cudaMalloc((void **)&h->d_symvals, sizeof(double) * (size_t)std::max<int64_t>(1, h->nsym) * h->dim * h->dim) != cudaSuccess) {
      cudaGetLastError(); h->d_symvals = nullptr; h->opt_sym_spmv = 0; return NLS_OK;
    }
    h->sym_version = -1;
  }
  if (h->sym_version != h->vals_version) {
    const unsigned g = (unsigned)((h->n_owned + 7) / 8);
    if (h->dim == 2) k_sym_pack<2><<<g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_dpos, h->d_symptr, h->d_vals, h->d_symvals);
    else             k_sym_pack<3><<<g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_dpos, h->d_symptr, h->d_vals, h->d_symvals);
    NLS_KCHECK(h);
    h->sym_version = h->vals_version;
  }
  h->sym_active = true;
  return NLS_OK;
}

int nls_sym_build_index(nls_context *h) {
  // host: position of the diagonal block in every block row, prefix sums of the upper counts
  h->have_sym_index = false;
  if (h->dim < 2 || h->n_owned == 0) return NLS_OK;
  const int64_t no = h->n_owned;
  std::vector<int32_t> dpos((size_t)no);
  std::vector<int64_t> symptr((size_t)no + 1, 0);
  for (int64_t a = 0; a < no; ++a) {
    const int32_t *lo = h->h_bcol.data() + h->h_browptr[a], *hi = h->h_bcol.data() + h->h_browptr[a + 1];
    const int32_t *d = std::lower_bound(lo, hi, (int32_t)a);
    dpos[a] = (int32_t)(d - lo);
    symptr[a + 1] = symptr[a] + (hi - d);
  }
  h->nsym = symptr[no];
  const int64_t nlow = h->nnzb - h->nsym;
  if (h->nsym >= ((int64_t)1 << 32)) return NLS_OK;
  if (h->d_dpos) cudaFree(h->d_dpos);
  if (h->d_symptr) cudaFree(h->d_symptr);
  if (h->d_lowslot) cudaFree(h->d_lowslot);
  if (h->d_symvals) { cudaFree(h->d_symvals); h->d_symvals = nullptr; }
  h->d_dpos = nullptr; h->d_symptr = nullptr; h->d_lowslot = nullptr;
  NLS_CUDA(h, cudaMalloc((void **)&h->d_dpos, sizeof(int32_t) * no));
  NLS_CUDA(h, cudaMalloc((void **)&h->d_symptr, sizeof(int64_t) * (no + 1)));
  NLS_CUDA(h, cudaMalloc((void **)&h->d_lowslot, sizeof(uint32_t) * (size_t)std::max<int64_t>(1, nlow)));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_dpos, dpos.data(), sizeof(int32_t) * no, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(h->d_symptr, symptr.data(), sizeof(int64_t) * (no + 1), cudaMemcpyHostToDevice, h->stream));
  k_sym_index<<<(unsigned)((no + 7) / 8), 256, 0, h->stream>>>(no, h->d_browptr, h->d_bcol, h->d_dpos, h->d_symptr, h->d_lowslot);
  NLS_KCHECK(h);
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_sym_index = true;
  return NLS_OK;
}

int nls_launch_spmv(nls_context *h, const double *x, double *y, bool masked, bool with_dot) {
  if (h->nrows == 0) return NLS_OK;
  const uint8_t *con = masked ? h->d_con : nullptr;
  if (h->sym_active && h->dim >= 2) return h->dim == 2 ? symspmv_t<2>(h, x, y, con, with_dot) : symspmv_t<3>(h, x, y, con, with_dot);
  if (h->opt_spmv == 1 && h->dim >= 2) {   // block-aware traversal of the same CSR values
    const int t = h->opt_spmv_lanes;
    if (h->dim == 2) {
      if (t == 2) return bspmv_t<2, 2>(h, x, y, con, with_dot);
      if (t == 8) return bspmv_t<2, 8>(h, x, y, con, with_dot);
      return bspmv_t<2, 4>(h, x, y, con, with_dot);
    }
    if (t == 8) return bspmv_t<3, 8>(h, x, y, con, with_dot);
    if (t == 16) return bspmv_t<3, 16>(h, x, y, con, with_dot);
    return bspmv_t<3, 4>(h, x, y, con, with_dot);   // measured best at C3 (profiles/r1d_spmv_variants.txt)
  }
  const double avg = (double)h->nnz / (double)h->nrows;
  if (avg <= 3.0) return spmv_t<1>(h, x, y, con, with_dot);
  if (avg <= 6.0) return spmv_t<2>(h, x, y, con, with_dot);
  if (avg <= 12.0) return spmv_t<4>(h, x, y, con, with_dot);
  if (avg <= 40.0) return spmv_t<8>(h, x, y, con, with_dot);
  return spmv_t<16>(h, x, y, con, with_dot);
}

// y = K x with the single-precision copy of the CSR values (the smoother of the multigrid preconditioner: half the HBM
// traffic of the product; accumulation and vectors stay FP64)
__global__ void __launch_bounds__(256) k_vals_to_f32(int64_t n, const double *__restrict__ v, float *__restrict__ o) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) o[i] = (float)v[i];
}
int nls_launch_vals_to_f32(nls_context *h, float *out) {
  k_vals_to_f32<<<vec_grid(h, h->nnz), 256, 0, h->stream>>>(h->nnz, h->d_vals, out);
  NLS_KCHECK(h);
  return NLS_OK;
}
int nls_launch_spmv_f32(nls_context *h, const float *vals32, const double *x, double *y) {
  if (h->dim < 2 || h->nrows == 0) NLS_FAIL(h, NLS_ERR_STATE, "single-precision product: 2-D and 3-D models only");
  int64_t g = (h->n_owned * 4 + 255) / 256;
  static const int64_t cap2 = resident_blocks(h, k_bspmv<2, 4, false, float>), cap3 = resident_blocks(h, k_bspmv<3, 4, false, float>);
  const int64_t cap = h->dim == 2 ? cap2 : cap3;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  if (h->dim == 2) k_bspmv<2, 4, false, float><<<(unsigned)g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_bcol, vals32, x, y, h->d_con, h->d_ctl, red_of(h));
  else k_bspmv<3, 4, false, float><<<(unsigned)g, 256, 0, h->stream>>>(h->n_owned, h->d_browptr, h->d_bcol, vals32, x, y, h->d_con, h->d_ctl, red_of(h));
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_extract_dinv(nls_context *h) {
  if (h->nrows == 0) return NLS_OK;
  k_extract_dinv<<<(unsigned)((h->nrows + 255) / 256), 256, 0, h->stream>>>(h->nrows, h->d_rowptr, h->d_diag, h->d_vals, h->d_con, h->d_dinv);
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_dev_dot(nls_context *h, int64_t n, const double *x, const double *y, double *out) {
  k_dot<<<vec_grid(h, n), 256, 0, h->stream>>>(n, x, y, nullptr, h->d_scal, red_of(h));
  NLS_KCHECK(h);
  if (h->distributed) { int rc = nls_allreduce_sum(h, h->d_scal, 1); if (rc) return rc; }
  NLS_CUDA(h, cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  *out = h->h_scal[0];
  return nls_p2p_failed(h);       // a timed-out peer exchange leaves stale ghosts / partial sums: never report NLS_OK over it
}

int nls_dev_dot_local(nls_context *h, int64_t n, const double *x, const double *y, double *out) {
  k_dot<<<vec_grid(h, n), 256, 0, h->stream>>>(n, x, y, nullptr, h->d_scal, red_of(h));
  NLS_KCHECK(h);
  NLS_CUDA(h, cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  *out = h->h_scal[0];
  return NLS_OK;
}

// sqrt(sum v_i^2) over owned rows, optionally skipping constrained DoFs -- norm(dofs(fem, R))
double nls_dev_masked_norm2(nls_context *h, const double *v, bool masked, int *rc) {
  *rc = NLS_OK;
  k_dot<<<vec_grid(h, h->nrows), 256, 0, h->stream>>>(h->nrows, v, v, masked ? h->d_con : nullptr, h->d_scal, red_of(h));
  h->launches++;
  if (h->distributed) { *rc = nls_allreduce_sum(h, h->d_scal, 1); if (*rc) return 0.0; }
  cudaError_t e = cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { h->err = std::string("norm: ") + cudaGetErrorString(e); *rc = NLS_ERR_CUDA; return 0.0; }
  if ((*rc = nls_p2p_failed(h)) != NLS_OK) return 0.0;
  return sqrt(h->h_scal[0]);
}

int nls_launch_axpy(nls_context *h, int64_t n, double a, const double *x, double *y) {
  if (n == 0) return NLS_OK;
  k_axpy<<<vec_grid(h, n), 256, 0, h->stream>>>(n, a, x, y);
  NLS_KCHECK(h);
  return NLS_OK;
}

// ---- communication: neighbour halo (ghost-DoF) exchange + scalar allreduce over NCCL ----
int nls_halo_exchange(nls_context *h, double *v) {
  if (!h->distributed || h->nneigh == 0) return NLS_OK;
  if (h->p2p.on) return nls_p2p_halo_exchange(h, v);
  const int dim = h->dim;
  cudaStream_t st = h->halo_stream ? h->halo_stream : h->stream;
  const int64_t nsend = h->send_ptr[h->nneigh], nrecv = h->recv_ptr[h->nneigh];
  if (nsend > 0) {
    k_pack<<<(unsigned)((nsend * dim + 255) / 256), 256, 0, st>>>(nsend, dim, h->d_send_nodes, v, h->d_sendbuf);
    NLS_KCHECK(h);
  }
  NcclApi *nc = h->nccl;
  if (nc->GroupStart() != 0) NLS_FAIL(h, NLS_ERR_NCCL, "ncclGroupStart failed");
  for (int q = 0; q < h->nneigh; ++q) {
    const int64_t s0 = h->send_ptr[q] * dim, s1 = h->send_ptr[q + 1] * dim;
    const int64_t r0 = h->recv_ptr[q] * dim, r1 = h->recv_ptr[q + 1] * dim;
    int e1 = 0, e2 = 0;
    if (s1 > s0) e1 = nc->Send(h->d_sendbuf + s0, (size_t)(s1 - s0), NCCL_DYN_FLOAT64, h->neigh[q], h->comm, st);
    if (r1 > r0) e2 = nc->Recv(h->d_recvbuf + r0, (size_t)(r1 - r0), NCCL_DYN_FLOAT64, h->neigh[q], h->comm, st);
    if (e1 || e2) { nc->GroupEnd(); NLS_FAIL(h, NLS_ERR_NCCL, std::string("ncclSend/Recv: ") + nc->GetErrorString(e1 ? e1 : e2)); }
  }
  int ge = nc->GroupEnd();
  if (ge != 0) NLS_FAIL(h, NLS_ERR_NCCL, std::string("ncclGroupEnd: ") + nc->GetErrorString(ge));
  if (nrecv > 0) {
    k_unpack<<<(unsigned)((nrecv * dim + 255) / 256), 256, 0, st>>>(nrecv, dim, h->d_recv_nodes, h->d_recvbuf, v);
    NLS_KCHECK(h);
  }
  return NLS_OK;
}

int nls_allreduce_sum(nls_context *h, double *dptr, int count) {
  if (!h->distributed) return NLS_OK;
  if (h->p2p.on) return nls_p2p_allreduce(h, dptr, count);
  int e = h->nccl->AllReduce(dptr, dptr, (size_t)count, NCCL_DYN_FLOAT64, NCCL_DYN_SUM, h->comm, h->stream);
  if (e != 0) NLS_FAIL(h, NLS_ERR_NCCL, std::string("ncclAllReduce: ") + h->nccl->GetErrorString(e));
  return NLS_OK;
}

// ------------------------------------------------------------------------------------
// Whole Jacobi-PCG solve in ONE thread block, for small systems (the reference's own problem sizes: a few hundred
// DoFs). The multi-kernel loop below costs three dependent launches per iteration (~14 us even graph-replayed);
// here an iteration is one pass over the block's rows and three block-wide reductions. Same arithmetic, masks
// and convergence tests as k_pcg_init / k_spmv<DOT> / k_pcg_update / k_pcg_pupdate; only the summation order of the
// reductions differs. Rows i = tid, tid + 1024, ... (at most PCG_SMALL_RPT per thread; A p stays in registers).
// ------------------------------------------------------------------------------------
#define PCG_SMALL_THREADS 1024
#define PCG_SMALL_RPT 8

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*sh)[32]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = warp_sum(v[i]);
    if (lane == 0) sh[i][wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = lane < (int)(blockDim.x >> 5) ? sh[i][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) sh[i][0] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sh[i][0];
  __syncthreads();
}

__global__ void __launch_bounds__(PCG_SMALL_THREADS)
k_pcg_small(int nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind, const int32_t *__restrict__ diag,
            const double *__restrict__ vals, const double *__restrict__ b, const uint8_t *__restrict__ con,
            double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r, double *p, PcgCtl *ctl,
            double tol2, int maxits) {
  __shared__ double sh[2][32];
  const int tid = threadIdx.x, nthr = blockDim.x;   // nthr: a multiple of 32, nthr * PCG_SMALL_RPT >= nrows
  double v[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < PCG_SMALL_RPT; ++k) {
    const int i = tid + k * nthr;
    if (i < nrows) {
      const double di = con[i] ? 0.0 : 1.0 / vals[rowptr[i] + diag[i]];
      const double ri = con[i] ? 0.0 : b[i];
      const double zi = di * ri;
      dinv[i] = di; x[i] = 0.0; r[i] = ri; p[i] = zi;
      v[0] += ri * zi; v[1] += ri * ri;
    }
  }
  block_sum<2>(v, sh);
  double rz = v[0], rr = v[1];
  const double bb = v[1];
  int it = 0, done = (bb == 0.0) ? 1 : 0;
  while (!done && it < maxits) {
    double Ap[PCG_SMALL_RPT];
    double d1[1] = {0.0};
#pragma unroll
    for (int k = 0; k < PCG_SMALL_RPT; ++k) {
      const int i = tid + k * nthr;
      Ap[k] = 0.0;
      if (i < nrows) {
        double s = 0.0;
        for (int64_t q = rowptr[i], e = rowptr[i + 1]; q < e; ++q) s += vals[q] * p[colind[q]];
        if (con[i]) s = 0.0;
        Ap[k] = s;
        d1[0] += p[i] * s;
      }
    }
    block_sum<1>(d1, sh);
    const double pAp = d1[0];
    if (!(pAp > 0.0)) { done = 3; break; }
    const double alpha = rz / pAp;
    v[0] = 0.0; v[1] = 0.0;
#pragma unroll
    for (int k = 0; k < PCG_SMALL_RPT; ++k) {
      const int i = tid + k * nthr;
      if (i < nrows) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[k];
        r[i] = ri;
        v[0] += ri * (dinv[i] * ri); v[1] += ri * ri;
      }
    }
    block_sum<2>(v, sh);              // also the barrier after which every thread is done reading p
    ++it;
    rr = v[1];
    if (!(rr == rr)) { done = 3; break; }
    if (rr <= tol2 * bb) { done = 1; break; }
    const double beta = v[0] / rz;
    rz = v[0];
#pragma unroll
    for (int k = 0; k < PCG_SMALL_RPT; ++k) {
      const int i = tid + k * nthr;
      if (i < nrows) p[i] = dinv[i] * r[i] + beta * p[i];
    }
    __syncthreads();                  // the new p is complete before the next matrix-vector product reads it
  }
  if (tid == 0) {
    ctl->iters = it; ctl->done = done; ctl->rr = rr; ctl->bb = bb; ctl->rz = rz; ctl->tol2 = tol2; ctl->maxits = maxits;
  }
}

// ------------------------------------------------------------------------------------
// Jacobi-PCG driver on the handle's current CSR values. rhs = h->d_b (masked inside),
// solution -> h->d_dU (zero on constrained DoFs = expand(), fem.jl:98-102).
//   fixed_iters > 0: run exactly that many iterations without convergence checks (bench).
//   status: 0 converged, 1 maxits, 2 breakdown
// ------------------------------------------------------------------------------------
int nls_pcg_solve(nls_context *h, double rtol, int maxits, int fixed_iters, int *its, double *relres, int *status) {
  int rc;
  const int64_t n = h->nrows;
  RedScratch rs = red_of(h);
  // grid-topology 3-D meshes: multigrid (or block-Jacobi) preconditioner, kernels_mg.cu
  if (fixed_iters == 0) {
    bool use = false;
    if ((rc = nls_mg_agreed(h, &use))) return rc;       // collective when distributed: all ranks take the same path
    if (use) {
      h->mg.mode = h->opt_precond == 1 ? 1 : 2;
      return nls_pcg_solve_mg(h, rtol, maxits, its, relres, status);
    }
  }
  if (!h->distributed && h->opt_pcg_small && fixed_iters == 0 && n > 0 && n <= (int64_t)PCG_SMALL_THREADS * PCG_SMALL_RPT &&
      h->nnz <= (int64_t)1 << 21) {
    const int nthr = (int)std::min<int64_t>(PCG_SMALL_THREADS, std::max<int64_t>(64, ((n + PCG_SMALL_RPT - 1) / PCG_SMALL_RPT + 31) / 32 * 32 * 2));
    k_pcg_small<<<1, std::min(nthr, PCG_SMALL_THREADS), 0, h->stream>>>((int)n, h->d_rowptr, h->d_colind, h->d_diag, h->d_vals, h->d_b, h->d_con,
                                                       h->d_dinv, h->d_dU, h->d_r, h->d_p, h->d_ctl, rtol * rtol, maxits);
    NLS_KCHECK(h);
    NLS_CUDA(h, cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(PcgCtl), cudaMemcpyDeviceToHost, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    const PcgCtl &c = *h->h_ctl;
    if (its) *its = c.iters;
    if (relres) *relres = c.bb > 0.0 ? sqrt(c.rr / c.bb) : 0.0;
    if (status) *status = (c.done == 1) ? 0 : (c.done == 3 ? 2 : 1);
    return NLS_OK;
  }
  if ((rc = nls_launch_extract_dinv(h))) return rc;
  if ((rc = nls_sym_prepare(h))) return rc;       // the solver's products read the packed upper blocks (K = K^T)
  struct SymScope { nls_context *h; ~SymScope() { h->sym_active = false; } } sym_scope{h};
  k_pcg_init<<<vec_grid(h, n), 256, 0, h->stream>>>(n, h->d_b, h->d_con, h->d_dinv, h->d_dU, h->d_r, h->d_p, h->d_ctl, rs);
  NLS_KCHECK(h);
  if ((rc = nls_allreduce_sum(h, &h->d_ctl->red[1], 2))) return rc;
  const double tol2 = fixed_iters > 0 ? -1.0 : rtol * rtol;
  k_pcg_after_init<<<1, 1, 0, h->stream>>>(h->d_ctl, tol2, maxits);
  NLS_KCHECK(h);
  if ((rc = nls_halo_exchange(h, h->d_p))) return rc;
  const int total = fixed_iters > 0 ? fixed_iters : maxits;
  const int chunk = fixed_iters > 0 ? fixed_iters : 16;
  int launched = 0;
  h->h_ctl->done = 0; h->h_ctl->iters = 0;
  // One PCG iteration = 3 kernels (+ 3 exchanges when distributed). Single GPU: a full chunk of 16 iterations is
  // captured once into a CUDA graph and replayed (the kernels take fixed pointers and read alpha/beta/done from the
  // device control block), which removes the per-launch host cost that dominates small models.
  auto iterate = [&](int todo) -> int {
    for (int i = 0; i < todo; ++i) {
      if ((rc = nls_launch_spmv(h, h->d_p, h->d_Ap, true, true))) return rc;
      if ((rc = nls_allreduce_sum(h, &h->d_ctl->red[0], 1))) return rc;
      k_pcg_update<<<vec_grid(h, n), 256, 0, h->stream>>>(n, h->d_p, h->d_Ap, h->d_dinv, h->d_dU, h->d_r, h->d_ctl, rs);
      NLS_KCHECK(h);
      if ((rc = nls_allreduce_sum(h, &h->d_ctl->red[1], 2))) return rc;
      k_pcg_pupdate<<<vec_grid(h, n), 256, 0, h->stream>>>(n, h->d_r, h->d_dinv, h->d_p, h->d_ctl);
      NLS_KCHECK(h);
      if ((rc = nls_halo_exchange(h, h->d_p))) return rc;
    }
    return NLS_OK;
  };
  const bool use_graph = !h->distributed && h->opt_pcg_graph && fixed_iters == 0 && chunk == 16;
  while (launched < total) {
    const int todo = (total - launched < chunk) ? (total - launched) : chunk;
    if (use_graph && todo == chunk) {
      PcgGraph &G = h->pcg_graph;
      const bool same = G.exec && G.vals == h->d_vals && G.con == h->d_con && G.n == n && G.nnz == h->nnz &&
                        G.spmv == h->opt_spmv + (h->sym_active ? 16 : 0) && G.lanes == h->opt_spmv_lanes;
      if (!same) {
        if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
        cudaGraph_t graph = nullptr;
        const int64_t l0 = h->launches;
        NLS_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        rc = iterate(chunk);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        h->launches = l0;                              // captured, not launched
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) { h->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return NLS_ERR_CUDA; }
        ce = cudaGraphInstantiate(&G.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { G.exec = nullptr; h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return NLS_ERR_CUDA; }
        G.vals = h->d_vals; G.con = h->d_con; G.n = n; G.nnz = h->nnz; G.spmv = h->opt_spmv + (h->sym_active ? 16 : 0); G.lanes = h->opt_spmv_lanes;
      }
      NLS_CUDA(h, cudaGraphLaunch(G.exec, h->stream));
      h->launches += 3 * chunk;
    } else {
      if ((rc = iterate(todo))) return rc;
    }
    launched += todo;
    NLS_CUDA(h, cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(PcgCtl), cudaMemcpyDeviceToHost, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    { const int rcp = nls_p2p_failed(h); if (rcp) return rcp; }
    if (h->h_ctl->done) break;
  }
  const PcgCtl &c = *h->h_ctl;
  if (its) *its = c.iters;
  if (relres) *relres = c.bb > 0.0 ? sqrt(c.rr / c.bb) : 0.0;
  if (status) *status = (c.done == 1) ? 0 : (c.done == 3 ? 2 : (fixed_iters > 0 ? 0 : 1));
  return NLS_OK;
}


// ------------------------------------------------------------------------------------
// Jacobi-preconditioned MINRES (Paige & Saunders) on the free block of the handle's current CSR values: the
// symmetric-INDEFINITE solve the arc-length continuation needs past a limit point (the reference's
// bordered dense `\` of newtonarclength_fem.jl:104 becomes two of these). Vectors stay on the device; the
// Lanczos / Givens scalars are 3 doubles per iteration on the host (one synchronisation per dot product): this
// solver serves continuation runs, not the throughput path. rhs = h->d_b (masked inside), solution -> h->d_dU.
// (Round 2: the scalars moved to a device control block; the comment above describes round 1.)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_lincomb3(int64_t n, double *y, double a, const double *x1, double b, const double *x2, double c, const double *x3) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = a * x1[i];
    if (x2) v = fma(b, x2[i], v);
    if (x3) v = fma(c, x3[i], v);
    y[i] = v;
  }
}
__global__ void __launch_bounds__(256)
k_prec_abs(int64_t n, const double *__restrict__ dinv, const uint8_t *__restrict__ con, const double *__restrict__ r, double *__restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = con[i] ? 0.0 : fabs(dinv[i]) * r[i];
}

// MINRES control block: the Lanczos / Givens recurrences live on the device (no host synchronisation per iteration;
// the host reads the block back once per chunk of 16 iterations, like the PCG loop)
struct MinresCtl {
  double beta, oldb, dbar, epsln, phibar, cs, sn, beta1, rtol;
  double c_w0, c_w1, c_w2, phi;      // coefficients of the direction update, set by k_minres_scalars
  double red[2];                     // [0] = v.y (alfa), [1] = r2.y (beta^2)
  int done;                          // 0 running, 1 converged, 2 breakdown
  int itn;
};

__global__ void k_minres_start(MinresCtl *c, double rtol) {       // red[1] = beta1^2 on entry
  const double b1 = c->red[1];
  c->rtol = rtol; c->itn = 0;
  c->done = (b1 > 0.0) ? 0 : (b1 == 0.0 ? 1 : 2);
  c->beta1 = b1 > 0.0 ? sqrt(b1) : 0.0;
  c->beta = c->beta1; c->oldb = 0.0; c->dbar = 0.0; c->epsln = 0.0; c->phibar = c->beta1; c->cs = -1.0; c->sn = 0.0;
}
// v = y / beta
__global__ void __launch_bounds__(256) k_minres_v(int64_t n, const MinresCtl *c, const double *__restrict__ y, double *__restrict__ v) {
  if (c->done) return;
  const double s = 1.0 / c->beta;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = s * y[i];
}
// y -= (beta / oldb) r1 (from the second iteration on), then alfa = v.y
__global__ void __launch_bounds__(256) k_minres_y1(int64_t n, MinresCtl *c, const double *__restrict__ r1, const double *__restrict__ v,
                                                    double *__restrict__ y, RedScratch rs) {
  if (c->done) return;
  const bool sub = c->itn >= 1;
  const double f = sub ? c->beta / c->oldb : 0.0;
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double yi = y[i];
    if (sub) { yi = fma(-f, r1[i], yi); y[i] = yi; }
    acc[0] += v[i] * yi;
  }
  grid_reduce<1>(acc, rs.partial, rs.counter, &c->red[0]);
}
// y -= (alfa / beta) r2; r1 <- r2 is a pointer swap on the host; r2new = y; y = M^-1 r2new; beta^2 = r2new.y
__global__ void __launch_bounds__(256) k_minres_y2(int64_t n, MinresCtl *c, const double *__restrict__ r2, double *__restrict__ r2new,
                                                    const double *__restrict__ dinv, const uint8_t *__restrict__ con, double *__restrict__ y, RedScratch rs) {
  if (c->done) return;
  const double f = c->red[0] / c->beta;
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double t = fma(-f, r2[i], y[i]);
    r2new[i] = t;
    const double z = con[i] ? 0.0 : fabs(dinv[i]) * t;
    y[i] = z;
    acc[0] += t * z;
  }
  grid_reduce<1>(acc, rs.partial, rs.counter, &c->red[1]);
}
// the scalar recurrences of one iteration (Paige & Saunders), convergence test
__global__ void k_minres_scalars(MinresCtl *c) {
  if (c->done) return;
  const double alfa = c->red[0], b2 = c->red[1];
  c->itn += 1;
  if (!(b2 >= 0.0)) { c->done = 2; return; }                       // NaN or an indefinite preconditioner
  c->oldb = c->beta;
  const double beta = sqrt(b2);
  c->beta = beta;
  const double oldeps = c->epsln, delta = c->cs * c->dbar + c->sn * alfa, gbar = c->sn * c->dbar - c->cs * alfa;
  c->epsln = c->sn * beta; c->dbar = -c->cs * beta;
  const double gamma = fmax(sqrt(gbar * gbar + beta * beta), 2.220446049250313e-16);
  c->cs = gbar / gamma; c->sn = beta / gamma;
  c->phi = c->cs * c->phibar;
  c->phibar = c->sn * c->phibar;
  c->c_w0 = 1.0 / gamma; c->c_w1 = -oldeps / gamma; c->c_w2 = -delta / gamma;
}
// w = c0 v + c1 w1 + c2 w2; x += phi w; then the convergence flag (after the update, like the host loop did)
__global__ void __launch_bounds__(256) k_minres_w(int64_t n, MinresCtl *c, const double *__restrict__ v, const double *__restrict__ w1,
                                                   const double *__restrict__ w2, double *__restrict__ w, double *__restrict__ x) {
  if (c->done) return;
  const double c0 = c->c_w0, c1 = c->c_w1, c2 = c->c_w2, phi = c->phi;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = fma(c0, v[i], fma(c1, w1[i], c2 * w2[i]));
    w[i] = wi;
    x[i] = fma(phi, wi, x[i]);
  }
}
__global__ void k_minres_test(MinresCtl *c) {
  if (c->done) return;
  if (fabs(c->phibar) <= c->rtol * c->beta1 || c->beta == 0.0) c->done = 1;
}

int nls_minres_solve(nls_context *h, double rtol, int maxits, int *its, double *relres, int *status) {
  int rc;
  const int64_t n = h->nrows, nl = h->ndof_local();
  for (int i = 0; i < 7; ++i)
    if (!h->d_mr[i]) {
      NLS_CUDA(h, cudaMalloc((void **)&h->d_mr[i], sizeof(double) * (size_t)std::max<int64_t>(1, nl)));
      NLS_CUDA(h, cudaMemsetAsync(h->d_mr[i], 0, sizeof(double) * (size_t)std::max<int64_t>(1, nl), h->stream));
    }
  if (!h->d_mctl) {
    NLS_CUDA(h, cudaMalloc((void **)&h->d_mctl, sizeof(MinresCtl)));
    NLS_CUDA(h, cudaHostAlloc((void **)&h->h_mctl, sizeof(MinresCtl), cudaHostAllocDefault));
  }
  MinresCtl *ctl = reinterpret_cast<MinresCtl *>(h->d_mctl), *hctl = reinterpret_cast<MinresCtl *>(h->h_mctl);
  double *r1 = h->d_mr[0], *r2 = h->d_mr[1], *y = h->d_mr[2], *v = h->d_mr[3], *w = h->d_mr[4], *w1 = h->d_mr[5], *w2 = h->d_mr[6];
  double *x = h->d_dU;
  const unsigned g = vec_grid(h, n);
  RedScratch rs = red_of(h);
  if ((rc = nls_launch_extract_dinv(h))) return rc;
  if ((rc = nls_sym_prepare(h))) return rc;
  struct SymScope { nls_context *h; ~SymScope() { h->sym_active = false; } } sym_scope{h};
  NLS_CUDA(h, cudaMemsetAsync(ctl, 0, sizeof(MinresCtl), h->stream));
  NLS_CUDA(h, cudaMemsetAsync(x, 0, sizeof(double) * n, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(w, 0, sizeof(double) * n, h->stream));
  NLS_CUDA(h, cudaMemsetAsync(w2, 0, sizeof(double) * n, h->stream));
  // r1 = r2 = b, y = M^-1 b (0 on constrained rows: they carry no equation), beta1^2 = b.y
  k_lincomb3<<<g, 256, 0, h->stream>>>(n, r1, 1.0, h->d_b, 0.0, nullptr, 0.0, nullptr); NLS_KCHECK(h);
  k_lincomb3<<<g, 256, 0, h->stream>>>(n, r2, 1.0, h->d_b, 0.0, nullptr, 0.0, nullptr); NLS_KCHECK(h);
  k_prec_abs<<<g, 256, 0, h->stream>>>(n, h->d_dinv, h->d_con, h->d_b, y); NLS_KCHECK(h);
  k_dot<<<g, 256, 0, h->stream>>>(n, r1, y, nullptr, &ctl->red[1], rs); NLS_KCHECK(h);
  if ((rc = nls_allreduce_sum(h, &ctl->red[1], 1))) return rc;
  k_minres_start<<<1, 1, 0, h->stream>>>(ctl, rtol); NLS_KCHECK(h);
  int queued = 0;
  hctl->done = 0; hctl->itn = 0;
  while (queued < maxits) {
    const int todo = std::min(16, maxits - queued);
    for (int k = 0; k < todo; ++k) {
      k_minres_v<<<g, 256, 0, h->stream>>>(n, ctl, y, v); NLS_KCHECK(h);
      if (h->distributed) {
        NLS_CUDA(h, cudaMemcpyAsync(h->d_p, v, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
        if ((rc = nls_halo_exchange(h, h->d_p))) return rc;
      }
      if ((rc = nls_launch_spmv(h, h->distributed ? h->d_p : v, y, true, false))) return rc;      // y = A v on the free block
      k_minres_y1<<<g, 256, 0, h->stream>>>(n, ctl, r1, v, y, rs); NLS_KCHECK(h);
      if ((rc = nls_allreduce_sum(h, &ctl->red[0], 1))) return rc;
      k_minres_y2<<<g, 256, 0, h->stream>>>(n, ctl, r2, r1, h->d_dinv, h->d_con, y, rs); NLS_KCHECK(h);   // the new r2 overwrites the old r1
      std::swap(r1, r2);                                 // r1 <- old r2, r2 <- the vector just written
      if ((rc = nls_allreduce_sum(h, &ctl->red[1], 1))) return rc;
      k_minres_scalars<<<1, 1, 0, h->stream>>>(ctl); NLS_KCHECK(h);
      std::swap(w1, w2); std::swap(w2, w);               // w1 <- w2, w2 <- w (old w1 is recycled as the new w)
      k_minres_w<<<g, 256, 0, h->stream>>>(n, ctl, v, w1, w2, w, x); NLS_KCHECK(h);
      k_minres_test<<<1, 1, 0, h->stream>>>(ctl); NLS_KCHECK(h);
    }
    queued += todo;
    NLS_CUDA(h, cudaMemcpyAsync(hctl, ctl, sizeof(MinresCtl), cudaMemcpyDeviceToHost, h->stream));
    NLS_CUDA(h, cudaStreamSynchronize(h->stream));
    { const int rcp = nls_p2p_failed(h); if (rcp) return rcp; }
    if (hctl->done) break;
  }
  if (its) *its = hctl->itn;
  if (relres) *relres = hctl->beta1 > 0.0 ? fabs(hctl->phibar) / hctl->beta1 : 0.0;
  if (status) *status = hctl->done == 1 ? 0 : (hctl->done == 2 ? 2 : 1);
  return NLS_OK;
}
