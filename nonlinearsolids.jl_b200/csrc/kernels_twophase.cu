// kernels_twophase.cu -- "element records, then row owners" assembly of the Q1 Neo-Hookean residual
// and CSR Jacobian (2-D and 3-D). Same contract as the shared-memory gather kernel
// (kernels_gather.cu): no atomics, no zero-fill, every CSR row and R entry written exactly once,
// contributions summed in ascending element order -- the order of the reference's serial element
// loop (compute_dinternalforce, src/materials/defaults.jl:77-81 + assemble!, src/fem/element.jl:70).
//
// A hexahedron's K_e (24 x 24) cannot be exchanged through shared memory the way the 2-D kernel
// does it (300 symmetric doubles per element), and FP64 atomics bound the scatter version at
// ~1.9 G RED per C3 assembly (profiles/r1g). So the exchange goes through an element-record
// scratch in HBM:
//   kernel 1 (k_elem_q1_*_store): the element integrals; the SYMMETRIC half of K_e (3-D: 36 of
//     64 node blocks, 2-D: 10 of 16) and f_e are written structure-of-arrays,
//     scratch[entry][element], so stores coalesce across elements;
//   kernel 2 (k_rows_from_records): one thread per (owned node, component) walks the node's
//     adjacent elements in ascending order, reads its row of each record (the mirrored blocks
//     transposed), accumulates the CSR row in shared memory and writes it out coalesced.
// The host splits the node range into chunks so the scratch stays below a byte budget; a chunk
// needs the records of every element adjacent to its nodes (elements on a chunk boundary are
// evaluated by both neighbours).
#include <algorithm>
#include <cstring>
#include "asm_common.cuh"
#include "hex_element.cuh"

// ---- record layout ---------------------------------------------------------------------------
// 2-D: blocks (a,b), a <= b, slot = {0,4,7,9}[a] + (b-a); entry = slot*4 + i*2 + k; f_e at 40 + a*2 + i.
// 3-D: circulant half: slot(a,d) for block (a, (a+d)&7): d < 4 -> a*4 + d; d == 4 (a < 4) -> 32 + a;
//      entry = slot*9 + i*3 + k; f_e at 324 + a*3 + i.
template <int DIM> struct RecLayout;
template <> struct RecLayout<2> { static constexpr int NEN = 4, NBLK = 10, FE = 40, NENT = 48; };
template <> struct RecLayout<3> { static constexpr int NEN = 8, NBLK = 36, FE = 324, NENT = 348; };

// where row-owner a finds block (a,b): slot index and whether it is stored transposed
template <int DIM>
__device__ __forceinline__ int rec_slot(int a, int b, int &tr);
template <>
__device__ __forceinline__ int rec_slot<2>(int a, int b, int &tr) {
  tr = a > b;
  const int lo = tr ? b : a, hi = tr ? a : b;
  const int bd = (lo == 0) ? 0 : (lo == 1 ? 4 : (lo == 2 ? 7 : 9));
  return bd + (hi - lo);
}
template <>
__device__ __forceinline__ int rec_slot<3>(int a, int b, int &tr) {
  const int d = (b - a) & 7;
  if (d < 4) { tr = 0; return a * 4 + d; }
  if (d == 4) { tr = a >= 4; return 32 + (a & 3); }
  tr = 1;
  return b * 4 + ((a - b) & 7);
}

struct TwoPhaseArgs {
  int64_t n0, n1;          // node range of this chunk (kernel 2)
  int64_t e_lo, e_hi;      // element range held by the scratch (kernel 1), e_hi exclusive
  int64_t estride;         // scratch column stride in elements
  int64_t np;              // stride of esrc (owned nodes, padded)
  int maxadj, rs;          // adjacent elements per node (max), shared-memory row stride in doubles (odd)
  const int32_t *esrc;     // [maxadj][np] adjacent element ids ascending, -1 = none
  const uint64_t *asrc;    // [np] local node index of the owner in each adjacent element, 4 bits each
  double *scratch;         // [NENT][estride]
};

// ------------------------------------------------------------------------------------
// kernel 1, 2-D: one thread per element (symmetric half of K_e: 40 accumulators + f_e)
// ------------------------------------------------------------------------------------
template <bool DO_R, bool DO_K>
__global__ void __launch_bounds__(128)
k_elem_q1_2d_store(const AsmArgs A, const TwoPhaseArgs P, const __grid_constant__ TabQ1 T, const __grid_constant__ MatParams mp) {
  const int64_t ec = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t e = P.e_lo + ec;
  if (e >= P.e_hi) return;
  const int4 c4 = reinterpret_cast<const int4 *>(A.conn)[e];
  const int nd[4] = {c4.x, c4.y, c4.z, c4.w};
  double X[4][2], u[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double2 x = reinterpret_cast<const double2 *>(A.coords)[nd[a]];
    const double2 v = reinterpret_cast<const double2 *>(A.U)[nd[a]];
    X[a][0] = x.x; X[a][1] = x.y; u[a][0] = v.x; u[a][1] = v.y;
  }
  const double mu = mp.p[0], lam = mp.p[1];
  double fe[8], Kb[10][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) fe[i] = 0.0;
#pragma unroll
  for (int s = 0; s < 10; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) Kb[s][j] = 0.0;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    double J[4] = {0, 0, 0, 0}, Ji[4], F[4] = {1, 0, 0, 1}, Fi[4], G[4][2], g[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double d0 = T.dN[q][a][0], d1 = T.dN[q][a][1];
      G[a][0] = d0; G[a][1] = d1;
      J[0] = fma(X[a][0], d0, J[0]); J[1] = fma(X[a][0], d1, J[1]);
      J[2] = fma(X[a][1], d0, J[2]); J[3] = fma(X[a][1], d1, J[3]);
    }
    const double detJ = inv2(J, Ji);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double d0 = G[a][0], d1 = G[a][1];
      G[a][0] = fma(d0, Ji[0], d1 * Ji[2]);
      G[a][1] = fma(d0, Ji[1], d1 * Ji[3]);
#pragma unroll
      for (int i = 0; i < 2; ++i) { F[i * 2] = fma(u[a][i], G[a][0], F[i * 2]); F[i * 2 + 1] = fma(u[a][i], G[a][1], F[i * 2 + 1]); }
    }
    const double detF = inv2(F, Fi);
    const double lnJ = log(detF);
    const double wd = T.w[q] * detJ;
    const double c1 = wd * mu, c2 = wd * lam, cl = c2 * lnJ, c3 = c1 - cl;
    if (DO_R) {
      double Pk[4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int Jc = 0; Jc < 2; ++Jc) Pk[i * 2 + Jc] = fma(c1, F[i * 2 + Jc] - Fi[Jc * 2 + i], cl * Fi[Jc * 2 + i]);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 2; ++i) fe[a * 2 + i] = fma(Pk[i * 2], G[a][0], fma(Pk[i * 2 + 1], G[a][1], fe[a * 2 + i]));
    }
    if (DO_K) {
#pragma unroll
      for (int a = 0; a < 4; ++a) { g[a][0] = fma(G[a][0], Fi[0], G[a][1] * Fi[2]); g[a][1] = fma(G[a][0], Fi[1], G[a][1] * Fi[3]); }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double c1G[2] = {c1 * G[a][0], c1 * G[a][1]};
        const double c2g[2] = {c2 * g[a][0], c2 * g[a][1]};
        const double c3g[2] = {c3 * g[a][0], c3 * g[a][1]};
#pragma unroll
        for (int b = a; b < 4; ++b) {
          const int s = (a == 0 ? 0 : (a == 1 ? 4 : (a == 2 ? 7 : 9))) + (b - a);
          const double dab = fma(c1G[0], G[b][0], c1G[1] * G[b][1]);
          Kb[s][0] += fma(c2g[0], g[b][0], fma(c3g[0], g[b][0], dab));
          Kb[s][1] += fma(c2g[0], g[b][1], c3g[1] * g[b][0]);
          Kb[s][2] += fma(c2g[1], g[b][0], c3g[0] * g[b][1]);
          Kb[s][3] += fma(c2g[1], g[b][1], fma(c3g[1], g[b][1], dab));
        }
      }
    }
  }
  double *sc = P.scratch + ec;
  if (DO_K) {
#pragma unroll
    for (int s = 0; s < 10; ++s)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[(int64_t)(s * 4 + j) * P.estride] = Kb[s][j];
  }
  if (DO_R) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[(int64_t)(40 + i) * P.estride] = fe[i];
  }
}

// ------------------------------------------------------------------------------------
// kernel 1, 3-D: eight lanes per element (K_e does not fit one thread's registers).
//   phase 0: lane a gathers node a (X_a, u_a) into shared memory
//   phase 1: lane q evaluates quadrature point q: J, dN/dX, F, F^-1, lnJ, P, g -> shared memory
//   phase 2: lane a accumulates the circulant half of row block a: blocks (a, (a+d)&7), d = 0..4
//            (every unordered node pair exactly once, the four antipodal pairs twice) and f_a
//   store  : 45 + 3 values per lane, 4 consecutive elements per 32-byte sector
// ------------------------------------------------------------------------------------
#define T8_EPB 16

template <bool DO_R, bool DO_K>
__global__ void __launch_bounds__(128, 3)
k_elem_q1_3d_store(const AsmArgs A, const TwoPhaseArgs P, const __grid_constant__ TabQ1 T, const __grid_constant__ MatParams mp) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int le = tid >> 3;
  const int a = tid & 7;
  const int64_t ec = blockIdx.x * (int64_t)T8_EPB + le;
  const int64_t e = P.e_lo + ec;
  if (e >= P.e_hi) return;                  // whole 8-lane groups leave together
  const unsigned gmask = 0xffu << (8 * ((tid & 31) >> 3));
  double fe[3], Kr[5][3][3];
  hex_element_circulant<DO_R, DO_K>(smem + le * HEX_REC, a, gmask, A.conn[e * 8 + a], A.coords, A.U, T, mp.p[0], mp.p[1], Kr, fe);
  double *sc = P.scratch + ec;
  if (DO_K) {
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      if (d == 4 && a >= 4) break;
      const int slot = (d < 4) ? a * 4 + d : 32 + a;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) sc[(int64_t)(slot * 9 + i * 3 + k) * P.estride] = Kr[d][i][k];
    }
  }
  if (DO_R) {
#pragma unroll
    for (int i = 0; i < 3; ++i) sc[(int64_t)(324 + a * 3 + i) * P.estride] = fe[i];
  }
}

// ------------------------------------------------------------------------------------
// kernel 2: row owners. CTA = DIM warps x 32 nodes; warp i owns component i of 32 consecutive
// nodes, so every load of a record entry is 32 consecutive elements on a structured mesh.
// Rows are accumulated in shared memory (stride rs doubles per (node, component), odd -> no bank
// conflicts) and written out as contiguous row segments.
// ------------------------------------------------------------------------------------
template <int DIM, bool DO_R, bool DO_K>
__global__ void __launch_bounds__(32 * DIM)
k_rows_from_records(const AsmArgs A, const TwoPhaseArgs P) {
  typedef RecLayout<DIM> L;
  constexpr int NEN = L::NEN;
  extern __shared__ __align__(16) double buf[];
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t node = P.n0 + blockIdx.x * (int64_t)32 + lane;
  const bool valid = node < P.n1;
  double *row = buf + ((size_t)i * 32 + lane) * P.rs;
  int64_t b0 = 0;
  int deg = 0;
  if (valid) { b0 = A.browptr[node]; deg = (int)(A.browptr[node + 1] - b0); }
  if (DO_K)
    for (int t = 0; t < DIM * deg; ++t) row[t] = 0.0;
  double rsum = 0.0;
  if (valid) {
    const uint64_t am = P.asrc[node];
    const int64_t es = P.estride;
    for (int s = 0; s < P.maxadj; ++s) {
      const int32_t e = P.esrc[(int64_t)s * P.np + node];
      if (e < 0) break;
      const int a = (int)((am >> (4 * s)) & 15);
      const double *sc = P.scratch + ((int64_t)e - P.e_lo);
      if (DO_R) rsum += sc[(int64_t)(L::FE + a * DIM + i) * es];
      if (DO_K) {
        // column-block positions of the element's nodes in this node's block row (NEN x uint16)
        int pb[NEN];
        if (NEN == 8) {
          const uint4 pp = reinterpret_cast<const uint4 *>(A.pos)[(int64_t)e * 8 + a];
          const unsigned pw[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
          for (int b = 0; b < 8; ++b) pb[b] = (int)((pw[b >> 1] >> (16 * (b & 1))) & 0xffff);
        } else {
          const uint2 pp = reinterpret_cast<const uint2 *>(A.pos)[(int64_t)e * 4 + a];
          pb[0] = (int)(pp.x & 0xffff); pb[1] = (int)(pp.x >> 16); pb[2] = (int)(pp.y & 0xffff); pb[3] = (int)(pp.y >> 16);
        }
        double v[NEN][DIM];
#pragma unroll
        for (int b = 0; b < NEN; ++b) {
          int tr;
          const int slot = rec_slot<DIM>(a, b, tr);
          const int64_t ent0 = slot * DIM * DIM + (tr ? i : i * DIM);
          const int64_t step = tr ? DIM : 1;
#pragma unroll
          for (int k = 0; k < DIM; ++k) v[b][k] = sc[(ent0 + k * step) * es];
        }
#pragma unroll
        for (int b = 0; b < NEN; ++b) {
          double *r = row + pb[b] * DIM;
          double o[DIM];
#pragma unroll
          for (int k = 0; k < DIM; ++k) o[k] = r[k];
#pragma unroll
          for (int k = 0; k < DIM; ++k) r[k] = o[k] + v[b][k];
        }
      }
    }
    if (DO_R) A.R[node * DIM + i] = rsum;
  }
  if (DO_K) {
    __syncwarp();
    for (int p = 0; p < 32; ++p) {
      const int64_t b0p = __shfl_sync(0xffffffffu, b0, p);
      const int len = DIM * __shfl_sync(0xffffffffu, deg, p);
      double *dst = A.vals + (int64_t)DIM * DIM * b0p + (int64_t)i * len;
      const double *src = buf + ((size_t)i * 32 + p) * P.rs;
      for (int t = lane; t < len; t += 32) dst[t] = src[t];
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------
void nls_twophase_free(nls_context *h) {
  TwoPhasePlan &t = h->tplan;
  if (t.esrc) cudaFree(t.esrc);
  if (t.asrc) cudaFree(t.asrc);
  if (t.scratch) cudaFree(t.scratch);
  t = TwoPhasePlan();
}

// Builds the node -> adjacent-element lists and the chunking. Leaves t.ok false (atomic scatter is
// used instead) when the mesh does not fit the packed formats.
int nls_twophase_build(nls_context *h) {
  nls_twophase_free(h);
  TwoPhasePlan &t = h->tplan;
  if (h->dim < 2 || h->n_owned == 0 || h->nel == 0) return NLS_OK;
  const int nen = h->nen, dim = h->dim;
  const int64_t no = h->n_owned, nel = h->nel;
  const int32_t *conn = h->h_conn.data();
  std::vector<int32_t> cnt((size_t)no, 0);
  for (int64_t k = 0; k < nel * nen; ++k) if (conn[k] < no) cnt[conn[k]]++;
  int maxadj = 0, maxdeg = 0;
  for (int64_t a = 0; a < no; ++a) {
    maxadj = std::max(maxadj, (int)cnt[a]);
    maxdeg = std::max<int>(maxdeg, (int)(h->h_browptr[a + 1] - h->h_browptr[a]));
  }
  if (maxadj > 16 || maxadj == 0) return NLS_OK;
  const int rs = (dim * maxdeg) | 1;
  if ((size_t)dim * 32 * rs * sizeof(double) > (size_t)200 * 1024) return NLS_OK;
  const int64_t np = (no + 31) & ~(int64_t)31;
  std::vector<int32_t> esrc((size_t)maxadj * np, -1), emin((size_t)no, INT32_MAX), emax((size_t)no, -1);
  std::vector<uint64_t> asrc((size_t)np, 0);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int64_t e = 0; e < nel; ++e)          // ascending element id per node by construction
    for (int a = 0; a < nen; ++a) {
      const int32_t nd = conn[e * nen + a];
      if (nd >= no) continue;
      const int s = cnt[nd]++;
      esrc[(size_t)s * np + nd] = (int32_t)e;
      asrc[nd] |= (uint64_t)a << (4 * s);
      emin[nd] = std::min<int32_t>(emin[nd], (int32_t)e);
      emax[nd] = std::max<int32_t>(emax[nd], (int32_t)e);
    }
  // chunks of consecutive nodes whose adjacent elements span at most `cap` element ids
  const int nent = dim == 2 ? RecLayout<2>::NENT : RecLayout<3>::NENT;
  const size_t budget = (size_t)(h->opt_scratch_mb > 0 ? h->opt_scratch_mb : 8192) << 20;
  const int64_t cap = std::max<int64_t>(1024, (int64_t)(budget / ((size_t)nent * sizeof(double))) - 64);
  int64_t widest = 0;
  t.chunks.clear();
  for (int64_t n0 = 0; n0 < no;) {
    int64_t lo = INT32_MAX, hi = -1, n1 = n0;
    while (n1 < no) {
      const int64_t l2 = std::min<int64_t>(lo, emin[n1]), h2 = std::max<int64_t>(hi, emax[n1]);
      if (h2 >= l2 && h2 - l2 + 1 > cap && n1 > n0) break;
      lo = l2; hi = h2; ++n1;
    }
    // whole 32-node groups per chunk except for the last one, so kernel-2 warps stay aligned
    if (n1 < no && n1 - n0 > 32) {
      const int64_t n1a = n0 + ((n1 - n0) & ~(int64_t)31);
      if (n1a != n1) {
        n1 = n1a; lo = INT32_MAX; hi = -1;
        for (int64_t a = n0; a < n1; ++a) { lo = std::min<int64_t>(lo, emin[a]); hi = std::max<int64_t>(hi, emax[a]); }
      }
    }
    TwoPhaseChunk c;
    c.n0 = n0; c.n1 = n1;
    if (hi < lo) { c.e_lo = 0; c.e_hi = 0; } else { c.e_lo = lo & ~(int64_t)3; c.e_hi = hi + 1; }
    widest = std::max(widest, c.e_hi - c.e_lo);
    t.chunks.push_back(c);
    n0 = n1;
  }
  t.estride = (widest + 31) & ~(int64_t)31;
  if (t.estride == 0) t.estride = 32;
  t.np = np; t.maxadj = maxadj; t.rs = rs; t.nent = nent;
  const size_t sbytes = (size_t)nent * (size_t)t.estride * sizeof(double);
  NLS_CUDA(h, cudaMalloc((void **)&t.esrc, sizeof(int32_t) * esrc.size()));
  NLS_CUDA(h, cudaMalloc((void **)&t.asrc, sizeof(uint64_t) * asrc.size()));
  NLS_CUDA(h, cudaMemcpyAsync(t.esrc, esrc.data(), sizeof(int32_t) * esrc.size(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMemcpyAsync(t.asrc, asrc.data(), sizeof(uint64_t) * asrc.size(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  t.scratch_bytes = sbytes;
  t.ok = true;
  return NLS_OK;
}

#define TP_KCHECK(h, what) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string(what) + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

template <int DIM, bool DO_R, bool DO_K>
static int launch_twophase(nls_context *h, const AsmArgs &a) {
  const TwoPhasePlan &t = h->tplan;
  bool &attr_set = h->attr_tp3d;
  const size_t shm1 = sizeof(double) * T8_EPB * HEX_REC;
  const size_t shm2 = sizeof(double) * DIM * 32 * (size_t)t.rs;
  if (DIM == 3 && !attr_set) {
    cudaFuncSetAttribute(k_elem_q1_3d_store<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    cudaFuncSetAttribute(k_elem_q1_3d_store<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    cudaFuncSetAttribute(k_elem_q1_3d_store<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    attr_set = true;
  }
  bool &rows_set = h->attr_rows[DIM - 2][DO_R ? 1 : 0][DO_K ? 1 : 0];
  if (DO_K && !rows_set) {
    cudaFuncSetAttribute(k_rows_from_records<DIM, DO_R, DO_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    rows_set = true;
  }
  for (const TwoPhaseChunk &c : t.chunks) {
    TwoPhaseArgs p{c.n0, c.n1, c.e_lo, c.e_hi, t.estride, t.np, t.maxadj, t.rs, t.esrc, t.asrc, t.scratch};
    const int64_t ne = c.e_hi - c.e_lo;
    if (ne > 0) {
      if (DIM == 2) k_elem_q1_2d_store<DO_R, DO_K><<<(unsigned)((ne + 127) / 128), 128, 0, h->stream>>>(a, p, h->tabq, h->mat);
      else k_elem_q1_3d_store<DO_R, DO_K><<<(unsigned)((ne + T8_EPB - 1) / T8_EPB), 128, shm1, h->stream>>>(a, p, h->tabq, h->mat);
      TP_KCHECK(h, "element-record kernel launch: ");
    }
    const int64_t nn = c.n1 - c.n0;
    k_rows_from_records<DIM, DO_R, DO_K><<<(unsigned)((nn + 31) / 32), 32 * DIM, DO_K ? shm2 : 0, h->stream>>>(a, p);
    TP_KCHECK(h, "row-owner kernel launch: ");
  }
  return NLS_OK;
}

int nls_launch_twophase(nls_context *h, bool want_r, bool want_k) {
  AsmArgs a{};
  a.nel = h->nel; a.n_owned = h->n_owned; a.conn = h->d_conn; a.coords = h->d_coords; a.U = h->d_U;
  a.R = h->d_R; a.vals = h->d_vals; a.browptr = h->d_browptr; a.pos = h->d_pos;
  if (h->dim == 2) {
    if (want_r && want_k) return launch_twophase<2, true, true>(h, a);
    if (want_r) return launch_twophase<2, true, false>(h, a);
    return launch_twophase<2, false, true>(h, a);
  }
  if (want_r && want_k) return launch_twophase<3, true, true>(h, a);
  if (want_r) return launch_twophase<3, true, false>(h, a);
  return launch_twophase<3, false, true>(h, a);
}
