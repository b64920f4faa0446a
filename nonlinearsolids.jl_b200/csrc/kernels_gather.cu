// kernels_gather.cu -- gather ("row-owner") assembly of the 2-D Q1 Neo-Hookean residual and CSR
// Jacobian: no atomics, no zero-fill, every CSR row and R entry written exactly once, contributions
// summed in ascending element order -- the order of the reference's serial element loop
// (compute_dinternalforce, src/materials/defaults.jl:77-81 + assemble!, src/fem/element.jl:70) --
// so results are bit-reproducible run to run.
//
// Persistent CTAs walk node clusters (host plan: tiles of 15 x 7 nodes, nls_host_cluster_plan: their 16 x 8 = 128
// touching elements are exactly one per thread, one warp per SM sub-partition). Everything a cluster needs is
// staged into shared memory with cp.async (LDGSTS), so no global-load latency sits on the critical path:
//   group A (k+2): header + list of nodes touched by cluster k+2           (contiguous copy, two clusters ahead)
//   group B1(k+1): coordinates + displacements of those nodes (16-byte gathers through the list) and the local
//                  connectivity: the phase-1 inputs, double buffered, one cluster ahead
//   group B2(k)  : row owners, 32-bit row bases relative to the cluster, per-block tasks and source lists: the
//                  phase-2 lists, single buffered, fetched while phase 1 of the same cluster runs
//   compute(k)   : phase 1  one thread per element: the symmetric half of K_e over the 4 quadrature points (rolled
//                           loop, 40 register accumulators; f_e accumulated in the element's record) -> smem
//                  phase 2  one task per CSR block of the cluster's rows: sums the element blocks that land in it
//                           (ascending element order, branch-free padded list) and writes its two 16-byte row
//                           segments; adjacent tasks write adjacent segments, so the stores coalesce.
//                           One task per node sums R.
// Clusters that touch a ghost node (multi-GPU) are listed last: they are launched separately, after the ghost
// exchange that runs on a second stream beside the interior clusters has landed (nls_api.cu: assemble()).
// History and counters behind this design: DESIGN.md section 5, profiles/r1a..r1k.
#include <algorithm>
#include <cstring>
#include "asm_common.cuh"

#define G2_TX 15
#define G2_TY 7
#define G2_NC (G2_TX * G2_TY)   // owned nodes per cluster: a 15 x 7 tile is touched by 16 x 8 = 128 elements -- one per thread,
                                // four full warps (one per SM sub-partition), 22 % of the elements evaluated twice
#define G2_NCP 112              // padded stride of the per-node plan arrays (16-byte rows)
#define G2_THREADS 128
#define G2_KS 50          // doubles per element in shared memory: 10 symmetric 2x2 blocks (40) + f_e (8) + 2 pad;
                          // 25 x 16 B (odd) -> the reduce-scatter's STS.128 are conflict-free
#define G2_MAXSRC 8       // most elements around one node the packed block lists can hold
#define G2_KREG 10        // symmetric blocks accumulated in registers (of 10); any rest lives in the element's smem record

#define G2_HDR 8          // int32 words of the cluster header in front of the touched-node list

struct GatherArgs {
  int64_t ncl;            // this launch walks clusters [c_begin, ncl)
  int64_t c_begin;
  int maxt, maxe, maxb, nsw;  // padded strides: touched nodes, elements, CSR blocks; source words per block (2 or 4)
  const int32_t *tn;      // [ncl][G2_HDR + maxt]: {nt, ne, nb, 0, base_lo, base_hi, 0, 0}, node ids; base = smallest block-row start
  const uint32_t *lconn;  // [ncl][maxe]: 4 local node indices (bytes) into the touched-node list
  const int32_t *cn;      // [ncl][G2_NCP] owned node ids (-1 = none)
  const int32_t *b0;      // [ncl][G2_NCP] block-row start of the node, relative to the cluster's base
  const int32_t *deg;     // [ncl][G2_NCP] block-row length
  const uint16_t *btask;  // [ncl][maxb]: one task per CSR block: owner nl | column block jb << 7
  const uint32_t *bsrc;   // [ncl][nsw][maxb]: 16-bit source codes, two per word: offset of the 2x2 block in Ks (15 bits) |
                          //   transposed << 15; unused sources point at the zero record Ks[maxe]
  const uint16_t *rsrc;   // [ncl][G2_NCP][2 nsw]: offsets in Ks of the node's f_e pairs, ascending element; unused -> zero record
  unsigned long long *prof; // optional per-phase cycle counters (nls_set_option "profile_phases"), else NULL
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One quadrature point of one quad: geometry chain (J, J^-1, F, F^-1, log), then this point's share of the
// symmetric half of K_e and of f_e is added to the accumulators:
//   Kb: blocks (a,b), a <= b, in the order (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3), each [i][k] row-major.
template <bool DO_R, bool DO_K>
// The nodal coordinates / displacements are re-read from the staged copy in shared memory at every point
// (8 LDS.128) instead of being held across the loop: 32 registers the accumulators need.
__device__ __forceinline__ void q1_2d_qp_accumulate(const double2 *__restrict__ xu, const unsigned lc, const double *__restrict__ Tq,
                                                    const double w, const double mu, const double lam,
                                                    double (&Kb)[G2_KREG][4], double2 *__restrict__ rec, const bool first) {
  double J[4] = {0, 0, 0, 0}, Ji[4], F[4] = {1, 0, 0, 1}, Fi[4], G[4][2], g[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double2 x = xu[2 * ((lc >> (8 * a)) & 0xff)];
    const double d0 = Tq[a * 2], d1 = Tq[a * 2 + 1];
    G[a][0] = d0; G[a][1] = d1;
    J[0] = fma(x.x, d0, J[0]); J[1] = fma(x.x, d1, J[1]);
    J[2] = fma(x.y, d0, J[2]); J[3] = fma(x.y, d1, J[3]);
  }
  const double detJ = inv2(J, Ji);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double2 uu = xu[2 * ((lc >> (8 * a)) & 0xff) + 1];
    const double d0 = G[a][0], d1 = G[a][1];
    G[a][0] = fma(d0, Ji[0], d1 * Ji[2]);
    G[a][1] = fma(d0, Ji[1], d1 * Ji[3]);
    F[0] = fma(uu.x, G[a][0], F[0]); F[1] = fma(uu.x, G[a][1], F[1]);
    F[2] = fma(uu.y, G[a][0], F[2]); F[3] = fma(uu.y, G[a][1], F[3]);
  }
  const double detF = inv2(F, Fi);
  const double lnJ = log(detF);
  const double wd = w * detJ;
  const double c1 = wd * mu, c2 = wd * lam, wcl = c2 * lnJ, c3 = c1 - wcl;
  if (DO_R) {
    const double P00 = fma(c1, F[0] - Fi[0], wcl * Fi[0]), P01 = fma(c1, F[1] - Fi[2], wcl * Fi[2]);
    const double P10 = fma(c1, F[2] - Fi[1], wcl * Fi[1]), P11 = fma(c1, F[3] - Fi[3], wcl * Fi[3]);
    // f_e is accumulated in the element's own shared-memory record (thread-private, no synchronisation):
    // 16 registers fewer across the loop than keeping it beside the 40 K accumulators
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double2 o = rec[20 + a];
      if (first) o = make_double2(0.0, 0.0);
      o.x = fma(P00, G[a][0], fma(P01, G[a][1], o.x));
      o.y = fma(P10, G[a][0], fma(P11, G[a][1], o.y));
      rec[20 + a] = o;
    }
  }
  if (DO_K) {
#pragma unroll
    for (int a = 0; a < 4; ++a) { g[a][0] = fma(G[a][0], Fi[0], G[a][1] * Fi[2]); g[a][1] = fma(G[a][0], Fi[1], G[a][1] * Fi[3]); }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double c1Gx = c1 * G[a][0], c1Gy = c1 * G[a][1];
      const double c2g0 = c2 * g[a][0], c2g1 = c2 * g[a][1], c3g0 = c3 * g[a][0], c3g1 = c3 * g[a][1];
      const double d0 = c2g0 + c3g0, d1 = c2g1 + c3g1;
#pragma unroll
      for (int b = a; b < 4; ++b) {
        const int s = ((a == 0) ? 0 : (a == 1 ? 4 : (a == 2 ? 7 : 9))) + (b - a);
        if (s < G2_KREG) {
          Kb[s][0] = fma(d0, g[b][0], fma(c1Gx, G[b][0], fma(c1Gy, G[b][1], Kb[s][0])));
          Kb[s][1] = fma(c2g0, g[b][1], fma(c3g1, g[b][0], Kb[s][1]));
          Kb[s][2] = fma(c2g1, g[b][0], fma(c3g0, g[b][1], Kb[s][2]));
          Kb[s][3] = fma(d1, g[b][1], fma(c1Gx, G[b][0], fma(c1Gy, G[b][1], Kb[s][3])));
        } else {      // the last blocks are accumulated in the element's own record like f_e: registers for a third CTA per SM
          double2 o0 = rec[2 * s], o1 = rec[2 * s + 1];
          if (first) { o0 = make_double2(0.0, 0.0); o1 = make_double2(0.0, 0.0); }
          o0.x = fma(d0, g[b][0], fma(c1Gx, G[b][0], fma(c1Gy, G[b][1], o0.x)));
          o0.y = fma(c2g0, g[b][1], fma(c3g1, g[b][0], o0.y));
          o1.x = fma(c2g1, g[b][0], fma(c3g0, g[b][1], o1.x));
          o1.y = fma(d1, g[b][1], fma(c1Gx, G[b][0], fma(c1Gy, G[b][1], o1.y)));
          rec[2 * s] = o0; rec[2 * s + 1] = o1;
        }
      }
    }
  }
}

// Residency: the fused pass needs 222 registers to stay spill-free (two CTAs per SM); the Jacobian-only and
// residual-only passes fit 168 and run three CTAs per SM (measured: fused 0.124 ms at 2 vs 0.134 ms at 3 with
// spills; Jacobian-only 0.114 -> 0.105 ms at 3).
template <bool DO_R, bool DO_K>
__global__ void __launch_bounds__(G2_THREADS, (DO_R && DO_K) ? 2 : 3)
k_asm_q1_2d_gather(const AsmArgs A, const GatherArgs P, const __grid_constant__ TabQ1 T,
                   const __grid_constant__ MatParams mp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- shared-memory carve-up (all sub-buffers 16-byte aligned); 74.5 KB at C2 -> three CTAs per SM ----
  const int tn_words = G2_HDR + P.maxt;
  double *Ks = reinterpret_cast<double *>(smem_raw);                                  // [maxe + 1][G2_KS], last record = zeros
  double *Tsm = Ks + (size_t)(P.maxe + 1) * G2_KS;                                    // 40 doubles: dN[4][4][2], w[4], pad
  int32_t *tnb = reinterpret_cast<int32_t *>(Tsm + 40);                               // [3][tn_words]
  unsigned char *b1base = reinterpret_cast<unsigned char *>(tnb + 3 * tn_words);     // [2] stages of the phase-1 inputs
  const size_t b1_xu = 0, b1_lc = (size_t)P.maxt * 32, b1_size = b1_lc + (size_t)P.maxe * 4;
  unsigned char *B2 = b1base + 2 * b1_size;                                           // single stage of the phase-2 lists
  const size_t b2_b0 = 0, b2_cn = b2_b0 + G2_NCP * 4, b2_dg = b2_cn + G2_NCP * 4, b2_rs = b2_dg + G2_NCP * 4,
               b2_bt = b2_rs + (size_t)G2_NCP * 4 * P.nsw, b2_bs = b2_bt + (((size_t)P.maxb * 2 + 15) & ~(size_t)15);
  const int tid = threadIdx.x;
  if (tid < 32) Tsm[tid] = T.dN[tid >> 3][(tid >> 1) & 3][tid & 1];
  else if (tid < 36) Tsm[tid] = T.w[tid - 32];
  else if (tid >= 64 && tid < 64 + G2_KS) Ks[(size_t)P.maxe * G2_KS + tid - 64] = 0.0;

  auto issue_A = [&](int64_t c, int k) {          // header + node list of cluster c -> tnb[k % 3]
    if (c < P.ncl) {
      const int32_t *src = P.tn + c * tn_words;
      int32_t *dst = tnb + (k % 3) * tn_words;
      for (int w = tid * 4; w < tn_words; w += G2_THREADS * 4) cp_async16(dst + w, src + w);
    }
    cp_async_commit();
  };
  auto issue_B1 = [&](int64_t c, int k) {         // phase-1 inputs of cluster c -> stage k & 1 (one cluster ahead)
    if (c < P.ncl) {
      unsigned char *B = b1base + (size_t)(k & 1) * b1_size;
      const int32_t *tn = tnb + (k % 3) * tn_words;
      const int nt = tn[0];
      double *xu = reinterpret_cast<double *>(B + b1_xu);
      for (int t = tid; t < nt; t += G2_THREADS) {
        const int id = tn[G2_HDR + t];
        cp_async16(xu + 4 * t, A.coords + 2 * (int64_t)id);
        cp_async16(xu + 4 * t + 2, A.U + 2 * (int64_t)id);
      }
      for (int w = tid * 4; w < P.maxe; w += G2_THREADS * 4) cp_async16(B + b1_lc + w * 4, P.lconn + c * P.maxe + w);
    }
    cp_async_commit();
  };
  auto issue_B2 = [&](int64_t c) {                // phase-2 lists of cluster c (same iteration: they land during phase 1)
    for (int w = tid * 4; w < G2_NCP; w += G2_THREADS * 4) {
      cp_async16(B2 + b2_b0 + w * 4, P.b0 + c * G2_NCP + w);
      cp_async16(B2 + b2_cn + w * 4, P.cn + c * G2_NCP + w);
      cp_async16(B2 + b2_dg + w * 4, P.deg + c * G2_NCP + w);
    }
    const int nrs = G2_NCP * 2 * P.nsw;           // uint16 entries, multiple of 8
    for (int w = tid * 8; w < nrs; w += G2_THREADS * 8) cp_async16(B2 + b2_rs + w * 2, P.rsrc + c * nrs + w);
    const int nbt = (P.maxb + 7) & ~7;            // uint16 entries
    for (int w = tid * 8; w < nbt; w += G2_THREADS * 8) cp_async16(B2 + b2_bt + w * 2, P.btask + c * nbt + w);
    const int nbw = P.nsw * P.maxb;               // uint32 entries, multiple of 4
    for (int w = tid * 4; w < nbw; w += G2_THREADS * 4) cp_async16(B2 + b2_bs + w * 4, P.bsrc + c * nbw + w);
    cp_async_commit();
  };

  const double mu = mp.p[0], lam = mp.p[1];
  const int64_t c0 = P.c_begin + blockIdx.x, cstep = gridDim.x;
  issue_A(c0, 0);
  issue_A(c0 + cstep, 1);
  cp_async_wait<1>();            // A(0) landed
  __syncthreads();
  issue_B1(c0, 0);

  int k = 0;
  for (int64_t c = c0; c < P.ncl; c += cstep, ++k) {
    const long long tp0 = P.prof ? clock64() : 0;
    issue_A(c + 2 * cstep, k + 2);
    cp_async_wait<1>();          // everything but A(k+2): A(k+1) and B1(k) have landed
    __syncthreads();             // ... and every warp is done with the previous cluster's Ks and phase-2 lists
    issue_B2(c);                 // lands while phase 1 runs
    issue_B1(c + cstep, k + 1);  // the next cluster's phase-1 inputs
    const long long tp1 = P.prof ? clock64() : 0;

    const unsigned char *B = b1base + (size_t)(k & 1) * b1_size;
    const int32_t *hdr = tnb + (k % 3) * tn_words;
    const int ne = hdr[1], nb = hdr[2];
    const int64_t base = (int64_t)(uint32_t)hdr[4] | ((int64_t)hdr[5] << 32);
    // ---- phase 1: one thread per element, quadrature loop rolled (the accumulators are the register budget) ----
    {
      const double *xu = reinterpret_cast<const double *>(B + b1_xu);
      const unsigned *lcv = reinterpret_cast<const unsigned *>(B + b1_lc);
      for (int sl = tid; sl < ne; sl += G2_THREADS) {
        const unsigned lc = lcv[sl];
        double Kb[G2_KREG][4];
#pragma unroll
        for (int s = 0; s < G2_KREG; ++s)
#pragma unroll
          for (int j = 0; j < 4; ++j) Kb[s][j] = 0.0;
        double2 *dst = reinterpret_cast<double2 *>(Ks + (size_t)sl * G2_KS);
#pragma unroll 1
        for (int q = 0; q < 4; ++q)
          q1_2d_qp_accumulate<DO_R, DO_K>(reinterpret_cast<const double2 *>(xu), lc, Tsm + q * 8, Tsm[32 + q], mu, lam, Kb, dst, q == 0);
        if (DO_K) {
#pragma unroll
          for (int s = 0; s < G2_KREG; ++s) { dst[2 * s] = make_double2(Kb[s][0], Kb[s][1]); dst[2 * s + 1] = make_double2(Kb[s][2], Kb[s][3]); }
        }
      }
    }
    cp_async_wait<1>();          // everything but B1(k+1): the phase-2 lists B2(k) have landed
    __syncthreads();
    const long long tp2 = P.prof ? clock64() : 0;
    // ---- phase 2: one task per CSR block of the cluster's rows: sum the element blocks that land in it, in
    //      ascending element order (padded list: unused sources read the zero record), and write its two row
    //      segments (adjacent tasks -> adjacent 16 B). R: one task per node over its f_e list. ----
    if (DO_K) {
      const uint16_t *bt = reinterpret_cast<const uint16_t *>(B2 + b2_bt);
      const uint32_t *bs = reinterpret_cast<const uint32_t *>(B2 + b2_bs);
      const int32_t *b0v = reinterpret_cast<const int32_t *>(B2 + b2_b0);
      const int32_t *dgv = reinterpret_cast<const int32_t *>(B2 + b2_dg);
      const unsigned ks_s = (unsigned)__cvta_generic_to_shared(Ks);
      double *vbase = A.vals + 4 * base;
      auto ld2 = [](unsigned addr, double2 &lo_, double2 &hi_) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(lo_.x), "=d"(lo_.y) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(hi_.x), "=d"(hi_.y) : "r"(addr));
      };
      if (P.nsw == 2) {
        // structured-like meshes (<= 4 elements around a node): all four sources of a task are loaded up front and
        // two tasks are interleaved, so 16 LDS.128 are in flight per thread instead of a dependent chain
#pragma unroll 2
        for (int blk = tid; blk < nb; blk += G2_THREADS) {
          const uint32_t w0 = bt[blk], c0w = bs[blk], c1w = bs[P.maxb + blk];
          double2 x[4][2];
          ld2(ks_s + ((c0w & 0x7fffu) << 3), x[0][0], x[0][1]);
          ld2(ks_s + ((c0w >> 13) & 0x3fff8u), x[1][0], x[1][1]);
          ld2(ks_s + ((c1w & 0x7fffu) << 3), x[2][0], x[2][1]);
          ld2(ks_s + ((c1w >> 13) & 0x3fff8u), x[3][0], x[3][1]);
          const bool tr[4] = {(c0w & 0x8000u) != 0, (c0w & 0x80000000u) != 0, (c1w & 0x8000u) != 0, (c1w & 0x80000000u) != 0};
          double2 lo = make_double2(0.0, 0.0), hi = make_double2(0.0, 0.0);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {      // ascending element order
            lo.x += x[s4][0].x; lo.y += tr[s4] ? x[s4][1].x : x[s4][0].y;
            hi.x += tr[s4] ? x[s4][0].y : x[s4][1].x; hi.y += x[s4][1].y;
          }
          const int nl = w0 & 0x7f, jb = w0 >> 7;
          double *row = vbase + 4 * (int64_t)b0v[nl] + 2 * jb;
          *reinterpret_cast<double2 *>(row) = lo;
          *reinterpret_cast<double2 *>(row + 2 * dgv[nl]) = hi;
        }
      } else {
        for (int blk = tid; blk < nb; blk += G2_THREADS) {
          const uint32_t w0 = bt[blk];
          const int nl = w0 & 0x7f, jb = w0 >> 7;
          double2 lo = make_double2(0.0, 0.0), hi = make_double2(0.0, 0.0);
          for (int sw = 0; sw < P.nsw; ++sw) {
            const uint32_t cw = bs[sw * P.maxb + blk];
            double2 x0, x1, y0, y1;
            ld2(ks_s + ((cw & 0x7fffu) << 3), x0, x1);
            ld2(ks_s + ((cw >> 13) & 0x3fff8u), y0, y1);
            const bool tr0 = (cw & 0x8000u) != 0, tr1 = (cw & 0x80000000u) != 0;
            lo.x += x0.x; lo.y += tr0 ? x1.x : x0.y; hi.x += tr0 ? x0.y : x1.x; hi.y += x1.y;
            lo.x += y0.x; lo.y += tr1 ? y1.x : y0.y; hi.x += tr1 ? y0.y : y1.x; hi.y += y1.y;
          }
          double *row = vbase + 4 * (int64_t)b0v[nl] + 2 * jb;
          *reinterpret_cast<double2 *>(row) = lo;
          *reinterpret_cast<double2 *>(row + 2 * dgv[nl]) = hi;
        }
      }
    }
    if (DO_R) {
      const int32_t *cnv = reinterpret_cast<const int32_t *>(B2 + b2_cn);
      const uint32_t *rsv = reinterpret_cast<const uint32_t *>(B2 + b2_rs);
      for (int nl = tid; nl < G2_NC; nl += G2_THREADS) {
        const int node = cnv[nl];
        if (node < 0) continue;
        double2 rs = make_double2(0.0, 0.0);
        for (int sw = 0; sw < P.nsw; ++sw) {
          const uint32_t rw = rsv[nl * P.nsw + sw];
          const double2 f0 = *reinterpret_cast<const double2 *>(Ks + (rw & 0xffffu));
          const double2 f1 = *reinterpret_cast<const double2 *>(Ks + (rw >> 16));
          rs.x += f0.x; rs.y += f0.y;
          rs.x += f1.x; rs.y += f1.y;
        }
        *reinterpret_cast<double2 *>(A.R + 2 * (int64_t)node) = rs;
      }
    }
    if (P.prof && tid == 0) {
      const long long tp3 = clock64();
      atomicAdd(P.prof + 0, (unsigned long long)(tp1 - tp0));   // staging: waits + issue
      atomicAdd(P.prof + 1, (unsigned long long)(tp2 - tp1));   // phase 1
      atomicAdd(P.prof + 3, (unsigned long long)(tp3 - tp2));   // phase 2 (thread 0's own tasks)
      atomicAdd(P.prof + 4, 1ull);                               // clusters
    }
  }
  cp_async_wait<0>();
}

// ---- host side: pack the cluster plan into fixed-stride device arrays ---------------------------
int nls_assembly_mode(nls_context *h) {
  if (h->opt_assembly == 0 || h->dim < 2 || h->mat.kind == NLS_MAT_J2_SMALL_STRAIN) return NLS_ASM_ATOMIC;
  if (h->dim == 2 && h->opt_assembly == 1 && h->have_plan) return NLS_ASM_GATHER;
  TwoPhasePlan &t = h->tplan;
  if (h->dim == 3 && h->opt_assembly == 1 && h->lplan.ok) return NLS_ASM_LINES;
  if (h->dim == 3 && h->opt_assembly == 1 && !h->opt_twophase_3d) return NLS_ASM_ATOMIC;
  if (t.ok && !t.scratch) {       // first use: the record scratch; without room for it keep the atomic scatter
    if (cudaMalloc((void **)&t.scratch, t.scratch_bytes) != cudaSuccess) { cudaGetLastError(); t.scratch = nullptr; t.ok = false; }
  }
  if (t.ok) return NLS_ASM_TWOPHASE;
  return (h->dim == 2 && h->have_plan) ? NLS_ASM_GATHER : NLS_ASM_ATOMIC;
}
bool nls_assembly_overwrites(nls_context *h) { return nls_assembly_mode(h) != NLS_ASM_ATOMIC; }
int nls_gather_nc(void) { return G2_NC; }
void nls_gather_tile(int *t) { t[0] = G2_TX; t[1] = G2_TY; t[2] = 1; }

static size_t gather_smem(int maxt, int maxe, int maxb, int nsw) {
  const size_t b1 = (size_t)maxt * 32 + (size_t)maxe * 4;
  const size_t b2 = (size_t)G2_NCP * 12 + (size_t)G2_NCP * 4 * nsw + (((size_t)maxb * 2 + 15) & ~(size_t)15) + (size_t)nsw * maxb * 4;
  return (size_t)(maxe + 1) * G2_KS * 8 + 40 * 8 + (size_t)3 * (G2_HDR + maxt) * 4 + 2 * b1 + b2;
}

void nls_gather_free(nls_context *h) {
  GatherPlanDev &g = h->gplan;
  if (g.tn) cudaFree(g.tn);
  if (g.lconn) cudaFree(g.lconn);
  if (g.cn) cudaFree(g.cn);
  if (g.b0) cudaFree(g.b0);
  if (g.deg) cudaFree(g.deg);
  if (g.btask) cudaFree(g.btask);
  if (g.bsrc) cudaFree(g.bsrc);
  if (g.rsrc) cudaFree(g.rsrc);
  g = GatherPlanDev();
  h->have_plan = false;
}

// Sets h->have_plan when the mesh fits the kernel's packed formats; otherwise leaves it false and
// another scheme is used instead.
int nls_gather_build(nls_context *h, const AsmPlan &plan) {
  nls_gather_free(h);
  if (h->dim != 2 || plan.ncl == 0) return NLS_OK;
  const int nen = 4;
  const int64_t ncl = plan.ncl;
  const int32_t *conn = h->h_conn.data();
  int maxdeg = 0;
  for (int64_t a = 0; a < h->n_owned; ++a) maxdeg = std::max<int>(maxdeg, (int)(h->h_browptr[a + 1] - h->h_browptr[a]));
  if ((size_t)(plan.max_elems + 5) * G2_KS > 0x7fff || maxdeg > 127 || plan.max_adj > G2_MAXSRC) return NLS_OK;
  std::vector<std::vector<int32_t>> touched((size_t)ncl);
  int maxt = 0, maxb = 0;
  for (int64_t c = 0; c < ncl; ++c) {
    std::vector<int32_t> &t = touched[c];
    for (int32_t s = plan.eptr[c]; s < plan.eptr[c + 1]; ++s)
      t.insert(t.end(), conn + (int64_t)plan.elems[s] * nen, conn + (int64_t)(plan.elems[s] + 1) * nen);
    std::sort(t.begin(), t.end());
    t.erase(std::unique(t.begin(), t.end()), t.end());
    maxt = std::max<int>(maxt, (int)t.size());
    int nb = 0;
    int64_t lo = INT64_MAX, hi = 0;
    for (int nl = 0; nl < G2_NC; ++nl) {
      const int32_t a = plan.nodes[c * G2_NC + nl];
      if (a < 0) continue;
      nb += (int)(h->h_browptr[a + 1] - h->h_browptr[a]);
      lo = std::min(lo, h->h_browptr[a]); hi = std::max(hi, h->h_browptr[a]);
    }
    if (nb > 0 && hi - lo > 0x1fffffff) return NLS_OK;      // block rows of one cluster too far apart for 32-bit offsets
    maxb = std::max(maxb, nb);
  }
  if (maxt > 255) return NLS_OK;
  const int MAXT = (maxt + 3) & ~3, MAXE = (plan.max_elems + 3) & ~3, MAXB = (maxb + 7) & ~7;
  const int NSW = plan.max_adj <= 4 ? 2 : 4;       // 16-bit source codes, two per word
  if (gather_smem(MAXT, MAXE, MAXB, NSW) > (size_t)110 * 1024) return NLS_OK;   // at least two CTAs per SM
  // clusters that touch a ghost node go last: the multi-GPU assembly runs the others while the halo is in flight
  std::vector<int64_t> order((size_t)ncl);
  int64_t n_interior = 0;
  {
    std::vector<int64_t> bnd;
    for (int64_t c = 0; c < ncl; ++c) {
      if (!touched[c].empty() && touched[c].back() >= h->n_owned) bnd.push_back(c); else order[n_interior++] = c;
    }
    std::copy(bnd.begin(), bnd.end(), order.begin() + n_interior);
  }
  const int tnw = G2_HDR + MAXT;
  std::vector<int32_t> tn((size_t)ncl * tnw, 0), cn((size_t)ncl * G2_NCP, -1), deg((size_t)ncl * G2_NCP, 0), b0((size_t)ncl * G2_NCP, 0);
  std::vector<uint32_t> lconn((size_t)ncl * MAXE, 0);
  const uint32_t none = (uint32_t)(MAXE * G2_KS);   // the zero record, not transposed
  std::vector<uint32_t> bsrc((size_t)ncl * NSW * MAXB, none | (none << 16));
  std::vector<uint16_t> btask((size_t)ncl * MAXB, 0);
  std::vector<uint16_t> rsrc((size_t)ncl * G2_NCP * 2 * NSW, (uint16_t)(none + 40));
  static const int bd[4] = {0, 4, 7, 9};           // first symmetric block of local row a
  std::vector<int> nsrc_of;
  for (int64_t c = 0; c < ncl; ++c) {         // c = position in the device arrays, pc = cluster of the host plan
    const int64_t pc = order[c];
    const std::vector<int32_t> &t = touched[pc];
    const int32_t e0 = plan.eptr[pc], ne = plan.eptr[pc + 1] - e0;
    tn[c * tnw] = (int32_t)t.size(); tn[c * tnw + 1] = ne;
    std::copy(t.begin(), t.end(), tn.begin() + c * tnw + G2_HDR);
    for (int32_t s = 0; s < ne; ++s) {
      uint32_t packed = 0;
      for (int a = 0; a < nen; ++a) {
        const int32_t nd = conn[(int64_t)plan.elems[e0 + s] * nen + a];
        packed |= (uint32_t)(std::lower_bound(t.begin(), t.end(), nd) - t.begin()) << (8 * a);
      }
      lconn[c * MAXE + s] = packed;
    }
    int64_t base = INT64_MAX;
    for (int nl = 0; nl < G2_NC; ++nl) {
      const int32_t a = plan.nodes[pc * G2_NC + nl];
      if (a >= 0) base = std::min(base, h->h_browptr[a]);
    }
    if (base == INT64_MAX) base = 0;
    tn[c * tnw + 4] = (int32_t)(uint32_t)(base & 0xffffffffll); tn[c * tnw + 5] = (int32_t)(base >> 32);
    uint16_t *bt = btask.data() + (size_t)c * MAXB;
    uint32_t *bs = bsrc.data() + (size_t)c * NSW * MAXB;
    int nb = 0;
    for (int nl = 0; nl < G2_NC; ++nl) {
      const int32_t a = plan.nodes[pc * G2_NC + nl];
      if (a < 0) continue;
      cn[c * G2_NCP + nl] = a;
      b0[c * G2_NCP + nl] = (int32_t)(h->h_browptr[a] - base);
      const int dg = (int)(h->h_browptr[a + 1] - h->h_browptr[a]);
      deg[c * G2_NCP + nl] = dg;
      const int32_t *lo = h->h_bcol.data() + h->h_browptr[a], *hi = lo + dg;
      const int blk0 = nb;
      nsrc_of.assign(dg, 0);
      for (int jb = 0; jb < dg; ++jb) bt[blk0 + jb] = (uint16_t)(nl | (jb << 7));
      int nrs = 0;
      nb += dg;
      for (int32_t s = 0; s < ne; ++s) {          // ascending slot = ascending element id
        const int32_t *ce = conn + (int64_t)plan.elems[e0 + s] * nen;
        int al = -1;
        for (int q = 0; q < nen; ++q) if (ce[q] == a) al = q;
        if (al < 0) continue;
        if (nrs >= 2 * NSW) return NLS_OK;
        rsrc[((size_t)c * G2_NCP + nl) * 2 * NSW + nrs++] = (uint16_t)(s * G2_KS + 40 + al * 2);
        for (int b = 0; b < nen; ++b) {
          const int jb = (int)(std::lower_bound(lo, hi, ce[b]) - lo);
          const int k = nsrc_of[jb]++;
          if (k >= 2 * NSW) return NLS_OK;        // more elements on one block than the packed list holds
          const int lo_ = std::min(al, b), hi_ = std::max(al, b);
          const uint32_t code = (uint32_t)(s * G2_KS + (bd[lo_] + (hi_ - lo_)) * 4) | (al > b ? 0x8000u : 0u);
          uint32_t &wsrc = bs[(size_t)(k / 2) * MAXB + blk0 + jb];
          wsrc = (k & 1) ? ((wsrc & 0x0000ffffu) | (code << 16)) : ((wsrc & 0xffff0000u) | code);
        }
      }
    }
    tn[c * tnw + 2] = nb;
  }
  GatherPlanDev &g = h->gplan;
  g.ncl = ncl; g.ncl_interior = n_interior; g.maxt = MAXT; g.maxe = MAXE; g.maxb = MAXB; g.nsw = NSW;
#define UP(field, vec) do { NLS_CUDA(h, cudaMalloc((void **)&g.field, sizeof(vec[0]) * vec.size())); \
  NLS_CUDA(h, cudaMemcpyAsync(g.field, vec.data(), sizeof(vec[0]) * vec.size(), cudaMemcpyHostToDevice, h->stream)); } while (0)
  UP(tn, tn); UP(lconn, lconn); UP(cn, cn); UP(b0, b0); UP(deg, deg); UP(btask, btask); UP(bsrc, bsrc); UP(rsrc, rsrc);
#undef UP
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_plan = true;
  return NLS_OK;
}

int nls_launch_gather_2d(nls_context *h, bool want_r, bool want_k, int64_t c_begin, int64_t c_end) {
  AsmArgs a{};
  a.nel = h->nel; a.n_owned = h->n_owned; a.conn = h->d_conn; a.coords = h->d_coords; a.U = h->d_U;
  a.R = h->d_R; a.vals = h->d_vals; a.browptr = h->d_browptr; a.pos = h->d_pos;
  const GatherPlanDev &g = h->gplan;
  if (c_end < 0) c_end = g.ncl;
  if (c_end <= c_begin) return NLS_OK;
  GatherArgs p{c_end, c_begin, g.maxt, g.maxe, g.maxb, g.nsw, g.tn, g.lconn, g.cn, g.b0, g.deg, g.btask, g.bsrc, g.rsrc,
               h->opt_profile_phases ? reinterpret_cast<unsigned long long *>(h->d_prof) : nullptr};
  const size_t shm = gather_smem(g.maxt, g.maxe, g.maxb, g.nsw);
  if (!h->attr_gather) {       // per device: raised once per handle to the opt-in maximum
    cudaFuncSetAttribute(k_asm_q1_2d_gather<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    cudaFuncSetAttribute(k_asm_q1_2d_gather<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    cudaFuncSetAttribute(k_asm_q1_2d_gather<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_optin_smem);
    h->attr_gather = true;
  }
  // persistent CTAs: as many per SM as the shared memory admits (3 for <= 4 elements around a node, else 2)
  const int fit = (int)std::min<size_t>((want_r && want_k) ? 2 : 3, ((size_t)h->max_optin_smem + 1024) / (shm + 1024));
  const int per_sm = h->opt_gather_ctas > 0 ? h->opt_gather_ctas : std::max(1, fit);
  const unsigned grid = (unsigned)std::min<int64_t>(c_end - c_begin, (int64_t)h->sm_count * per_sm);
  if (want_r && want_k) k_asm_q1_2d_gather<true, true><<<grid, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  else if (want_r)      k_asm_q1_2d_gather<true, false><<<grid, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  else                  k_asm_q1_2d_gather<false, true><<<grid, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("gather kernel launch: ") + cudaGetErrorString(e); return NLS_ERR_CUDA; }
  h->launches++;
  return NLS_OK;
}
