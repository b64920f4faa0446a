// kernels_gather.cu -- gather ("row-owner") assembly of the 2-D Q1 Neo-Hookean residual and CSR
// Jacobian: no atomics, no zero-fill, every CSR row and R entry written exactly once, contributions
// summed in ascending element order -- the order of the reference's serial element loop
// (compute_dinternalforce, src/materials/defaults.jl:77-81 + assemble!, src/fem/element.jl:70) --
// so results are bit-reproducible run to run.
//
// Persistent CTAs walk node clusters (host plan: 8x8-node Morton tiles, nls_host_cluster_plan).
// Everything a cluster needs is staged into shared memory with cp.async (LDGSTS) while the
// previous cluster is being computed, so no global-load latency sits on the critical path:
//   group A(k+2): the list of nodes touched by cluster k+2           (contiguous copy)
//   group B(k+1): coordinates + displacements of those nodes (16-byte gathers through the list),
//                 local connectivity, row owners, CSR row bases, gather source lists
//   compute(k)  : phase 1  one thread per element: K_e (symmetric half accumulated, 8x8 stored)
//                          and f_e over the 4 quadrature points -> shared memory
//                 phase 2  one thread per (owned node, component): walks the node's source list
//                          (element slot, local row, local column block -> CSR column block,
//                          sorted by column block then element) and writes the row once.
// History and counters behind this design: DESIGN.md section 5, profiles/r1a..r1e.
#include <algorithm>
#include <cstring>
#include "asm_common.cuh"

#define G2_NC 64          // owned nodes per cluster
#define G2_THREADS 128
#define G2_QS 24          // doubles per (element, quadrature point) record
#define G2_KS (4 * G2_QS + 2)   // doubles per element in shared memory: 4 records, later overwritten by the
                                // 10 symmetric 2x2 blocks + fe[8]; 49 x 16 B (odd stride -> conflict-free across elements)

struct GatherArgs {
  int64_t ncl;
  int maxt, maxe, ns, rp; // padded strides: touched nodes, elements, sources per node, row-buffer doubles per node
  const int32_t *tn;      // [ncl][4 + maxt]: {nt, ne, 0, 0}, node ids
  const uint32_t *lconn;  // [ncl][maxe]: 4 local node indices (bytes) into the touched-node list
  const int32_t *cn;      // [ncl][G2_NC] owned node ids (-1 = none)
  const int64_t *b0;      // [ncl][G2_NC] block-row start of the node
  const int32_t *deg;     // [ncl][G2_NC] block-row length
  const uint32_t *nsrc;   // [ncl][G2_NC][ns]: Ks offset (14 bits) | lower << 14 | diag << 15 | jb << 16, ~0 = end
  unsigned long long *prof; // optional per-phase cycle counters (nls_set_option "profile_phases"), else NULL
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// phase 1a: one quadrature point. Record (24 doubles): {G_a[2], g_a[2]} x 4 nodes, c1, c2, c3, wd*P[4], pad.
template <bool DO_R, bool DO_K>
__device__ __forceinline__ void q1_2d_qp_record(const double (&X)[4][2], const double (&u)[4][2], const double *__restrict__ Tq,
                                                const double w, const double mu, const double lam, double2 *__restrict__ rec) {
  double J[4] = {0, 0, 0, 0}, Ji[4], F[4] = {1, 0, 0, 1}, Fi[4], G[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double d0 = Tq[a * 2], d1 = Tq[a * 2 + 1];
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
  const double wd = w * detJ;
  const double wmu = wd * mu, wcl = wd * lam * lnJ;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    rec[2 * a] = make_double2(G[a][0], G[a][1]);
    rec[2 * a + 1] = make_double2(fma(G[a][0], Fi[0], G[a][1] * Fi[2]), fma(G[a][0], Fi[1], G[a][1] * Fi[3]));
  }
  rec[8] = make_double2(wmu, wd * lam);
  rec[9] = make_double2(wmu - wcl, fma(wmu, F[0] - Fi[0], wcl * Fi[0]));                        // c3, P00
  rec[10] = make_double2(fma(wmu, F[1] - Fi[2], wcl * Fi[2]), fma(wmu, F[2] - Fi[1], wcl * Fi[1]));   // P01, P10
  rec[11] = make_double2(fma(wmu, F[3] - Fi[3], wcl * Fi[3]), 0.0);                             // P11
}

// phase 1b: K_e (symmetric half) and f_e of one element from its 4 records; result overwrites the
// start of the element's own record area (thread-private, so no synchronisation is needed):
// blocks (a,b), a <= b, in the order (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3), each [i][k]
// row-major; then fe[8].
template <bool DO_R, bool DO_K>
__device__ __forceinline__ void q1_2d_accumulate_sym(double *__restrict__ el) {
  double fe[8], Kd[4][4], Ko[6][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) fe[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Kd[i][j] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ko[i][j] = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double2 *rec = reinterpret_cast<const double2 *>(el + q * G2_QS);
    double2 G[4], g[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { G[a] = rec[2 * a]; g[a] = rec[2 * a + 1]; }
    const double2 c12 = rec[8], c3p = rec[9];
    if (DO_R) {
      const double2 p1 = rec[10], p2 = rec[11];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        fe[a * 2] = fma(c3p.y, G[a].x, fma(p1.x, G[a].y, fe[a * 2]));
        fe[a * 2 + 1] = fma(p1.y, G[a].x, fma(p2.x, G[a].y, fe[a * 2 + 1]));
      }
    }
    if (DO_K) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double c1Gx = c12.x * G[a].x, c1Gy = c12.x * G[a].y;
        const double c2g0 = c12.y * g[a].x, c2g1 = c12.y * g[a].y, c3g0 = c3p.x * g[a].x, c3g1 = c3p.x * g[a].y;
        const double d0 = c2g0 + c3g0, d1 = c2g1 + c3g1;
        Kd[a][0] = fma(d0, g[a].x, fma(c1Gx, G[a].x, fma(c1Gy, G[a].y, Kd[a][0])));
        Kd[a][1] = fma(c2g0, g[a].y, fma(c3g1, g[a].x, Kd[a][1]));
        Kd[a][2] = fma(c2g1, g[a].x, fma(c3g0, g[a].y, Kd[a][2]));
        Kd[a][3] = fma(d1, g[a].y, fma(c1Gx, G[a].x, fma(c1Gy, G[a].y, Kd[a][3])));
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
          const int ob = (a == 0) ? (b - 1) : (a == 1 ? b + 1 : 5);   // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) -> 0..5
          Ko[ob][0] = fma(d0, g[b].x, fma(c1Gx, G[b].x, fma(c1Gy, G[b].y, Ko[ob][0])));
          Ko[ob][1] = fma(c2g0, g[b].y, fma(c3g1, g[b].x, Ko[ob][1]));
          Ko[ob][2] = fma(c2g1, g[b].x, fma(c3g0, g[b].y, Ko[ob][2]));
          Ko[ob][3] = fma(d1, g[b].y, fma(c1Gx, G[b].x, fma(c1Gy, G[b].y, Ko[ob][3])));
        }
      }
    }
  }
  double2 *d2 = reinterpret_cast<double2 *>(el);
  if (DO_K) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int bd = (a == 0) ? 0 : (a == 1 ? 4 : (a == 2 ? 7 : 9));
      d2[bd * 2] = make_double2(Kd[a][0], Kd[a][1]);
      d2[bd * 2 + 1] = make_double2(Kd[a][2], Kd[a][3]);
#pragma unroll
      for (int b = a + 1; b < 4; ++b) {
        const int ob = (a == 0) ? (b - 1) : (a == 1 ? b + 1 : 5);
        const int bo = bd + (b - a);
        d2[bo * 2] = make_double2(Ko[ob][0], Ko[ob][1]);
        d2[bo * 2 + 1] = make_double2(Ko[ob][2], Ko[ob][3]);
      }
    }
  }
  if (DO_R) {
#pragma unroll
    for (int a = 0; a < 4; ++a) d2[20 + a] = make_double2(fe[a * 2], fe[a * 2 + 1]);
  }
}

template <bool DO_R, bool DO_K>
__global__ void __launch_bounds__(G2_THREADS, 2)
k_asm_q1_2d_gather(const AsmArgs A, const GatherArgs P, const __grid_constant__ TabQ1 T,
                   const __grid_constant__ MatParams mp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- shared-memory carve-up (all sub-buffers 16-byte aligned) ----
  const int tn_words = 4 + P.maxt;
  double *Qs = reinterpret_cast<double *>(smem_raw);                                  // [maxe][G2_KS]
  double *Tsm = Qs + (size_t)P.maxe * G2_KS + 4;                                      // 4 zero doubles (dummy source), then 40 doubles
  int32_t *tnb = reinterpret_cast<int32_t *>(Tsm + 40);                               // [3][tn_words]
  unsigned char *bbase = reinterpret_cast<unsigned char *>(tnb + 3 * tn_words);      // [2] stage buffers
  const size_t b_xu = 0, b_b0 = b_xu + (size_t)P.maxt * 32, b_lc = b_b0 + G2_NC * 8, b_cn = b_lc + (size_t)P.maxe * 4,
               b_dg = b_cn + G2_NC * 4, b_ns = b_dg + G2_NC * 4, b_size = b_ns + (size_t)G2_NC * P.ns * 4;
  const int tid = threadIdx.x;
  if (tid < 32) Tsm[tid] = T.dN[tid >> 3][(tid >> 1) & 3][tid & 1];
  else if (tid < 36) Tsm[tid] = T.w[tid - 32];
  else if (tid < 40) Tsm[tid - 40] = 0.0;   // the zero slot just below Tsm

  auto issue_A = [&](int64_t c, int k) {          // node list of cluster c -> tnb[k % 3]
    if (c < P.ncl) {
      const int32_t *src = P.tn + c * tn_words;
      int32_t *dst = tnb + (k % 3) * tn_words;
      for (int w = tid * 4; w < tn_words; w += G2_THREADS * 4) cp_async16(dst + w, src + w);
    }
    cp_async_commit();
  };
  auto issue_B = [&](int64_t c, int k) {          // everything else of cluster c -> stage k & 1
    if (c < P.ncl) {
      unsigned char *B = bbase + (size_t)(k & 1) * b_size;
      const int32_t *tn = tnb + (k % 3) * tn_words;
      const int nt = tn[0];
      double *xu = reinterpret_cast<double *>(B + b_xu);
      for (int t = tid; t < nt; t += G2_THREADS) {
        const int id = tn[4 + t];
        cp_async16(xu + 4 * t, A.coords + 2 * (int64_t)id);
        cp_async16(xu + 4 * t + 2, A.U + 2 * (int64_t)id);
      }
      for (int w = tid * 2; w < G2_NC; w += G2_THREADS * 2) cp_async16(B + b_b0 + w * 8, P.b0 + c * G2_NC + w);
      for (int w = tid * 4; w < P.maxe; w += G2_THREADS * 4) cp_async16(B + b_lc + w * 4, P.lconn + c * P.maxe + w);
      for (int w = tid * 4; w < G2_NC; w += G2_THREADS * 4) {
        cp_async16(B + b_cn + w * 4, P.cn + c * G2_NC + w);
        cp_async16(B + b_dg + w * 4, P.deg + c * G2_NC + w);
      }
      const int nsw = G2_NC * P.ns;   // uint32 entries, multiple of 4
      for (int w = tid * 4; w < nsw; w += G2_THREADS * 4) cp_async16(B + b_ns + w * 4, P.nsrc + c * nsw + w);
    }
    cp_async_commit();
  };

  const double mu = mp.p[0], lam = mp.p[1];
  const int64_t c0 = blockIdx.x, cstep = gridDim.x;
  issue_A(c0, 0);
  issue_A(c0 + cstep, 1);
  cp_async_wait<1>();            // A(0) landed
  __syncthreads();
  issue_B(c0, 0);

  int k = 0;
  for (int64_t c = c0; c < P.ncl; c += cstep, ++k) {
    const long long tp0 = P.prof ? clock64() : 0;
    issue_A(c + 2 * cstep, k + 2);
    cp_async_wait<1>();          // everything but A(k+2): A(k+1) and B(k) have landed
    __syncthreads();
    issue_B(c + cstep, k + 1);   // flies while cluster k is computed
    const long long tp1 = P.prof ? clock64() : 0;

    const unsigned char *B = bbase + (size_t)(k & 1) * b_size;
    const int ne = tnb[(k % 3) * tn_words + 1];
    // ---- phase 1a: one thread per (element, quadrature point): the latency-bound geometry chain
    //      (J, J^-1, F, F^-1, log) gets 4x the thread-level parallelism of a per-element thread ----
    {
      const double *xu = reinterpret_cast<const double *>(B + b_xu);
      const unsigned *lcv = reinterpret_cast<const unsigned *>(B + b_lc);
      for (int w = tid; w < ne * 4; w += G2_THREADS) {
        const int sl = w >> 2, q = w & 3;
        const unsigned lc = lcv[sl];
        double X[4][2], u[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int t = (lc >> (8 * a)) & 0xff;
          const double2 x = reinterpret_cast<const double2 *>(xu)[2 * t];
          const double2 v = reinterpret_cast<const double2 *>(xu)[2 * t + 1];
          X[a][0] = x.x; X[a][1] = x.y; u[a][0] = v.x; u[a][1] = v.y;
        }
        q1_2d_qp_record<DO_R, DO_K>(X, u, Tsm + q * 8, Tsm[32 + q], mu, lam,
                                    reinterpret_cast<double2 *>(Qs + (size_t)sl * G2_KS + q * G2_QS));
      }
    }
    __syncthreads();
    const long long tp2 = P.prof ? clock64() : 0;
    // ---- phase 1b: one thread per element: pure FMA stream over its 4 records ----
    if (tid < ne) q1_2d_accumulate_sym<DO_R, DO_K>(Qs + (size_t)tid * G2_KS);
    __syncthreads();
    const long long tp3 = P.prof ? clock64() : 0;
    // ---- phase 2: one thread per (owned node, component) writes its row ----
    {
      const int nl = tid >> 1, i = tid & 1;
      const int node = reinterpret_cast<const int32_t *>(B + b_cn)[nl];
      if (node >= 0) {
        const int64_t b0 = reinterpret_cast<const int64_t *>(B + b_b0)[nl];
        const int deg = reinterpret_cast<const int32_t *>(B + b_dg)[nl];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(B + b_ns) + (size_t)nl * P.ns;
        double *row = A.vals + 4 * b0 + (int64_t)i * 2 * deg;
        double rsum = 0.0;
        double2 acc = make_double2(0.0, 0.0);
        // branch-free walk: every list has exactly P.ns entries (padded with sources that point at the
        // zero slot); bit 14 = block stored transposed, bit 15 = diagonal block (adds f_e), bit 26 = last
        // source of its column block (flush)
#pragma unroll 4
        for (int m = 0; m < P.ns; ++m) {
          const uint32_t code = src[m];
          const int off = code & 0x3fff, tr = (code >> 14) & 1;
          if (DO_K) {
            const int xo = off + i * (2 - tr);
            acc.x += Qs[xo]; acc.y += Qs[xo + 1 + tr];
            if (code & (1u << 26)) {
              *reinterpret_cast<double2 *>(row + 2 * ((code >> 16) & 0xff)) = acc;
              acc = make_double2(0.0, 0.0);
            }
          }
          if (DO_R && (code & 0x8000u)) rsum += Qs[(off / G2_KS) * G2_KS + 40 + ((code >> 24) & 3) * 2 + i];
        }
        if (DO_R) A.R[2 * (int64_t)node + i] = rsum;
      }
    }
    __syncthreads();             // Qs and stage k & 1 are free again
    if (P.prof && tid == 0) {
      const long long tp4 = clock64();
      atomicAdd(P.prof + 0, (unsigned long long)(tp1 - tp0));   // staging wait + issue
      atomicAdd(P.prof + 1, (unsigned long long)(tp2 - tp1));   // phase 1a
      atomicAdd(P.prof + 2, (unsigned long long)(tp3 - tp2));   // phase 1b
      atomicAdd(P.prof + 3, (unsigned long long)(tp4 - tp3));   // phase 2
      atomicAdd(P.prof + 4, 1ull);                               // clusters
    }
  }
  cp_async_wait<0>();
}

// ---- host side: pack the cluster plan into fixed-stride device arrays ---------------------------
bool nls_assembly_overwrites(const nls_context *h) { return h->dim == 2 && h->opt_assembly == 1 && h->have_plan; }
int nls_gather_nc(void) { return G2_NC; }

static size_t gather_smem(int maxt, int maxe, int ns, int rp) {
  const size_t bsz = (size_t)maxt * 32 + G2_NC * 8 + (size_t)maxe * 4 + G2_NC * 4 + G2_NC * 4 + (size_t)G2_NC * ns * 4;
  (void)rp;   // staging the rows in shared memory for coalesced stores was measured slower
  return (size_t)maxe * G2_KS * 8 + 44 * 8 + (size_t)3 * (4 + maxt) * 4 + 2 * bsz;
}

void nls_gather_free(nls_context *h) {
  GatherPlanDev &g = h->gplan;
  if (g.tn) cudaFree(g.tn);
  if (g.lconn) cudaFree(g.lconn);
  if (g.cn) cudaFree(g.cn);
  if (g.b0) cudaFree(g.b0);
  if (g.deg) cudaFree(g.deg);
  if (g.nsrc) cudaFree(g.nsrc);
  g = GatherPlanDev();
  h->have_plan = false;
}

// Sets h->have_plan when the mesh fits the kernel's packed formats; otherwise leaves it false and
// the atomic-scatter kernel is used instead.
int nls_gather_build(nls_context *h, const AsmPlan &plan) {
  nls_gather_free(h);
  if (h->dim != 2 || plan.ncl == 0) return NLS_OK;
  const int nen = 4;
  const int64_t ncl = plan.ncl;
  const int32_t *conn = h->h_conn.data();
  int maxdeg = 0;
  for (int64_t a = 0; a < h->n_owned; ++a) maxdeg = std::max<int>(maxdeg, (int)(h->h_browptr[a + 1] - h->h_browptr[a]));
  if (plan.max_elems > G2_THREADS || (size_t)(plan.max_elems + 4) * G2_KS + 4 > 0x3fff || maxdeg > 255 || plan.max_adj > 8) return NLS_OK;
  std::vector<std::vector<int32_t>> touched((size_t)ncl);
  int maxt = 0;
  for (int64_t c = 0; c < ncl; ++c) {
    std::vector<int32_t> &t = touched[c];
    for (int32_t s = plan.eptr[c]; s < plan.eptr[c + 1]; ++s)
      t.insert(t.end(), conn + (int64_t)plan.elems[s] * nen, conn + (int64_t)(plan.elems[s] + 1) * nen);
    std::sort(t.begin(), t.end());
    t.erase(std::unique(t.begin(), t.end()), t.end());
    maxt = std::max<int>(maxt, (int)t.size());
  }
  if (maxt > 255) return NLS_OK;
  const int MAXT = (maxt + 3) & ~3, MAXE = (plan.max_elems + 3) & ~3, NS = (plan.max_adj * nen + 3) & ~3;
  const int RP = 4 * maxdeg + 2;   // doubles per node in the row buffer; (2 maxdeg + 1) x 16 B = odd
  if (gather_smem(MAXT, MAXE, NS, RP) > (size_t)110 * 1024) return NLS_OK;
  const int tnw = 4 + MAXT;
  std::vector<int32_t> tn((size_t)ncl * tnw, 0), cn((size_t)ncl * G2_NC, -1), deg((size_t)ncl * G2_NC, 0);
  std::vector<uint32_t> lconn((size_t)ncl * MAXE, 0);
  std::vector<int64_t> b0((size_t)ncl * G2_NC, 0);
  const uint32_t dummy = (uint32_t)(MAXE * G2_KS);   // offset of the zero slot, no flags
  std::vector<uint32_t> nsrc((size_t)ncl * G2_NC * NS, dummy);
  std::vector<uint32_t> codes;
  for (int64_t c = 0; c < ncl; ++c) {
    const std::vector<int32_t> &t = touched[c];
    const int32_t e0 = plan.eptr[c], ne = plan.eptr[c + 1] - e0;
    tn[c * tnw] = (int32_t)t.size(); tn[c * tnw + 1] = ne;
    std::copy(t.begin(), t.end(), tn.begin() + c * tnw + 4);
    for (int32_t s = 0; s < ne; ++s) {
      uint32_t packed = 0;
      for (int a = 0; a < nen; ++a) {
        const int32_t nd = conn[(int64_t)plan.elems[e0 + s] * nen + a];
        packed |= (uint32_t)(std::lower_bound(t.begin(), t.end(), nd) - t.begin()) << (8 * a);
      }
      lconn[c * MAXE + s] = packed;
    }
    for (int nl = 0; nl < G2_NC; ++nl) {
      const int32_t a = plan.nodes[c * G2_NC + nl];
      if (a < 0) continue;
      cn[c * G2_NC + nl] = a;
      b0[c * G2_NC + nl] = h->h_browptr[a];
      deg[c * G2_NC + nl] = (int32_t)(h->h_browptr[a + 1] - h->h_browptr[a]);
      const int32_t *lo = h->h_bcol.data() + h->h_browptr[a], *hi = h->h_bcol.data() + h->h_browptr[a + 1];
      codes.clear();
      for (int32_t s = 0; s < ne; ++s) {          // ascending slot = ascending element id
        const int32_t *ce = conn + (int64_t)plan.elems[e0 + s] * nen;
        int al = -1;
        for (int q = 0; q < nen; ++q) if (ce[q] == a) al = q;
        if (al < 0) continue;
        for (int b = 0; b < nen; ++b) {
          const uint32_t jb = (uint32_t)(std::lower_bound(lo, hi, ce[b]) - lo);
          codes.push_back((jb << 16) | ((uint32_t)s << 4) | ((uint32_t)al << 2) | (uint32_t)b);   // sort key: (jb, slot)
        }
      }
      std::sort(codes.begin(), codes.end());
      if ((int)codes.size() > NS) return NLS_OK;   // cannot happen: NS >= max_adj * nen
      for (size_t m = 0; m < codes.size(); ++m) {
        const uint32_t kk = codes[m], jb = kk >> 16, s = (kk >> 4) & 0xfff, al = (kk >> 2) & 3, b = kk & 3;
        static const int bd[4] = {0, 4, 7, 9};           // first symmetric block of local row a
        const uint32_t lo_ = std::min(al, b), hi_ = std::max(al, b);
        const uint32_t off = s * G2_KS + (uint32_t)(bd[lo_] + (hi_ - lo_)) * 4;
        const bool last = (m + 1 == codes.size()) || ((codes[m + 1] >> 16) != jb);
        nsrc[((size_t)c * G2_NC + nl) * NS + m] = off | (al > b ? 0x4000u : 0u) | (al == b ? 0x8000u : 0u) | (jb << 16) | (al << 24) |
                                                   (last ? (1u << 26) : 0u);
      }
    }
  }
  GatherPlanDev &g = h->gplan;
  g.ncl = ncl; g.maxt = MAXT; g.maxe = MAXE; g.ns = NS; g.rp = RP;
#define UP(field, vec) do { NLS_CUDA(h, cudaMalloc((void **)&g.field, sizeof(vec[0]) * vec.size())); \
  NLS_CUDA(h, cudaMemcpyAsync(g.field, vec.data(), sizeof(vec[0]) * vec.size(), cudaMemcpyHostToDevice, h->stream)); } while (0)
  UP(tn, tn); UP(lconn, lconn); UP(cn, cn); UP(b0, b0); UP(deg, deg); UP(nsrc, nsrc);
#undef UP
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_plan = true;
  return NLS_OK;
}

int nls_launch_gather_2d(nls_context *h, bool want_r, bool want_k) {
  AsmArgs a{};
  a.nel = h->nel; a.n_owned = h->n_owned; a.conn = h->d_conn; a.coords = h->d_coords; a.U = h->d_U;
  a.R = h->d_R; a.vals = h->d_vals; a.browptr = h->d_browptr; a.pos = h->d_pos;
  const GatherPlanDev &g = h->gplan;
  GatherArgs p{g.ncl, g.maxt, g.maxe, g.ns, g.rp, g.tn, g.lconn, g.cn, g.b0, g.deg, g.nsrc,
               h->opt_profile_phases ? reinterpret_cast<unsigned long long *>(h->d_prof) : nullptr};
  const size_t shm = gather_smem(g.maxt, g.maxe, g.ns, g.rp);
  static size_t shm_set = 0;
  if (shm > shm_set) {
    cudaFuncSetAttribute(k_asm_q1_2d_gather<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    cudaFuncSetAttribute(k_asm_q1_2d_gather<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    cudaFuncSetAttribute(k_asm_q1_2d_gather<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    shm_set = shm;
  }
  const int per_sm = h->opt_gather_ctas > 0 ? h->opt_gather_ctas : 2;
  const unsigned grid = (unsigned)std::min<int64_t>(g.ncl, (int64_t)h->sm_count * per_sm);
  if (want_r && want_k) k_asm_q1_2d_gather<true, true><<<grid, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  else if (want_r)      k_asm_q1_2d_gather<true, false><<<grid, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  else                  k_asm_q1_2d_gather<false, true><<<grid, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("gather kernel launch: ") + cudaGetErrorString(e); return NLS_ERR_CUDA; }
  h->launches++;
  return NLS_OK;
}
