// kernels_gather.cu -- gather ("row-owner") assembly of the 2-D Q1 Neo-Hookean residual and
// CSR Jacobian: no atomics, no zero-fill, every CSR row and R entry written exactly once, and
// contributions summed in ascending element order -- the order of the reference's serial element
// loop (compute_dinternalforce, src/materials/defaults.jl:77-81 + assemble!, src/fem/element.jl:70),
// so results are bit-reproducible run to run.
//
// One CTA per node cluster (host plan: nls_host_cluster_plan, 8x8-node tiles on structured grids,
// Morton tiles of rank-binned coordinates in general).
//   phase 1a  one thread per (element touching the cluster, quadrature point): J, dN/dX, F, F^-1,
//             lnJ, P and the spatial gradients g = F^-T dN/dX -> shared memory (24 doubles per qp)
//   phase 1b  one thread per (element, local node a) whose node belongs to the cluster: the row
//             block a of K_e (2 x 8) and f_a, accumulated over the 4 qps from shared memory.
//             Lanes of one element are adjacent, so the G_b/g_b reads are warp broadcasts.
//   phase 2   conflict-free accumulation into per-node row buffers in shared memory, in rounds:
//             round j adds, for every node, the contribution of its j-th adjacent element.
//   phase 3   rows copied to the CSR values / R.
// Why: ncu on the atomic-scatter kernel (profiles/r1a_ncu_asm2d_atomic.txt) shows 252 registers,
// 11% warps active, FP64 pipe 20%, 34 M RED sectors and 295 MB of DRAM traffic for 176 MB of
// algorithmic bytes; see DESIGN.md section 5.
#include <algorithm>
#include "asm_common.cuh"

#define G2_NC 64         // nodes per cluster
#define G2_THREADS 256   // >= items per cluster (4 per node on quad meshes)
#define G2_QS 24         // doubles per (element, qp) record: {G_a[2], g_a[2]} x 4 nodes, c1, c2, c3, wd*P[4], pad
#define G2_ES (4 * G2_QS + 2)   // doubles per element: 49 x 16 B (odd) -> lanes reading different elements hit different banks

struct PlanArgs {
  const int32_t *nodes, *eptr, *iptr;
  const int4 *conn;        // connectivity of the cluster's elements, contiguous per cluster
  const uint32_t *items;   // slot | a << 12 | node_local << 16 | rank << 24, sorted by (slot, a)
  const uint2 *ipos;       // per item: CSR block positions of the element's 4 nodes in the row of node a (4 x uint16)
  int max_elems, rs, rp, max_adj;
};

template <bool DO_R, bool DO_K>
__global__ void __launch_bounds__(G2_THREADS, 3)
k_asm_q1_2d_gather(const AsmArgs A, const PlanArgs P, const __grid_constant__ TabQ1 T,
                   const __grid_constant__ MatParams mp) {
  extern __shared__ __align__(16) double smem[];
  double *Q = smem;                                         // [max_elems] x G2_ES: phase 1a -> 1b
  double *racc = smem;                                      // [G2_NC][rp] row buffers, ALIASES Q (phase 2 -> 3)
  double *rres = smem + max((size_t)P.max_elems * G2_ES, (size_t)G2_NC * P.rp);   // [2*G2_NC]
  double *Tsm = rres + 2 * G2_NC;                           // [4][4][2] dN + [4] w
  const int tid = threadIdx.x;
  const int c = blockIdx.x;
  const int e0 = P.eptr[c], ne = P.eptr[c + 1] - e0;
  const int it0 = P.iptr[c], nit = P.iptr[c + 1] - it0;
  // issue every plan load up front so their latency overlaps phase 1a
  const bool have = tid < nit;
  unsigned it = 0;
  uint2 pp = make_uint2(0, 0);
  if (have) { it = P.items[it0 + tid]; if (DO_K) pp = P.ipos[it0 + tid]; }
  int my_node = -1, my_deg = 0;
  int64_t my_b0 = 0;
  if (tid < 2 * G2_NC) {
    my_node = P.nodes[c * G2_NC + (tid >> 1)];
    if (my_node >= 0 && DO_K) { my_b0 = A.browptr[my_node]; my_deg = (int)(A.browptr[my_node + 1] - my_b0); }
  }
  if (tid < 32) Tsm[tid] = T.dN[tid >> 3][(tid >> 1) & 3][tid & 1];
  else if (tid < 36) Tsm[tid] = T.w[tid - 32];
  if (tid < 2 * G2_NC) rres[tid] = 0.0;
  __syncthreads();

  const double mu = mp.p[0], lam = mp.p[1];
  // ---- phase 1a: geometry + kinematics of one quadrature point per thread ----
  for (int w = tid; w < ne * 4; w += G2_THREADS) {
    const int s = w >> 2, q = w & 3;
    const int4 c4 = P.conn[e0 + s];
    const int nd[4] = {c4.x, c4.y, c4.z, c4.w};
    double J[4] = {0, 0, 0, 0}, Ji[4], F[4] = {1, 0, 0, 1}, Fi[4], G[4][2], u[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double2 x = reinterpret_cast<const double2 *>(A.coords)[nd[a]];
      const double2 v = reinterpret_cast<const double2 *>(A.U)[nd[a]];
      u[a][0] = v.x; u[a][1] = v.y;
      const double d0 = Tsm[q * 8 + a * 2], d1 = Tsm[q * 8 + a * 2 + 1];
      G[a][0] = d0; G[a][1] = d1;
      J[0] = fma(x.x, d0, J[0]); J[1] = fma(x.x, d1, J[1]); J[2] = fma(x.y, d0, J[2]); J[3] = fma(x.y, d1, J[3]);
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
    const double wd = Tsm[32 + q] * detJ;
    double2 *rec = reinterpret_cast<double2 *>(Q + (size_t)s * G2_ES + q * G2_QS);
#pragma unroll
    for (int a = 0; a < 4; ++a) {   // per node: {G_a, g_a} = 32 B, so the lane-dependent reads of phase 1b do not conflict
      rec[2 * a] = make_double2(G[a][0], G[a][1]);
      rec[2 * a + 1] = make_double2(fma(G[a][0], Fi[0], G[a][1] * Fi[2]), fma(G[a][0], Fi[1], G[a][1] * Fi[3]));
    }
    const double cl = lam * lnJ, wmu = wd * mu, wcl = wd * cl;
    rec[8] = make_double2(wmu, wd * lam);
    rec[9] = make_double2(wmu - wcl, fma(wmu, F[0] - Fi[0], wcl * Fi[0]));                     // c3, P00
    rec[10] = make_double2(fma(wmu, F[1] - Fi[2], wcl * Fi[2]), fma(wmu, F[2] - Fi[1], wcl * Fi[1]));  // P01, P10
    rec[11] = make_double2(fma(wmu, F[3] - Fi[3], wcl * Fi[3]), 0.0);                           // P11
  }
  __syncthreads();

  // ---- phase 1b: row block a of K_e and f_a for this thread's (element, local node) item ----
  double K[4][2][2], f[2] = {0.0, 0.0};
  const int s = it & 0xfff, a = (it >> 12) & 7;
  const int nl = (it >> 16) & 0xff, rank = have ? (int)(it >> 24) : -1;
#pragma unroll
  for (int b = 0; b < 4; ++b) { K[b][0][0] = K[b][0][1] = K[b][1][0] = K[b][1][1] = 0.0; }
  if (have) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 *rec = reinterpret_cast<const double2 *>(Q + (size_t)s * G2_ES + q * G2_QS);
      const double2 Ga = rec[2 * a], ga = rec[2 * a + 1];
      const double2 c12 = rec[8], c3p = rec[9];
      if (DO_R) {
        const double2 p1 = rec[10], p2 = rec[11];
        f[0] = fma(c3p.y, Ga.x, fma(p1.x, Ga.y, f[0]));     // P00 G0 + P01 G1
        f[1] = fma(p1.y, Ga.x, fma(p2.x, Ga.y, f[1]));      // P10 G0 + P11 G1
      }
      if (DO_K) {
        // K_{ai,bk} += c1 (G_a.G_b) d_ik + c2 g_ai g_bk + c3 g_ak g_bi  -- 10 FMAs per b
        const double c1Gx = c12.x * Ga.x, c1Gy = c12.x * Ga.y;
        const double c2g0 = c12.y * ga.x, c2g1 = c12.y * ga.y;
        const double c3g0 = c3p.x * ga.x, c3g1 = c3p.x * ga.y;
        const double d0 = c2g0 + c3g0, d1 = c2g1 + c3g1;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double2 Gb = rec[2 * b], gb = rec[2 * b + 1];
          K[b][0][0] = fma(d0, gb.x, fma(c1Gx, Gb.x, fma(c1Gy, Gb.y, K[b][0][0])));
          K[b][0][1] = fma(c2g0, gb.y, fma(c3g1, gb.x, K[b][0][1]));
          K[b][1][0] = fma(c2g1, gb.x, fma(c3g0, gb.y, K[b][1][0]));
          K[b][1][1] = fma(d1, gb.y, fma(c1Gx, Gb.x, fma(c1Gy, Gb.y, K[b][1][1])));
        }
      }
    }
  }
  __syncthreads();                       // everyone is done reading Q: its storage becomes the row buffers
  if (DO_K) for (int k = tid; k < G2_NC * P.rp; k += G2_THREADS) racc[k] = 0.0;
  __syncthreads();
  // ---- phase 2: round j adds every node's j-th adjacent element (ascending element id) ----
  const int pb[4] = {(int)(pp.x & 0xffff), (int)(pp.x >> 16), (int)(pp.y & 0xffff), (int)(pp.y >> 16)};
  for (int j = 0; j < P.max_adj; ++j) {
    if (rank == j) {
      if (DO_K) {
        double *r0 = racc + (size_t)nl * P.rp, *r1 = r0 + P.rs;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          r0[2 * pb[b]] += K[b][0][0]; r0[2 * pb[b] + 1] += K[b][0][1];
          r1[2 * pb[b]] += K[b][1][0]; r1[2 * pb[b] + 1] += K[b][1][1];
        }
      }
      if (DO_R) { rres[nl * 2] += f[0]; rres[nl * 2 + 1] += f[1]; }
    }
    __syncthreads();
  }
  // ---- phase 3: write each row once ----
  if (my_node >= 0) {
    const int i = tid & 1;
    if (DO_K) {
      double *row = A.vals + 4 * my_b0 + (int64_t)i * 2 * my_deg;
      const double *acc = racc + (size_t)(tid >> 1) * P.rp + i * P.rs;
      for (int j = 0; j < 2 * my_deg; ++j) row[j] = acc[j];
    }
    if (DO_R) A.R[2 * (int64_t)my_node + i] = rres[tid];
  }
}

// ---- host side ---------------------------------------------------------------------------------
bool nls_assembly_overwrites(const nls_context *h) { return h->dim == 2 && h->opt_assembly == 1 && h->have_plan; }
int nls_gather_nc(void) { return G2_NC; }
int nls_gather_max_items(void) { return G2_THREADS; }
size_t nls_gather_smem(int max_elems, int maxdeg) {
  const size_t rp = (size_t)(4 * maxdeg + 1);
  return sizeof(double) * (std::max((size_t)max_elems * G2_ES, (size_t)G2_NC * rp) + 2 * G2_NC + 40);
}

int nls_launch_gather_2d(nls_context *h, bool want_r, bool want_k) {
  AsmArgs a{};
  a.nel = h->nel; a.n_owned = h->n_owned; a.conn = h->d_conn; a.coords = h->d_coords; a.U = h->d_U;
  a.R = h->d_R; a.vals = h->d_vals; a.browptr = h->d_browptr; a.pos = h->d_pos;
  PlanArgs p{h->d_cl_nodes, h->d_cl_eptr, h->d_cl_iptr, reinterpret_cast<const int4 *>(h->d_cl_conn), h->d_cl_items,
             reinterpret_cast<const uint2 *>(h->d_cl_ipos), h->plan_max_elems, 2 * h->plan_maxdeg, 4 * h->plan_maxdeg + 1, h->plan_max_adj};
  const size_t shm = nls_gather_smem(h->plan_max_elems, h->plan_maxdeg);
  static size_t shm_set = 0;
  if (shm > shm_set) {
    cudaFuncSetAttribute(k_asm_q1_2d_gather<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    cudaFuncSetAttribute(k_asm_q1_2d_gather<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    cudaFuncSetAttribute(k_asm_q1_2d_gather<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    shm_set = shm;
  }
  const unsigned g = (unsigned)h->plan_ncl;
  if (want_r && want_k) k_asm_q1_2d_gather<true, true><<<g, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  else if (want_r)      k_asm_q1_2d_gather<true, false><<<g, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  else                  k_asm_q1_2d_gather<false, true><<<g, G2_THREADS, shm, h->stream>>>(a, p, h->tabq, h->mat);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("gather kernel launch: ") + cudaGetErrorString(e); return NLS_ERR_CUDA; }
  h->launches++;
  return NLS_OK;
}
