// kernels_assembly.cu -- the element loop of NonlinearSolids.jl as sm_100a CUDA kernels.
//
// Replaces (paths relative to the reference source tree):
//   compute_internalforce / compute_dinternalforce   src/materials/defaults.jl:29-46, 68-82
//   _internalforce_el / _dinternalforce_el           src/materials/defaults.jl:2-26
//   plastic element kernels + update_state_el!       src/materials/linearelastoplasticity.jl:23-93
//   restrict! / unrestrict! / assemble!              src/fem/element.jl:64-70
//   Dirichlet / Neumann apply!                       src/fem/boundary.jl:29,39
// The 2-D / 3-D Q1 Neo-Hookean kernels are the builder-defined tensor-product extension
// (SURVEY.md 8c); they keep the same restrict -> qfunc -> integrate -> scatter structure.
//
// All arithmetic is FP64. Scatter is FP64 atomics (RED.E.ADD.F64) into R and the CSR
// values; element -> CSR slot comes from a precomputed per-element position map.
#include "nls_internal.h"
#include "asm_common.cuh"
#include "hex_element.cuh"

// ------------------------------------------------------------------------------------
// 1-D bar, order p = M-1 Lagrange on Chebyshev-2 nodes, nq Gauss points.
// One thread per element. Operation order follows defaults.jl:7-9,23 and
// quadrature.jl:69 (sum over q ascending) -- see SURVEY.md Appendix C.
// ------------------------------------------------------------------------------------
template <int M, bool DO_R, bool DO_K>
__global__ void __launch_bounds__(128)
k_asm_1d(const AsmArgs A, const __grid_constant__ Tab1D T, const __grid_constant__ MatParams mp) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.nel) return;
  int nd[M];
  double u[M], fe[M], Ke[M * M], B[M];
#pragma unroll
  for (int a = 0; a < M; ++a) { nd[a] = A.conn[e * M + a]; u[a] = A.U[nd[a]]; }
  const double dxdxi = (A.coords[nd[M - 1]] - A.coords[nd[0]]) / 2.0;  // length(el)/2, element.jl:15
  const double inv = 1.0 / dxdxi;
  const double area = mp.area;
  const int nq = T.nq;
  const bool plastic = mp.kind == NLS_MAT_ELASTOPLASTIC_1D;
  for (int q = 0; q < nq; ++q) {
#pragma unroll
    for (int i = 0; i < M; ++i) B[i] = T.dN[q][i] * inv;  // B = gradN' * (1/dxdxi)
    double s = 0.0, d = 0.0;
    if (plastic) {
      const int64_t k = e * nq + q;
      if (DO_R) {  // update_state_el!, linearelastoplasticity.jl:68-91 (residual pass only)
        double deps = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
          const double du = u[i] - A.u_n[e * M + i];
          deps = (i == 0) ? B[0] * du : deps + B[i] * du;
        }
        const double E = mp.p[0], Hk = mp.p[1], Ha = mp.p[2], Eep = mp.p[5];
        const double s_tr = A.st[5][k] + E * deps;
        const double rel = s_tr - A.st[6][k];
        const double sgn = rel > 0.0 ? 1.0 : (rel < 0.0 ? -1.0 : 0.0);
        const double f_tr = fabs(rel) - A.st[7][k];
        double alp, kap, gam;
        if (f_tr > 1e-5) {
          const double dg = f_tr / (E + Hk + Ha);
          s = s_tr - E * dg * sgn;
          alp = A.st[6][k] + Ha * dg * sgn;
          kap = A.st[7][k] + Hk * dg;
          gam = A.st[8][k] + dg;
          d = Eep;
        } else {
          s = s_tr; alp = A.st[6][k]; kap = A.st[7][k]; gam = A.st[8][k]; d = E;
        }
        A.st[0][k] = s; A.st[1][k] = alp; A.st[2][k] = kap; A.st[3][k] = gam; A.st[4][k] = d;
      } else {     // Jacobian pass reads the stored tangent, linearelastoplasticity.jl:42-47
        d = A.st[4][k];
      }
    } else {
      double eps = 0.0;  // eps = B * u
#pragma unroll
      for (int i = 0; i < M; ++i) eps = (i == 0) ? B[0] * u[0] : eps + B[i] * u[i];
      if (mp.kind == NLS_MAT_LINEAR_ELASTIC_1D) { s = mp.p[0] * eps; d = mp.p[0]; }
      else {  // exponential.jl:20-21
        const double ex = exp(-mp.p[1] * eps);
        s = mp.p[0] * (1.0 - ex);
        d = mp.p[1] * mp.p[0] * ex;
      }
    }
    const double w = T.w[q];
    if (DO_R) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const double v = ((B[i] * s) * area) * dxdxi;
        fe[i] = (q == 0) ? v * w : fe[i] + v * w;
      }
    }
    if (DO_K) {
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const double v = (((B[i] * d) * B[j]) * area) * dxdxi;
          Ke[i * M + j] = (q == 0) ? v * w : Ke[i * M + j] + v * w;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) {
    if (nd[i] >= A.n_owned) continue;
    if (DO_R) red_add(A.R + nd[i], fe[i]);
    if (DO_K) {
      const int64_t base = A.browptr[nd[i]];
#pragma unroll
      for (int j = 0; j < M; ++j) red_add(A.vals + base + A.pos[(e * M + i) * M + j], Ke[i * M + j]);
    }
  }
}

// ------------------------------------------------------------------------------------
// 2-D plane-strain Q1 compressible Neo-Hookean. One thread per element.
//   F = I + sum_a u_a (x) dN_a/dX,  P = mu (F - F^-T) + lam lnJ F^-T
//   K_{ai,bk} = w detJ [ mu d_ik (G_a.G_b) + lam g_ai g_bk + (mu - lam lnJ) g_ak g_bi ],
//   G_a = dN_a/dX,  g_a = F^-T G_a  (spatial form of the tangent of SURVEY.md 8c).
// ------------------------------------------------------------------------------------
// Element integrals of one quad: fe[8], Ke[8][8] (row-major DoF order a*2+i).
template <bool DO_R, bool DO_K>
__device__ __forceinline__ void q1_2d_element(const double (&X)[4][2], const double (&u)[4][2], const TabQ1 &T,
                                              const double mu, const double lam, double (&fe)[8], double (&Ke)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    fe[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) Ke[i][j] = 0.0;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double J[4] = {0, 0, 0, 0}, Ji[4], F[4] = {1, 0, 0, 1}, Fi[4], G[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int D = 0; D < 2; ++D) J[d * 2 + D] += X[a][d] * T.dN[q][a][D];
    const double detJ = inv2(J, Ji);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int D = 0; D < 2; ++D) G[a][D] = T.dN[q][a][0] * Ji[D] + T.dN[q][a][1] * Ji[2 + D];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int Jc = 0; Jc < 2; ++Jc) F[i * 2 + Jc] += u[a][i] * G[a][Jc];
    const double detF = inv2(F, Fi);
    const double lnJ = log(detF);
    const double wd = T.w[q] * detJ;
    if (DO_R) {
      double P[4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int Jc = 0; Jc < 2; ++Jc)
          P[i * 2 + Jc] = wd * (mu * (F[i * 2 + Jc] - Fi[Jc * 2 + i]) + lam * lnJ * Fi[Jc * 2 + i]);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 2; ++i) fe[a * 2 + i] += P[i * 2] * G[a][0] + P[i * 2 + 1] * G[a][1];
    }
    if (DO_K) {
      double g[4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 2; ++i) g[a][i] = G[a][0] * Fi[i] + G[a][1] * Fi[2 + i];
      const double c1 = wd * mu, c2 = wd * lam, c3 = wd * (mu - lam * lnJ);
      // K is symmetric: K[(b,k),(a,i)] = K[(a,i),(b,k)]; accumulate blocks b >= a only
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double c2g[2] = {c2 * g[a][0], c2 * g[a][1]};
        const double c3g[2] = {c3 * g[a][0], c3 * g[a][1]};
#pragma unroll
        for (int b = a; b < 4; ++b) {
          const double dab = c1 * (G[a][0] * G[b][0] + G[a][1] * G[b][1]);
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 2; ++k)
              Ke[a * 2 + i][b * 2 + k] += c2g[i] * g[b][k] + c3g[k] * g[b][i] + (i == k ? dab : 0.0);
        }
      }
    }
  }
  if (DO_K) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < r; ++c)
        if ((c >> 1) < (r >> 1)) Ke[r][c] = Ke[c][r];   // mirror the strictly-lower blocks
  }
}

template <bool DO_R, bool DO_K>
__global__ void __launch_bounds__(128)
k_asm_q1_2d(const AsmArgs A, const __grid_constant__ TabQ1 T, const __grid_constant__ MatParams mp) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.nel) return;
  const int4 c4 = reinterpret_cast<const int4 *>(A.conn)[e];
  const int nd[4] = {c4.x, c4.y, c4.z, c4.w};
  double X[4][2], u[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double2 x = reinterpret_cast<const double2 *>(A.coords)[nd[a]];
    const double2 v = reinterpret_cast<const double2 *>(A.U)[nd[a]];
    X[a][0] = x.x; X[a][1] = x.y; u[a][0] = v.x; u[a][1] = v.y;
  }
  double fe[8], Ke[8][8];
  q1_2d_element<DO_R, DO_K>(X, u, T, mp.p[0], mp.p[1], fe, Ke);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (nd[a] >= A.n_owned) continue;
    if (DO_R) { red_add(A.R + 2 * (int64_t)nd[a], fe[a * 2]); red_add(A.R + 2 * (int64_t)nd[a] + 1, fe[a * 2 + 1]); }
    if (DO_K) {
      const int64_t b0 = A.browptr[nd[a]];
      const int64_t deg = A.browptr[nd[a] + 1] - b0;
      const uint2 pp = reinterpret_cast<const uint2 *>(A.pos)[e * 4 + a];  // 4 x uint16
      const int pb[4] = {(int)(pp.x & 0xffff), (int)(pp.x >> 16), (int)(pp.y & 0xffff), (int)(pp.y >> 16)};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double *row = A.vals + 4 * b0 + i * 2 * deg;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          red_add(row + 2 * pb[b], Ke[a * 2 + i][b * 2]);
          red_add(row + 2 * pb[b] + 1, Ke[a * 2 + i][b * 2 + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// 3-D Q1 hexahedron, compressible Neo-Hookean, atomic scatter. Element integrals by eight lanes per
// element (hex_element.cuh: circulant symmetric half, 40 of 64 node blocks). Scatter = FP64 RED:
//   each lane parks its 5 blocks AND their mirror images in a 24 x 24 shared-memory tile; the 8 lanes
//   then walk the 24 element rows together, lane j adding columns j, j+8, j+16, so adjacent lanes hit
//   adjacent CSR entries (runs of >= 3): the L2 atomic rate roughly doubles against one isolated double
//   per lane (profiles/r1_ubench: 195 -> 355 G RED/s). Rows of nodes owned by another rank are skipped.
// ------------------------------------------------------------------------------------
#define H8_EPB 8                        // elements per CTA (64 threads): small CTAs keep ~5 resident per SM
#define H8_TS 25                        // row stride of the K_e tile (24 entries + 1: conflict-free)
#define H8_ESTRIDE (24 * H8_TS + 32)    // per element: max(HEX_REC = 544, K_e tile 600 + scatter meta 28) doubles

// one-time: lap[(a*5+d)][e] = sum_q w detJ (G_a . G_b), b = (a+d)&7 -- the lane/slot layout the assembly kernel reads back
__global__ void __launch_bounds__(8 * H8_EPB)
k_lap_build_3d(const AsmArgs A, const __grid_constant__ TabQ1 T, double *lap, int64_t stride) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, le = tid >> 3, a = tid & 7;
  const int64_t e = blockIdx.x * (int64_t)H8_EPB + le;
  if (e >= A.nel) return;
  const unsigned gmask = 0xffu << (8 * ((tid & 31) >> 3));
  double fe[3], Kr[5][3][3];
  hex_element_circulant<false, true, false, true>(smem + le * HEX_REC, a, gmask, A.conn[e * 8 + a], A.coords, A.coords, T, 1.0, 0.0, Kr, fe);
#pragma unroll
  for (int d = 0; d < 5; ++d) lap[(int64_t)(a * 5 + d) * stride + e] = Kr[d][0][0];
}

template <bool DO_R, bool DO_K, bool USE_LAP>
__global__ void __launch_bounds__(8 * H8_EPB, 5)
k_asm_q1_3d(const AsmArgs A, const __grid_constant__ TabQ1 T, const __grid_constant__ MatParams mp) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int le = tid >> 3;        // element within block
  const int a = tid & 7;          // lane within element
  const int64_t e = blockIdx.x * (int64_t)H8_EPB + le;
  if (e >= A.nel) return;         // whole 8-lane groups leave together
  constexpr int ES = DO_K ? H8_ESTRIDE : HEX_REC;   // the residual-only pass needs no K_e tile
  double *S = smem + le * ES;
  const unsigned gmask = 0xffu << (8 * ((tid & 31) >> 3));
  const int node = A.conn[e * 8 + a];
  double fe[3], Kr[5][3][3], lap5[5];
  if (DO_K && USE_LAP) {
#pragma unroll
    for (int d = 0; d < 5; ++d) lap5[d] = A.lap[(int64_t)(a * 5 + d) * A.lap_stride + e];
  }
  hex_element_circulant<DO_R, DO_K, USE_LAP>(S, a, gmask, node, A.coords, A.U, T, mp.p[0], mp.p[1], Kr, fe, lap5);

  const bool owned = node < A.n_owned;
  if (DO_R && owned) {
    double *r = A.R + 3 * (int64_t)node;
    red_add(r, fe[0]); red_add(r + 1, fe[1]); red_add(r + 2, fe[2]);
  }
  if (DO_K) {
    __syncwarp(gmask);                       // every lane of the element is done reading the qp records
    double *Tt = S;                          // [24 rows (a*3+i)][H8_TS]
    int64_t *b0m = reinterpret_cast<int64_t *>(S + 24 * H8_TS);          // [8] block-row start, -1 = not owned
    int *degm = reinterpret_cast<int *>(S + 24 * H8_TS + 8);             // [8]
    uint4 *posm = reinterpret_cast<uint4 *>(S + 24 * H8_TS + 12);        // [8] x 8 uint16
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const int b = (a + d) & 7;
      if (d == 4 && a >= 4) break;           // the antipodal pair was computed by both lanes: the lower one writes it
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          Tt[(a * 3 + i) * H8_TS + b * 3 + k] = Kr[d][i][k];
          if (d > 0) Tt[(b * 3 + k) * H8_TS + a * 3 + i] = Kr[d][i][k];   // K_{bk,ai} = K_{ai,bk}
        }
    }
    int64_t b0 = -1;
    int deg = 0;
    if (owned) { b0 = A.browptr[node]; deg = (int)(A.browptr[node + 1] - b0); }
    b0m[a] = b0; degm[a] = deg;
    posm[a] = reinterpret_cast<const uint4 *>(A.pos)[e * 8 + a];
    __syncwarp(gmask);
    const unsigned short *pos16 = reinterpret_cast<const unsigned short *>(posm);
    int cb[3], ck[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { const int c = a + 8 * m; cb[m] = c / 3; ck[m] = c - 3 * (c / 3); }
#pragma unroll 1
    for (int ar = 0; ar < 8; ++ar) {
      const int64_t rb0 = b0m[ar];
      if (rb0 < 0) continue;
      const int rdeg = degm[ar];
      int off[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) off[m] = 3 * (int)pos16[ar * 8 + cb[m]] + ck[m];
      double *rowp = A.vals + 9 * rb0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double *tr = Tt + (ar * 3 + i) * H8_TS + a;
#pragma unroll
        for (int m = 0; m < 3; ++m) red_add(rowp + i * 3 * rdeg + off[m], tr[8 * m]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Boundary kernels and one-time index kernels
// ------------------------------------------------------------------------------------
// Dirichlet apply!: D[bc.nodes] .= bc.values (boundary.jl:29). The list is deduplicated on
// the host (last write wins, as sequential application would).
__global__ void k_apply_dirichlet(int64_t n, const int64_t *dofs, const double *vals, double *U) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n) U[dofs[k]] = vals[k];
}
// R .-= Fext with Fext[bc.nodes] .+= bc.values (residual.jl:24-25, boundary.jl:39); duplicates allowed
__global__ void k_neumann(int64_t n, const int64_t *dofs, const double *vals, double *R, int64_t nrows) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n && dofs[k] < nrows) atomicAdd(R + dofs[k], -vals[k]);
}

// DistributedForce (src/gallery/distributedforce.jl:7-27): F_ext,a = sum_q (J f N_a(xi_q)) w_q per element,
// and R -= F_ext (residual.jl:19-25). 1-D: J = length(e)/2 as in the reference; Q1: J = det(dX/dxi).
template <int M>
__global__ void k_body_force_1d(int64_t nel, int64_t n_owned, const int32_t *conn, const double *coords, double f,
                                const __grid_constant__ Tab1D T, double *R) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nel) return;
  int nd[M];
#pragma unroll
  for (int a = 0; a < M; ++a) nd[a] = conn[e * M + a];
  const double J = (coords[nd[M - 1]] - coords[nd[0]]) / 2.0;
  const double Jf = J * f;
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double s = 0.0;
    for (int q = 0; q < T.nq; ++q) { const double v = (Jf * T.N[q][a]) * T.w[q]; s = (q == 0) ? v : s + v; }
    if (nd[a] < n_owned) atomicAdd(R + nd[a], -s);
  }
}
template <int DIM>
__global__ void k_body_force_q1(int64_t nel, int64_t n_owned, const int32_t *conn, const double *coords, double f0, double f1,
                                double f2, const __grid_constant__ TabQ1 T, double *R) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nel) return;
  constexpr int NEN = 1 << DIM;
  int nd[NEN];
  double X[NEN][DIM], s[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    nd[a] = conn[e * NEN + a];
    s[a] = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[a][d] = coords[(int64_t)nd[a] * DIM + d];
  }
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
    const double wd = T.w[q] * (DIM == 2 ? inv2(J, Ji) : inv3(J, Ji));
#pragma unroll
    for (int a = 0; a < NEN; ++a) s[a] += wd * T.N[q][a];
  }
  const double f[3] = {f0, f1, f2};
#pragma unroll
  for (int a = 0; a < NEN; ++a)
    if (nd[a] < n_owned)
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (f[d] != 0.0) atomicAdd(R + (int64_t)nd[a] * DIM + d, -s[a] * f[d]);
}

// position of node b in the (ascending) block row of node a, for every element node pair
__global__ void k_build_pos(int64_t nel, int nen, int64_t n_owned, const int32_t *conn,
                            const int64_t *browptr, const int32_t *bcol, uint16_t *pos) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nel * nen) return;
  const int64_t e = t / nen;
  const int na = conn[t];
  uint16_t *out = pos + t * nen;
  if (na >= n_owned) { for (int b = 0; b < nen; ++b) out[b] = 0; return; }
  const int64_t lo0 = browptr[na], hi0 = browptr[na + 1];
  for (int b = 0; b < nen; ++b) {
    const int nb = conn[e * nen + b];
    int64_t lo = lo0, hi = hi0 - 1;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (bcol[mid] < nb) lo = mid + 1; else hi = mid; }
    out[b] = (uint16_t)(lo - lo0);
  }
}

// scalar CSR = dim x dim expansion of the node-level pattern; also the diagonal offsets
__global__ void k_expand_pattern(int64_t n_owned, int dim, const int64_t *browptr, const int32_t *bcol,
                                 int64_t *rowptr, int32_t *colind, int32_t *diag) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nrows = n_owned * dim;
  if (r > nrows) return;
  if (r == nrows) { rowptr[r] = (int64_t)dim * dim * browptr[n_owned]; return; }
  const int64_t a = r / dim;
  const int i = (int)(r % dim);
  const int64_t b0 = browptr[a], deg = browptr[a + 1] - b0;
  const int64_t start = (int64_t)dim * dim * b0 + (int64_t)i * dim * deg;
  rowptr[r] = start;
  for (int64_t j = 0; j < deg; ++j) {
    const int32_t nb = bcol[b0 + j];
    for (int k = 0; k < dim; ++k) colind[start + j * dim + k] = nb * dim + k;
    if (nb == a) diag[r] = (int32_t)(j * dim + i);
  }
}

// save_state! (linearelastoplasticity.jl:96-103): u_n <- restrict(U); *_n <- *
__global__ void k_commit_state(int64_t nel, int nen, int nq, const int32_t *conn, const double *U,
                               double *u_n, double *sig, double *alp, double *kap, double *gam,
                               double *sig_n, double *alp_n, double *kap_n, double *gam_n) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nel) return;
  for (int a = 0; a < nen; ++a) u_n[e * nen + a] = U[conn[e * nen + a]];
  for (int q = 0; q < nq; ++q) {
    const int64_t k = e * nq + q;
    sig_n[k] = sig[k]; alp_n[k] = alp[k]; kap_n[k] = kap[k]; gam_n[k] = gam[k];
  }
}

// ------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------
static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

#define NLS_KCHECK(h) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string("kernel launch: ") + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

template <int M>
static int launch_1d(nls_context *h, const AsmArgs &a, bool r, bool k) {
  const unsigned g = nblk(h->nel, 128);
  if (r && k) k_asm_1d<M, true, true><<<g, 128, 0, h->stream>>>(a, h->tab1, h->mat);
  else if (r) k_asm_1d<M, true, false><<<g, 128, 0, h->stream>>>(a, h->tab1, h->mat);
  else        k_asm_1d<M, false, true><<<g, 128, 0, h->stream>>>(a, h->tab1, h->mat);
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_assembly(nls_context *h, bool want_r, bool want_k) {
  if (!want_r && !want_k) return NLS_OK;
  if (h->nel == 0) return NLS_OK;
  AsmArgs a{};
  a.nel = h->nel; a.n_owned = h->n_owned; a.conn = h->d_conn; a.coords = h->d_coords; a.U = h->d_U;
  a.R = h->d_R; a.vals = h->d_vals; a.browptr = h->d_browptr; a.pos = h->d_pos;
  for (int i = 0; i < 9; ++i) a.st[i] = h->d_state[i];
  a.u_n = h->d_u_n;
  if (h->dim == 1) {
    switch (h->nen) {
      case 2: return launch_1d<2>(h, a, want_r, want_k);
      case 3: return launch_1d<3>(h, a, want_r, want_k);
      case 4: return launch_1d<4>(h, a, want_r, want_k);
      case 5: return launch_1d<5>(h, a, want_r, want_k);
      case 6: return launch_1d<6>(h, a, want_r, want_k);
      case 7: return launch_1d<7>(h, a, want_r, want_k);
      case 8: return launch_1d<8>(h, a, want_r, want_k);
      default: NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "1-D elements support p = 1..7");
    }
  }
  if (h->mat.kind == NLS_MAT_J2_SMALL_STRAIN) return nls_launch_j2(h, a, want_r, want_k);
  const int mode = nls_assembly_mode(h);
  if (mode == NLS_ASM_GATHER) return nls_launch_gather_2d(h, want_r, want_k);
  if (mode == NLS_ASM_TWOPHASE) return nls_launch_twophase(h, want_r, want_k);
  if (mode == NLS_ASM_LINES) return nls_launch_lines3(h, want_r, want_k);
  if (h->dim == 2) {
    const unsigned g = nblk(h->nel, 128);
    if (want_r && want_k) k_asm_q1_2d<true, true><<<g, 128, 0, h->stream>>>(a, h->tabq, h->mat);
    else if (want_r)      k_asm_q1_2d<true, false><<<g, 128, 0, h->stream>>>(a, h->tabq, h->mat);
    else                  k_asm_q1_2d<false, true><<<g, 128, 0, h->stream>>>(a, h->tabq, h->mat);
    NLS_KCHECK(h);
    return NLS_OK;
  }
  const unsigned g = nblk(h->nel, H8_EPB);
  const size_t shm = sizeof(double) * H8_EPB * (want_k ? H8_ESTRIDE : HEX_REC);
  bool &attr_set = h->attr_asm3d;   // the attribute is per device: one handle = one device
  if (!attr_set) {
    const int shm_max = h->max_optin_smem;
    cudaFuncSetAttribute(k_asm_q1_3d<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm_max);
    cudaFuncSetAttribute(k_asm_q1_3d<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm_max);
    cudaFuncSetAttribute(k_asm_q1_3d<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm_max);
    cudaFuncSetAttribute(k_asm_q1_3d<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm_max);
    cudaFuncSetAttribute(k_asm_q1_3d<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm_max);
    cudaFuncSetAttribute(k_lap_build_3d, cudaFuncAttributeMaxDynamicSharedMemorySize, shm_max);
    attr_set = true;
  }
  // the Laplacian part of the tangent: once per mesh (coordinates and connectivity only), if its 320 B/element fit
  if (want_k && h->opt_lap3d && !h->lap_valid && !h->d_lap) {
    h->lap_stride = (h->nel + 31) & ~(int64_t)31;
    if (cudaMalloc((void **)&h->d_lap, sizeof(double) * 40 * (size_t)h->lap_stride) != cudaSuccess) { cudaGetLastError(); h->d_lap = nullptr; h->opt_lap3d = 0; }
  }
  if (want_k && h->opt_lap3d && h->d_lap && !h->lap_valid) {
    k_lap_build_3d<<<g, 8 * H8_EPB, sizeof(double) * H8_EPB * HEX_REC, h->stream>>>(a, h->tabq, h->d_lap, h->lap_stride);
    NLS_KCHECK(h);
    h->lap_valid = true;
  }
  const bool lap = want_k && h->opt_lap3d && h->lap_valid;
  a.lap = lap ? h->d_lap : nullptr; a.lap_stride = h->lap_stride;
  if (want_r && want_k) {
    if (lap) k_asm_q1_3d<true, true, true><<<g, 8 * H8_EPB, shm, h->stream>>>(a, h->tabq, h->mat);
    else     k_asm_q1_3d<true, true, false><<<g, 8 * H8_EPB, shm, h->stream>>>(a, h->tabq, h->mat);
  } else if (want_r) {
    k_asm_q1_3d<true, false, false><<<g, 8 * H8_EPB, shm, h->stream>>>(a, h->tabq, h->mat);
  } else {
    if (lap) k_asm_q1_3d<false, true, true><<<g, 8 * H8_EPB, shm, h->stream>>>(a, h->tabq, h->mat);
    else     k_asm_q1_3d<false, true, false><<<g, 8 * H8_EPB, shm, h->stream>>>(a, h->tabq, h->mat);
  }
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_apply_dirichlet(nls_context *h, double *U) {
  if (h->ndir == 0) return NLS_OK;
  k_apply_dirichlet<<<nblk(h->ndir, 256), 256, 0, h->stream>>>(h->ndir, h->d_dir_dofs, h->d_dir_vals, U);
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_neumann(nls_context *h, double *R) {
  if (h->nneu == 0) return NLS_OK;
  k_neumann<<<nblk(h->nneu, 256), 256, 0, h->stream>>>(h->nneu, h->d_neu_dofs, h->d_neu_vals, R, h->nrows);
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_body_force(nls_context *h, double *R) {
  if (!h->have_body || h->nel == 0) return NLS_OK;
  const unsigned g = nblk(h->nel, 128);
  if (h->dim == 1) {
    switch (h->nen) {
#define BF(M) case M: k_body_force_1d<M><<<g, 128, 0, h->stream>>>(h->nel, h->n_owned, h->d_conn, h->d_coords, h->body[0], h->tab1, R); break;
      BF(2) BF(3) BF(4) BF(5) BF(6) BF(7) BF(8)
#undef BF
      default: NLS_FAIL(h, NLS_ERR_UNSUPPORTED, "1-D elements support p = 1..7");
    }
  } else if (h->dim == 2) {
    k_body_force_q1<2><<<g, 128, 0, h->stream>>>(h->nel, h->n_owned, h->d_conn, h->d_coords, h->body[0], h->body[1], 0.0, h->tabq, R);
  } else {
    k_body_force_q1<3><<<g, 128, 0, h->stream>>>(h->nel, h->n_owned, h->d_conn, h->d_coords, h->body[0], h->body[1], h->body[2], h->tabq, R);
  }
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_build_pos(nls_context *h) {
  if (h->nel == 0) return NLS_OK;
  k_build_pos<<<nblk(h->nel * h->nen, 256), 256, 0, h->stream>>>(h->nel, h->nen, h->n_owned, h->d_conn,
                                                                 h->d_browptr, h->d_bcol, h->d_pos);
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_expand_pattern(nls_context *h) {
  k_expand_pattern<<<nblk(h->nrows + 1, 256), 256, 0, h->stream>>>(h->n_owned, h->dim, h->d_browptr, h->d_bcol,
                                                                   h->d_rowptr, h->d_colind, h->d_diag);
  NLS_KCHECK(h);
  return NLS_OK;
}

int nls_launch_commit_state(nls_context *h, const double *dU) {
  if (h->mat.kind != NLS_MAT_ELASTOPLASTIC_1D || h->nel == 0) return NLS_OK;
  k_commit_state<<<nblk(h->nel, 256), 256, 0, h->stream>>>(h->nel, h->nen, h->nq, h->d_conn, dU, h->d_u_n,
      h->d_state[0], h->d_state[1], h->d_state[2], h->d_state[3],
      h->d_state[5], h->d_state[6], h->d_state[7], h->d_state[8]);
  NLS_KCHECK(h);
  return NLS_OK;
}
