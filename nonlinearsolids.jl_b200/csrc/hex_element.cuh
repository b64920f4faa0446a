// hex_element.cuh -- element integrals of one Q1 hexahedron (compressible Neo-Hookean) by EIGHT cooperating
// lanes (4 elements per warp); K_e (24 x 24) does not fit one thread's registers. Shared by the atomic-scatter
// kernel (kernels_assembly.cu) and the element-record kernel (kernels_twophase.cu).
//   phase 0: lane a gathers node a (X_a, u_a) into shared memory
//   phase 1: lane q evaluates quadrature point q: J, dN/dX, F, F^-1, lnJ, P, g -> shared memory
//   phase 2: lane a accumulates the CIRCULANT half of row block a: blocks (a, (a+d)&7), d = 0..4 -- every
//            unordered node pair exactly once (the four antipodal pairs twice) -- and f_a. K_e is symmetric,
//            so this is 40 of 64 node blocks (-37 % DFMA against owning the whole row block); the price is
//            that the eight lanes read eight different neighbours from shared memory instead of broadcasting.
//   F = I + sum_a u_a (x) dN_a/dX,  P = mu (F - F^-T) + lam lnJ F^-T
//   K_{ai,bk} = w detJ [ mu d_ik (G_a.G_b) + lam g_ai g_bk + (mu - lam lnJ) g_ak g_bi ],  G_a = dN_a/dX, g_a = F^-T G_a
// Shared-memory layout per element (doubles): 8 qp records of 62 {gG[8][6] = {g, G} per node, P[9], c[3], pad 2}, then Xu[8][6];
// 62 * 8 B = 31 16-byte units per lane -> conflict-free STS.128 in phase 1.
#pragma once
#include "asm_common.cuh"

#define HEX_QSTRIDE 62
#define HEX_REC (8 * HEX_QSTRIDE + 48)    // doubles of shared memory one element needs for phases 0-2

// S: this element's shared-memory area (>= HEX_REC doubles); ln = lane within the 8-lane group; gmask = the
// group's lanes. All 8 lanes of an element must call this together. Returns Kr[d] = block (ln, (ln+d)&7), fe.
// USE_LAP: the deformation-independent part of the tangent, mu * sum_q w detJ (G_a . G_b) (the vector Laplacian of the
// reference configuration), is not re-accumulated but read from `lap` (this lane's 5 values, precomputed once per mesh by
// the LAP_BUILD variant and stored without mu). That removes the G_b half of the neighbour reads from shared memory -- the
// binding resource of phase 2 -- and 3 of 21 DFMA per block and point.
// LAP_BUILD: only that sum is computed (returned in Kr[d][0][0]); U, mu, lam are ignored.
template <bool DO_R, bool DO_K, bool USE_LAP = false, bool LAP_BUILD = false>
__device__ __forceinline__ void hex_element_circulant(double *__restrict__ S, const int ln, const unsigned gmask, const int node,
                                                      const double *__restrict__ coords, const double *__restrict__ U,
                                                      const TabQ1 &T, const double mu, const double lam,
                                                      double (&Kr)[5][3][3], double (&fe)[3], const double *lap5 = nullptr) {
  double *Xu = S + 8 * HEX_QSTRIDE;
  {
    const double *xc = coords + 3 * (int64_t)node;
    const double *uc = U + 3 * (int64_t)node;
    Xu[ln * 6 + 0] = xc[0]; Xu[ln * 6 + 1] = xc[1]; Xu[ln * 6 + 2] = xc[2];
    Xu[ln * 6 + 3] = LAP_BUILD ? 0.0 : uc[0]; Xu[ln * 6 + 4] = LAP_BUILD ? 0.0 : uc[1]; Xu[ln * 6 + 5] = LAP_BUILD ? 0.0 : uc[2];
  }
  __syncwarp(gmask);
  {
    const int q = ln;
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Ji[9], F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Fi[9];
    double G[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double dn0 = T.dN[q][a][0], dn1 = T.dN[q][a][1], dn2 = T.dN[q][a][2];
      G[a][0] = dn0; G[a][1] = dn1; G[a][2] = dn2;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double x = Xu[a * 6 + d];
        J[d * 3 + 0] = fma(x, dn0, J[d * 3 + 0]); J[d * 3 + 1] = fma(x, dn1, J[d * 3 + 1]); J[d * 3 + 2] = fma(x, dn2, J[d * 3 + 2]);
      }
    }
    const double detJ = inv3(J, Ji);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double d0 = G[a][0], d1 = G[a][1], d2 = G[a][2];
#pragma unroll
      for (int D = 0; D < 3; ++D) G[a][D] = fma(d0, Ji[D], fma(d1, Ji[3 + D], d2 * Ji[6 + D]));
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double ui = Xu[a * 6 + 3 + i];
        F[i * 3 + 0] = fma(ui, G[a][0], F[i * 3 + 0]); F[i * 3 + 1] = fma(ui, G[a][1], F[i * 3 + 1]); F[i * 3 + 2] = fma(ui, G[a][2], F[i * 3 + 2]);
      }
    }
    const double detF = inv3(F, Fi);
    const double lnJ = log(detF);
    const double wd = T.w[q] * detJ;
    double *Q = S + q * HEX_QSTRIDE;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      double2 *dst = reinterpret_cast<double2 *>(Q + a * 6);
      const double g0 = fma(G[a][0], Fi[0], fma(G[a][1], Fi[3], G[a][2] * Fi[6]));
      const double g1 = fma(G[a][0], Fi[1], fma(G[a][1], Fi[4], G[a][2] * Fi[7]));
      const double g2 = fma(G[a][0], Fi[2], fma(G[a][1], Fi[5], G[a][2] * Fi[8]));
      dst[0] = make_double2(g0, g1);               // node record: {g0, g1, g2, G0, G1, G2}
      dst[1] = make_double2(g2, G[a][0]);
      dst[2] = make_double2(G[a][1], G[a][2]);
    }
    const double c1 = LAP_BUILD ? wd : wd * mu, cl = wd * lam * lnJ;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int Jc = 0; Jc < 3; ++Jc)
        Q[48 + i * 3 + Jc] = fma(c1, F[i * 3 + Jc] - Fi[Jc * 3 + i], cl * Fi[Jc * 3 + i]);
    Q[57] = c1; Q[58] = wd * lam; Q[59] = c1 - cl;
  }
  __syncwarp(gmask);

  const int a = ln;
#pragma unroll
  for (int i = 0; i < 3; ++i) fe[i] = 0.0;
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) Kr[d][i][k] = 0.0;
#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    const double *Q = S + q * HEX_QSTRIDE;
    const double2 t0 = reinterpret_cast<const double2 *>(Q + a * 6)[0];
    const double2 t1 = reinterpret_cast<const double2 *>(Q + a * 6)[1];
    const double2 t2 = reinterpret_cast<const double2 *>(Q + a * 6)[2];
    const double ga[3] = {t0.x, t0.y, t1.x}, Ga[3] = {t1.y, t2.x, t2.y};
    if (DO_R) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
        fe[i] = fma(Q[48 + i * 3], Ga[0], fma(Q[48 + i * 3 + 1], Ga[1], fma(Q[48 + i * 3 + 2], Ga[2], fe[i])));
    }
    if (DO_K) {
      const double c1 = Q[57], c2 = Q[58], c3 = Q[59];
      const double c1G[3] = {c1 * Ga[0], c1 * Ga[1], c1 * Ga[2]};
      const double c2g[3] = {c2 * ga[0], c2 * ga[1], c2 * ga[2]};
      const double c3g[3] = {c3 * ga[0], c3 * ga[1], c3 * ga[2]};
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const int b = (a + d) & 7;
        double Gb[3] = {0, 0, 0}, gb[3];
        if (d == 0) {
          Gb[0] = Ga[0]; Gb[1] = Ga[1]; Gb[2] = Ga[2]; gb[0] = ga[0]; gb[1] = ga[1]; gb[2] = ga[2];
        } else if (USE_LAP) {                      // g_b only: one LDS.128 + one LDS.64
          const double2 s0 = reinterpret_cast<const double2 *>(Q + b * 6)[0];
          gb[0] = s0.x; gb[1] = s0.y; gb[2] = Q[b * 6 + 2];
        } else {
          const double2 s0 = reinterpret_cast<const double2 *>(Q + b * 6)[0];
          const double2 s1 = reinterpret_cast<const double2 *>(Q + b * 6)[1];
          const double2 s2 = reinterpret_cast<const double2 *>(Q + b * 6)[2];
          gb[0] = s0.x; gb[1] = s0.y; gb[2] = s1.x; Gb[0] = s1.y; Gb[1] = s2.x; Gb[2] = s2.y;
        }
        if (LAP_BUILD) {
          Kr[d][0][0] += fma(c1G[0], Gb[0], fma(c1G[1], Gb[1], c1G[2] * Gb[2]));
        } else if (USE_LAP) {
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) Kr[d][i][k] = fma(c2g[i], gb[k], fma(c3g[k], gb[i], Kr[d][i][k]));
        } else {
          const double dab = fma(c1G[0], Gb[0], fma(c1G[1], Gb[1], c1G[2] * Gb[2]));
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k)
              Kr[d][i][k] += fma(c2g[i], gb[k], fma(c3g[k], gb[i], (i == k ? dab : 0.0)));
        }
      }
    }
  }
  if (DO_K && USE_LAP) {
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const double l = mu * lap5[d];
#pragma unroll
      for (int i = 0; i < 3; ++i) Kr[d][i][i] += l;
    }
  }
}
