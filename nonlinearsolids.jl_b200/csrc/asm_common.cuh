// asm_common.cuh -- argument block and small dense helpers shared by the assembly kernels.
#pragma once
#include "nls_internal.h"

struct AsmArgs {
  int64_t nel, n_owned;
  const int32_t *conn;
  const double *coords;
  const double *U;
  double *R;
  double *vals;
  const int64_t *browptr;
  const uint16_t *pos;
  double *st[9];   // sig, alp, kap, gam, dsig, sig_n, alp_n, kap_n, gam_n
  const double *u_n;
  const double *lap;       // 3-D: [40][lap_stride] precomputed sum_q w detJ (G_a . G_b) per lane slot (a*5+d), or NULL
  int64_t lap_stride;
};

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }


// ------------------------------------------------------------------------------------
// Small dense helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double inv2(const double *M, double *Mi) {
  const double det = M[0] * M[3] - M[1] * M[2];
  const double id = 1.0 / det;
  Mi[0] = M[3] * id; Mi[1] = -M[1] * id; Mi[2] = -M[2] * id; Mi[3] = M[0] * id;
  return det;
}
__device__ __forceinline__ double inv3(const double *M, double *Mi) {
  const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  const double id = 1.0 / det;
  Mi[0] = c00 * id; Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id; Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  Mi[3] = c01 * id; Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id; Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  Mi[6] = c02 * id; Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id; Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  return det;
}

