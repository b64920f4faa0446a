// kernels_j2.cu -- small-strain J2 elastoplasticity with linear isotropic hardening on Q1 quads (plane strain) and hexahedra:
// the stateful-material protocol of the reference generalised from the 1-D bar to 2-D / 3-D.
//
// Pattern followed: src/materials/linearelastoplasticity.jl:60-103 -- the internal-force pass runs the return mapping at every
// quadrature point from the COMMITTED state and the current displacement (update_state_el!), stores the trial state and the
// consistent-tangent data; the tangent pass reads what the last internal-force pass stored (state.dsigma in the reference);
// save_state! (poststep!, src/timesteppers/types.jl:74-78) commits trial -> committed only after the Newton solve of the load
// step has converged. Radial return (Simo & Hughes, box 3.2):
//   eps = sym grad u,  s_tr = 2G dev(eps - eps_p_n),  f = |s_tr| - sqrt(2/3) (sigma_y + H alpha_n)
//   f <= 0: sigma = K tr(eps - eps_p_n) I + s_tr                                  (theta = 1, thetabar = 0)
//   f > 0 : dgamma = f / (2G + 2H/3), n = s_tr / |s_tr|, sigma = sigma_tr - 2G dgamma n, eps_p = eps_p_n + dgamma n,
//           alpha = alpha_n + sqrt(2/3) dgamma, theta = 1 - 2G dgamma / |s_tr|, thetabar = 1 / (1 + H / 3G) - (1 - theta)
//   C = K I(x)I + 2G theta (I_sym - I(x)I/3) - 2G thetabar n(x)n
// One thread per element, FP64 RED scatter (the atomic-free schemes are specific to the Neo-Hookean kernels; this material is
// here for the protocol, not for the headline).
#include "asm_common.cuh"

#ifndef NLS_TRY
#define NLS_TRY(call) do { int rc_ = (call); if (rc_ != NLS_OK) return rc_; } while (0)
#endif

struct J2Args {
  double *comm, *trial, *tang;     // [7][nqp], [7][nqp], [8][nqp]: eps_p (xx, yy, zz, xy, yz, xz), alpha | n (6), theta, thetabar
  int64_t nqp;
};

template <int DIM, bool DO_R, bool DO_K>
__global__ void __launch_bounds__(64)
k_asm_j2(const AsmArgs A, const __grid_constant__ TabQ1 T, const __grid_constant__ MatParams mp, const J2Args S) {
  constexpr int NEN = 1 << DIM, NQ = NEN;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.nel) return;
  const double Kb = mp.p[0], G = mp.p[1], sy = mp.p[2], Hh = mp.p[3];
  int nd[NEN];
  double X[NEN][DIM], u[NEN][DIM];
  for (int a = 0; a < NEN; ++a) {
    nd[a] = A.conn[e * NEN + a];
    for (int d = 0; d < DIM; ++d) { X[a][d] = A.coords[(int64_t)nd[a] * DIM + d]; u[a][d] = A.U[(int64_t)nd[a] * DIM + d]; }
  }
  double Gq[NQ][NEN][DIM], wd[NQ], sig[NQ][6], nn[NQ][6], th[NQ], thb[NQ];
  for (int q = 0; q < NQ; ++q) {
    double J[DIM * DIM], Ji[DIM * DIM];
    for (int k = 0; k < DIM * DIM; ++k) J[k] = 0.0;
    for (int a = 0; a < NEN; ++a)
      for (int d = 0; d < DIM; ++d)
        for (int D = 0; D < DIM; ++D) J[d * DIM + D] += X[a][d] * T.dN[q][a][D];
    const double detJ = DIM == 2 ? inv2(J, Ji) : inv3(J, Ji);
    wd[q] = T.w[q] * detJ;
    for (int a = 0; a < NEN; ++a)
      for (int D = 0; D < DIM; ++D) {
        double s = 0.0;
        for (int E = 0; E < DIM; ++E) s += T.dN[q][a][E] * Ji[E * DIM + D];
        Gq[q][a][D] = s;
      }
    const int64_t iq = e * NQ + q;
    if (DO_R) {
      double gu[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int a = 0; a < NEN; ++a)
        for (int i = 0; i < DIM; ++i)
          for (int j = 0; j < DIM; ++j) gu[i][j] += u[a][i] * Gq[q][a][j];
      // strain (xx, yy, zz, xy, yz, xz), elastic trial strain
      const double eps[6] = {gu[0][0], gu[1][1], gu[2][2], 0.5 * (gu[0][1] + gu[1][0]), 0.5 * (gu[1][2] + gu[2][1]), 0.5 * (gu[0][2] + gu[2][0])};
      double epn[6], ee[6];
      for (int k = 0; k < 6; ++k) { epn[k] = S.comm[k * S.nqp + iq]; ee[k] = eps[k] - epn[k]; }
      const double alpha_n = S.comm[6 * S.nqp + iq];
      const double tr = ee[0] + ee[1] + ee[2];
      double s[6] = {2.0 * G * (ee[0] - tr / 3.0), 2.0 * G * (ee[1] - tr / 3.0), 2.0 * G * (ee[2] - tr / 3.0), 2.0 * G * ee[3], 2.0 * G * ee[4], 2.0 * G * ee[5]};
      const double qn = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]));
      const double f = qn - sqrt(2.0 / 3.0) * (sy + Hh * alpha_n);
      double dg = 0.0, theta = 1.0, thetab = 0.0, nv[6] = {0, 0, 0, 0, 0, 0};
      if (f > 0.0) {
        dg = f / (2.0 * G + 2.0 * Hh / 3.0);
        for (int k = 0; k < 6; ++k) nv[k] = s[k] / qn;
        theta = 1.0 - 2.0 * G * dg / qn;
        thetab = 1.0 / (1.0 + Hh / (3.0 * G)) - (1.0 - theta);
      }
      for (int k = 0; k < 6; ++k) {
        sig[q][k] = s[k] - 2.0 * G * dg * nv[k] + (k < 3 ? Kb * tr : 0.0);
        nn[q][k] = nv[k];
        S.trial[k * S.nqp + iq] = epn[k] + dg * nv[k];
        S.tang[k * S.nqp + iq] = nv[k];
      }
      S.trial[6 * S.nqp + iq] = alpha_n + sqrt(2.0 / 3.0) * dg;
      S.tang[6 * S.nqp + iq] = theta; S.tang[7 * S.nqp + iq] = thetab;
      th[q] = theta; thb[q] = thetab;
    } else {
      for (int k = 0; k < 6; ++k) nn[q][k] = S.tang[k * S.nqp + iq];
      th[q] = S.tang[6 * S.nqp + iq]; thb[q] = S.tang[7 * S.nqp + iq];
    }
  }
  // symmetric 3 x 3 from the 6-vector
  auto sym = [](const double *v, int i, int j) -> double {
    if (i == j) return v[i];
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    return lo == 0 ? (hi == 1 ? v[3] : v[5]) : v[4];
  };
  for (int a = 0; a < NEN; ++a) {
    if (nd[a] >= A.n_owned) continue;
    if (DO_R) {
      for (int i = 0; i < DIM; ++i) {
        double s = 0.0;
        for (int q = 0; q < NQ; ++q)
          for (int j = 0; j < DIM; ++j) s += wd[q] * sym(sig[q], i, j) * Gq[q][a][j];
        red_add(A.R + (int64_t)DIM * nd[a] + i, s);
      }
    }
    if (DO_K) {
      const int64_t b0 = A.browptr[nd[a]];
      const int64_t deg = A.browptr[nd[a] + 1] - b0;
      for (int b = 0; b < NEN; ++b) {
        double Bk[DIM][DIM];
        for (int i = 0; i < DIM; ++i) for (int k = 0; k < DIM; ++k) Bk[i][k] = 0.0;
        for (int q = 0; q < NQ; ++q) {
          const double c1 = wd[q] * (Kb - 2.0 / 3.0 * G * th[q]), c2 = wd[q] * G * th[q], c3 = wd[q] * 2.0 * G * thb[q];
          double gab = 0.0, na[DIM], nb[DIM];
          for (int j = 0; j < DIM; ++j) gab += Gq[q][a][j] * Gq[q][b][j];
          for (int i = 0; i < DIM; ++i) {
            na[i] = 0.0; nb[i] = 0.0;
            for (int j = 0; j < DIM; ++j) { na[i] += sym(nn[q], i, j) * Gq[q][a][j]; nb[i] += sym(nn[q], i, j) * Gq[q][b][j]; }
          }
          for (int i = 0; i < DIM; ++i)
            for (int k = 0; k < DIM; ++k)
              Bk[i][k] += c1 * Gq[q][a][i] * Gq[q][b][k] + c2 * ((i == k ? gab : 0.0) + Gq[q][a][k] * Gq[q][b][i]) - c3 * na[i] * nb[k];
        }
        const int pb = A.pos[(e * NEN + a) * NEN + b];
        for (int i = 0; i < DIM; ++i) {
          double *row = A.vals + (int64_t)DIM * DIM * b0 + (int64_t)i * DIM * deg + (int64_t)DIM * pb;
          for (int k = 0; k < DIM; ++k) red_add(row + k, Bk[i][k]);
        }
      }
    }
  }
}

__global__ void k_j2_copy(int64_t n, const double *__restrict__ src, double *__restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// ---- host side ----
void nls_j2_free(nls_context *h) { if (h->d_j2) cudaFree(h->d_j2); h->d_j2 = nullptr; h->j2_nqp = 0; }

// initialize_state!: eps_p = 0, alpha = 0 (committed and trial), elastic tangent data
int nls_j2_init(nls_context *h) {
  const int64_t nqp = h->nel * (int64_t)h->nen;
  if (h->j2_nqp != nqp) { nls_j2_free(h); }
  if (nqp == 0) return NLS_OK;
  if (!h->d_j2) { NLS_CUDA(h, cudaMalloc((void **)&h->d_j2, sizeof(double) * 22 * (size_t)nqp)); h->j2_nqp = nqp; }
  NLS_CUDA(h, cudaMemsetAsync(h->d_j2, 0, sizeof(double) * 22 * (size_t)nqp, h->stream));
  std::vector<double> ones((size_t)nqp, 1.0);                          // theta = 1 (elastic) until the first internal-force pass
  NLS_CUDA(h, cudaMemcpyAsync(h->d_j2 + (14 + 6) * nqp, ones.data(), sizeof(double) * (size_t)nqp, cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLS_OK;
}

// save_state!: committed <- trial
int nls_j2_commit(nls_context *h) {
  if (!h->d_j2) NLS_FAIL(h, NLS_ERR_STATE, "J2 state not initialised");
  const int64_t n = 7 * h->j2_nqp;
  k_j2_copy<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4096)), 256, 0, h->stream>>>(n, h->d_j2 + 7 * h->j2_nqp, h->d_j2);
  h->launches++;
  return NLS_OK;
}

// one per-qp array by name: "epsp_xx" ... "epsp_xz", "alpha" (+ "_n" = committed)
int nls_j2_get(nls_context *h, const char *name, double *out) {
  static const char *names[7] = {"epsp_xx", "epsp_yy", "epsp_zz", "epsp_xy", "epsp_yz", "epsp_xz", "alpha"};
  if (!h->d_j2) NLS_FAIL(h, NLS_ERR_STATE, "nls_get_state: material has no state");
  std::string s(name);
  bool committed = false;
  if (s.size() > 2 && s.substr(s.size() - 2) == "_n") { committed = true; s = s.substr(0, s.size() - 2); }
  for (int k = 0; k < 7; ++k)
    if (s == names[k]) {
      const double *src = h->d_j2 + ((committed ? 0 : 7) + k) * h->j2_nqp;
      NLS_CUDA(h, cudaMemcpyAsync(out, src, sizeof(double) * h->j2_nqp, cudaMemcpyDeviceToHost, h->stream));
      NLS_CUDA(h, cudaStreamSynchronize(h->stream));
      return NLS_OK;
    }
  NLS_FAIL(h, NLS_ERR_ARG, "nls_get_state: unknown state name");
}

int nls_launch_j2(nls_context *h, const AsmArgs &a, bool want_r, bool want_k) {
  if (!h->d_j2 || h->j2_nqp != h->nel * (int64_t)h->nen) NLS_TRY(nls_j2_init(h));
  J2Args S;
  S.nqp = h->j2_nqp; S.comm = h->d_j2; S.trial = h->d_j2 + 7 * S.nqp; S.tang = h->d_j2 + 14 * S.nqp;
  const unsigned g = (unsigned)((h->nel + 63) / 64);
#define J2_LAUNCH(D, R, K) k_asm_j2<D, R, K><<<g, 64, 0, h->stream>>>(a, h->tabq, h->mat, S)
  if (h->dim == 2) { if (want_r && want_k) J2_LAUNCH(2, true, true); else if (want_r) J2_LAUNCH(2, true, false); else J2_LAUNCH(2, false, true); }
  else { if (want_r && want_k) J2_LAUNCH(3, true, true); else if (want_r) J2_LAUNCH(3, true, false); else J2_LAUNCH(3, false, true); }
#undef J2_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("J2 kernel launch: ") + cudaGetErrorString(e); return NLS_ERR_CUDA; }
  h->launches++;
  return NLS_OK;
}
