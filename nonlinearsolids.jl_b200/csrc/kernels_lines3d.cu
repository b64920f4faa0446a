// kernels_lines3d.cu -- atomic-free, row-owner assembly of the 3-D Q1 Neo-Hookean residual and CSR Jacobian on meshes
// with grid topology (lexicographic hexahedral grids and their slab partitions, coordinates arbitrary).
//
// Replaces, like kernels_assembly.cu:  compute_internalforce / compute_dinternalforce (src/materials/defaults.jl:29-46,
// 68-82) with restrict! / unrestrict! / assemble! (src/fem/element.jl:64-70) folded in. The 24 x 24 element matrix is
// never formed: every 3 x 3 CSR block K_AB = sum over the elements e that hold both nodes, over their 8 quadrature
// points, of
//     c2 g_A (x) g_B + c3 g_B (x) g_A   (+ mu * L_AB on the diagonal, L = reference-configuration Laplacian, built once per mesh)
// is summed by one thread in a fixed order, together with its transpose K_BA (K is symmetric): no atomics, no zero-fill
// of the CSR values, no element-matrix exchange, bit-reproducible run to run.
//
// Work decomposition. Node lines run along x (the fastest node index), so the CSR rows of one line are ONE contiguous
// stream of doubles. A CTA owns the rows of a tile of L3_TY x L3_TZ lines and marches along x; per x-interval i (the
// elements between node planes i and i+1 of the (TY+1) x (TZ+1) element columns that touch the tile), in two halves of
// four quadrature points each (the records of eight do not fit beside the row staging):
//   phase A  one thread per (element column, quadrature point): J, H = sum u (x) dN, M = J + H, M^-1, ln det F,
//            g_a = M^-T dN_a (spatial gradients) and the scalars -> shared-memory records (SoA, 16-byte columns:
//            every LDS.128 / STS.128 of a warp is conflict-free)
//   phase B  one thread per line pair (L, L + d), d in {(0,0), (1,0), (-1,1), (0,1), (1,1)} in (y, z), with at least one
//            line in the tile (pairs across the tile border are evaluated by both tiles: 366 pair threads per 64 lines).
//            It accumulates the four blocks between nodes {i, i+1} of L and {i, i+1} of L + d over the one or two
//            (self: four) element columns that hold both lines -- symmetric and skew parts separately, 15 DFMA per block
//            and point instead of 18 -- and carries the (i+1, i+1) block and the two x-off-diagonal blocks into the next
//            interval, where they complete the rows of node i+1.
//   staging  at the end of the interval every pair thread drops its blocks of node plane i -- direct into the row of L,
//            transposed into the row of L + d -- into a shared-memory image of the tile's 64 complete CSR rows, and one
//            thread per line hands the row to the TMA unit: cp.async.bulk shared -> global of the 16-byte aligned part
//            (~1.9 KB per line and interval; the odd leading / trailing double is carried to the next interval's copy).
//            The copy drains while the next interval is computed; the SM issues no global store for the Jacobian.
// Every CSR row is written once, in full 16-byte units, by exactly one CTA. Row addresses are closed-form per line (host
// tables, verified against the CSR pattern when the plan is built); meshes without grid topology keep the atomic scatter.
#include <algorithm>
#include <cstring>
#include "asm_common.cuh"

#define L3_TY 8
#define L3_TZ 8
#define L3_NL (L3_TY * L3_TZ)            // lines per tile
#define L3_CY (L3_TY + 1)                // element columns per tile row
#define L3_NCOL (L3_CY * (L3_TZ + 1))    // 81 element columns touch a tile
#define L3_ZCOL L3_NCOL                  // an all-zero column: what pairs read where the mesh has no element
#define L3_NCOLP (L3_NCOL + 1)
#define L3_NY2 (L3_TY + 2)               // node lines cached per row
#define L3_NLC (L3_NY2 * (L3_TZ + 2))    // 100 node lines touch those columns
#define L3_NLCP 110
// first cell of row lz of the cached node lines in XU: 9 lz mod 16 is the residue, so that the lanes of phase A (consecutive
// element columns, 9 per row) read consecutive cells modulo 16 -- 8 different 16-byte bank groups per quarter-warp for the
// double2 loads; rows laid out without overlap in 110 cells (a plain 10-cell row stride made every XU load a two-way bank
// conflict: 96 M of the 139 M excess wavefronts at C3)
__constant__ unsigned char c_l3_xrow[L3_TZ + 2] = {0, 89, 66, 11, 100, 45, 22, 79, 56, 33};
#define L3_THREADS 384                   // 12 warps: 2 of self pairs, 5 of two-column pairs, 5 of one-column pairs
#define L3_NPT L3_THREADS                // Laplacian-table slots per interval: one per thread
#define L3_QH 4                          // quadrature points per half interval
#ifndef L3_PAIR_UNROLL
#define L3_PAIR_UNROLL 2                 // unroll factors of the two phase-B loops (measured)
#endif
#ifndef L3_SELF_UNROLL
#define L3_SELF_UNROLL 2
#endif
static constexpr int kL3PairUnroll = L3_PAIR_UNROLL, kL3SelfUnroll = L3_SELF_UNROLL;
#define L3_SLOT 246                      // staging doubles per line: 243 (27 blocks x 9) + alignment slack, 16-byte multiple

struct L3Line {                          // per node line (y, z): closed-form CSR addressing of its rows
  int64_t lb;                            // block-row start of the line's first node
  int32_t nl;                            // neighbour lines present (block-row length = nl * (2 or 3))
  int8_t rank[9];                        // rank of neighbour line (dz+1)*3 + (dy+1) among them in column order, -1 = absent
  int8_t pad[3];
};

struct L3Args {
  int nx, ny, nzp, nunits;
  const int32_t *pbase;                  // [nzp] first node id of each node plane
  const uint8_t *pown;                   // [nzp] plane holds owned nodes
  const L3Line *info;                    // [nzp * ny]
  const int4 *units;                     // {y0, z0, i0, i1}: tile origin and x-interval range
  const int64_t *lapoff;                 // [nunits] offset of the unit's Laplacian values
  double *lap;                           // [unit][interval (+1)][3][L3_NPT]
  const double *coords, *U;
  double *R, *vals;
  unsigned *counter;                     // [2] next unit, CTAs done
  double *carry;                         // [grid][18][L3_THREADS] x-off-diagonal blocks on their way to the next node plane's rows (L2-resident)
  int dbg_nostore;                       // experiments only: 1 = skip the global stores
  unsigned long long *prof;              // optional cycle counters (nls_set_option "profile_phases"): phase A, phase B, intervals
};

__device__ __forceinline__ void l3_cp8(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void l3_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void l3_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// TMA bulk copy shared -> global (no tensor map: a flat run of 16-byte units), tracked by bulk async-groups
__device__ __forceinline__ void l3_bulk_store(double *gdst, const double *ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l3_bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void l3_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the TMA unit) after the next barrier
__device__ __forceinline__ void l3_fence_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// stores of 3 consecutive doubles at an 8-byte aligned address (the residual): 16-byte stores wherever the alignment
// allows. Inline PTX on purpose: written as C++ the optimiser merges the stores of the two alignment branches into one
// unconditional 16-byte store that is misaligned for half of the addresses (seen in the SASS of the first version).
__device__ __forceinline__ void l3_stg1(double *p, double a) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(a)); }
__device__ __forceinline__ void l3_stg2(double *p, double a, double b) { asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b)); }
__device__ __forceinline__ void l3_st3(double *p, double a, double b, double c) {
  if ((reinterpret_cast<uintptr_t>(p) & 8) == 0) { l3_stg2(p, a, b); l3_stg1(p + 2, c); }
  else { l3_stg1(p, a); l3_stg2(p + 1, b, c); }
}

// 6 consecutive doubles into the staging image at an 8-byte aligned shared address: 16-byte stores where the alignment allows
// (conflict-free across the lines of a warp: the line slots are an odd number of 16-byte units apart)
__device__ __forceinline__ void l3_sts6(double *p, double a, double b, double c, double d, double e, double f) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(p);
  if ((s & 8u) == 0) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n\tst.shared.v2.f64 [%0+16], {%3, %4};\n\tst.shared.v2.f64 [%0+32], {%5, %6};"
                 ::"r"(s), "d"(a), "d"(b), "d"(c), "d"(d), "d"(e), "d"(f) : "memory");
  } else {
    asm volatile("st.shared.f64 [%0], %1;\n\tst.shared.v2.f64 [%0+8], {%2, %3};\n\tst.shared.v2.f64 [%0+24], {%4, %5};\n\tst.shared.f64 [%0+40], %6;"
                 ::"r"(s), "d"(a), "d"(b), "d"(c), "d"(d), "d"(e), "d"(f) : "memory");
  }
}
__device__ __forceinline__ void l3_sts3(double *p, double a, double b, double c) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(p);
  if ((s & 8u) == 0) asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n\tst.shared.f64 [%0+16], %3;" ::"r"(s), "d"(a), "d"(b), "d"(c) : "memory");
  else asm volatile("st.shared.f64 [%0], %1;\n\tst.shared.v2.f64 [%0+8], {%2, %3};" ::"r"(s), "d"(a), "d"(b), "d"(c) : "memory");
}

// one block accumulator: {D0, D1, D2, S01, S02, S12, W01, W02, W12}; K_ik = S_ik + W_ik, K_ki = S_ik - W_ik, K_ii = 2 D_i
__device__ __forceinline__ void l3_acc(double (&K)[9], const double (&p)[3], const double (&m)[3], const double (&b)[3]) {
  K[0] = fma(p[0], b[0], K[0]); K[1] = fma(p[1], b[1], K[1]); K[2] = fma(p[2], b[2], K[2]);
  K[3] = fma(p[0], b[1], fma(p[1], b[0], K[3]));
  K[4] = fma(p[0], b[2], fma(p[2], b[0], K[4]));
  K[5] = fma(p[1], b[2], fma(p[2], b[1], K[5]));
  K[6] = fma(m[0], b[1], fma(-m[1], b[0], K[6]));
  K[7] = fma(m[0], b[2], fma(-m[2], b[0], K[7]));
  K[8] = fma(m[1], b[2], fma(-m[2], b[1], K[8]));
}
__device__ __forceinline__ void l3_full(const double (&K)[9], const double dl, double (&F)[3][3]) {
  F[0][0] = fma(2.0, K[0], dl); F[1][1] = fma(2.0, K[1], dl); F[2][2] = fma(2.0, K[2], dl);
  F[0][1] = K[3] + K[6]; F[1][0] = K[3] - K[6];
  F[0][2] = K[4] + K[7]; F[2][0] = K[4] - K[7];
  F[1][2] = K[5] + K[8]; F[2][1] = K[5] - K[8];
}

// adjugate (transposed cofactors) and determinant of a 3 x 3 matrix: inverse = adj / det
__device__ __forceinline__ double l3_adj(const double (&M)[9], double (&A)[9]) {
  A[0] = M[4] * M[8] - M[5] * M[7]; A[1] = M[2] * M[7] - M[1] * M[8]; A[2] = M[1] * M[5] - M[2] * M[4];
  A[3] = M[5] * M[6] - M[3] * M[8]; A[4] = M[0] * M[8] - M[2] * M[6]; A[5] = M[2] * M[3] - M[0] * M[5];
  A[6] = M[3] * M[7] - M[4] * M[6]; A[7] = M[1] * M[6] - M[0] * M[7]; A[8] = M[0] * M[4] - M[1] * M[3];
  return M[0] * A[0] + M[1] * A[3] + M[2] * A[6];
}

// CSR geometry of the row of node i of a line: parity of its first double, blocks per neighbour line
struct L3RowGeo { int par, w; };
__device__ __forceinline__ L3RowGeo l3_geo(const int lbpar, const int nl, const int i, const int nxm1) {
  L3RowGeo g;
  g.w = (i == 0 || i == nxm1) ? 2 : 3;
  g.par = (i == 0) ? lbpar : ((lbpar + nl * (3 * i - 1)) & 1);
  return g;
}

template <bool DO_R, bool DO_K, bool LAP_BUILD, bool PROF>
__global__ void __launch_bounds__(L3_THREADS, 1)
k_asm_q1_3d_lines(const L3Args P, const __grid_constant__ TabQ1 T, const __grid_constant__ MatParams mp) {
  constexpr int NARR = DO_R ? 16 : 13;          // double2 arrays per quadrature point: 4 corner pairs x 3, scalars, stress x 3
  constexpr bool STAGE = DO_K && !LAP_BUILD;    // the Jacobian rows go through the staging image and the TMA unit
  constexpr int CC = 12 * L3_NCOLP, PR = 3 * L3_NCOLP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *STG = reinterpret_cast<double *>(smem_raw);                                  // [L3_NL][L3_SLOT]
  double2 *REC = reinterpret_cast<double2 *>(STG + L3_NL * L3_SLOT);                    // [L3_QH][NARR][L3_NCOLP]
  double *XU = reinterpret_cast<double *>(REC + L3_QH * 16 * L3_NCOLP);                 // [2 planes][3][L3_NLCP] double2: X, u of the cached node lines
  double *LAPS = XU + 2 * 6 * L3_NLCP;                                                  // [3][L3_THREADS] Laplacian parts of this interval's blocks
  double *TAB = LAPS + 3 * L3_THREADS;                                                  // dN[8][8][3], w[8]
  __shared__ int s_unit;
  __shared__ long long s_lapoff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < 192; k += L3_THREADS) TAB[k] = T.dN[k / 24][(k / 3) & 7][k % 3];
  if (tid < 8) TAB[192 + tid] = T.w[tid];
  for (int k = tid; k < L3_QH * NARR; k += L3_THREADS) REC[k * L3_NCOLP + L3_ZCOL] = make_double2(0.0, 0.0);
  const double mu = LAP_BUILD ? 0.0 : mp.p[0], lam = mp.p[1];
  const int nx = P.nx, ny = P.ny, nzp = P.nzp, nxm1 = nx - 1;

  for (;;) {
    if (tid == 0) { const int uu = (int)atomicAdd(P.counter, 1u); s_unit = uu; s_lapoff = uu < P.nunits ? (long long)P.lapoff[uu] : 0ll; }
    __syncthreads();
    const int u = s_unit;
    if (u >= P.nunits) break;
    const int4 un = P.units[u];
    const int y0 = un.x, z0 = un.y, i0 = un.z, i1 = un.w;
    const int ifirst = i0 > 0 ? i0 - 1 : 0;    // a chunk that starts inside the line first runs one interval without stores
    const int ilast = (i1 == nxm1) ? nxm1 : i1 - 1;   // last node plane whose rows this unit writes

    // ---- phase-B role of this thread (the same for every unit; derived again per unit so that it does not occupy registers across the interval loop): pair type and the tile-relative line A = (ay, az) ----
    // warps {0,1} self pairs; {2,3,6,7,11} pairs with two common element columns, d = (1,0) and (0,1); {4,5,8,9,10} pairs with
    // one, d = (1,1) and (-1,1): per SM sub-partition (warp & 3) the roles add up to about the same DFMA count
    int ptype = -1, ay = 0, az = 0;
    {
      int role, ord;
      switch (warp) {
        case 0: role = 0; ord = 0; break;  case 1: role = 0; ord = 1; break;
        case 2: role = 1; ord = 0; break;  case 3: role = 1; ord = 1; break;
        case 6: role = 1; ord = 2; break;  case 7: role = 1; ord = 3; break;  case 11: role = 1; ord = 4; break;
        case 4: role = 2; ord = 0; break;  case 5: role = 2; ord = 1; break;
        case 8: role = 2; ord = 2; break;  case 9: role = 2; ord = 3; break;  default: role = 2; ord = 4; break;
      }
      int k = ord * 32 + lane;
      if (role == 0) { ptype = 0; ay = k & 7; az = k >> 3; }
      else if (role == 1) {
        if (k < 72) { ptype = 1; ay = k % 9 - 1; az = k / 9; }                       // A from y = -1: the pair (-1, 0) belongs to line 0's rows
        else if (k < 144) { k -= 72; ptype = 3; ay = k & 7; az = (k >> 3) - 1; }
      } else if (k < 159 && k != 79) {
        if (k < 79) ptype = 4; else { ptype = 2; k -= 80; }      // from a multiple of 8: the lanes of a quarter-warp read 8 consecutive columns
        if (k < 64) { ay = k & 7; az = k >> 3; }
        else if (k < 72) { ay = ptype == 4 ? -1 : 8; az = k - 65; }                  // B in the tile, A beside it
        else { az = -1; ay = ptype == 4 ? k - 72 : k - 71; }                         // ... A below it
      }
    }
    const int dy = (ptype == 1 || ptype == 4) ? 1 : (ptype == 2 ? -1 : 0), dz = ptype >= 2 ? 1 : 0;
    const bool two = (ptype == 1 || ptype == 3);
    // corner pair (dy' + 2 dz') of line A and of line B inside the first and the second common element column
    const int oa0 = (ptype == 1 ? 2 : (ptype == 4 ? 0 : 1)) * PR, ob0 = (ptype == 2 ? 2 : 3) * PR, ob1 = (ptype == 1 ? 1 : 2) * PR;   // oa1 = 0
    // row issuers: cp.async.bulk blocks its issuer while the TMA queue (about eight requests) is full, and one interval's rows
    // keep the SM's store path busy for about 4 k cycles (32 B per clock; profiles/r2_ubench_bulk.txt). The hand-off therefore
    // sits where warps idle anyway: warp 11 holds none of the 324 phase-A items and issues 32 rows during the first half of
    // phase A; the five warps of one-column pairs finish the first half of phase B long before the others and issue 6-7 rows each.
    const int oord = (warp == 4) ? 0 : (warp == 5) ? 1 : (warp == 8) ? 2 : (warp == 9) ? 3 : (warp == 10) ? 4 : -1;
    const int iline = (warp == 11) ? lane : 32 + oord * 7 + lane;
    const bool isslot = STAGE && (warp == 11 || (oord >= 0 && lane < 7 && iline < L3_NL));

    // ---- per-unit role of this thread in phase B ----
    // geoA / geoB: where the pair's blocks go in the staging image -- staging offset of the line | neighbour lines << 16 |
    // rank of the other line among them << 20 | parity of the line's first block row << 24; -1 = that line is not this tile's
    int geoA = -1, geoB = -1;
    int c0 = L3_ZCOL, c1 = L3_ZCOL, c2 = L3_ZCOL, c3 = L3_ZCOL;     // element columns that hold both lines
    int nodeA0 = 0;
    if (ptype >= 0 && (DO_K || ptype == 0)) {
      const int y = y0 + ay, z = z0 + az, y2 = y + dy, z2 = z + dz;
      const bool ex = y >= 0 && y < ny && z >= 0 && z < nzp && y2 >= 0 && y2 < ny && z2 < nzp;
      const bool stA = ex && ay >= 0 && ay < L3_TY && az >= 0 && P.pown[z] != 0;
      const bool stB = ex && ptype != 0 && ay + dy >= 0 && ay + dy < L3_TY && az + dz < L3_TZ && P.pown[z2] != 0;
      if (stA || stB) {
        auto col_of = [&](int ey, int ez) -> int {
          return (ey >= 0 && ey < ny - 1 && ez >= 0 && ez < nzp - 1) ? (ez - z0 + 1) * L3_CY + (ey - y0 + 1) : L3_ZCOL;
        };
        if (ptype == 0) { c0 = col_of(y - 1, z - 1); c1 = col_of(y, z - 1); c2 = col_of(y - 1, z); c3 = col_of(y, z); }
        else if (ptype == 1) { c0 = col_of(y, z - 1); c1 = col_of(y, z); }
        else if (ptype == 2) { c0 = col_of(y - 1, z); }
        else if (ptype == 3) { c0 = col_of(y - 1, z); c1 = col_of(y, z); }
        else { c0 = col_of(y, z); }
        if (stA) {
          const L3Line *a = P.info + (z * ny + y);
          geoA = ((az * L3_TY + ay) * L3_SLOT) | (a->nl << 16) | ((int)a->rank[(dz + 1) * 3 + (dy + 1)] << 20) | ((int)(a->lb & 1) << 24);
          nodeA0 = P.pbase[z] + y * nx;
        }
        if (stB) {
          const L3Line *b = P.info + (z2 * ny + y2);
          geoB = (((az + dz) * L3_TY + ay + dy) * L3_SLOT) | (b->nl << 16) | ((int)b->rank[(1 - dz) * 3 + (1 - dy)] << 20) | ((int)(b->lb & 1) << 24);
        }
      }
    }
    const bool on = (geoA & geoB) >= 0;     // at least one of them is not -1
    // first double of the pair's piece in the row of node ip, and the distance between the three scalar rows
    auto piece = [&](const int geo, const int ip, int &rs) -> double * {
      const int nl = (geo >> 16) & 15, w = (ip == 0 || ip == nxm1) ? 2 : 3;
      const int par = (ip == 0) ? (geo >> 24) & 1 : ((geo >> 24) + nl * (3 * ip - 1)) & 1;
      rs = nl * 3 * w;
      return STG + (geo & 0xffff) + par + ((geo >> 20) & 15) * w * 3;
    };
    // issuer: does its line exist and belong to this rank?
    bool iss = false;
    if (isslot) {
      const int y = y0 + (iline & (L3_TY - 1)), z = z0 + iline / L3_TY;
      iss = y < ny && z < nzp && P.pown[z] != 0;
    }
    int pend = -1;                 // node plane whose rows still wait for their hand-off to the TMA unit

    // accumulators: Kd = (A_i, B_i), Ku = (A_i, B_i+1), Kl = (A_i+1, B_i), Kn = (A_i+1, B_i+1); self: Kd, Kn hold 6 symmetric sums.
    // The two x-off-diagonal blocks finished in interval i-1, (A_i, B_i-1) and (A_i-1, B_i), belong to the rows of node plane
    // i: they wait in the CTA's carry scratch (global, L2-resident, one coalesced column per thread) -- not in registers --
    // and cp.async drops them into the staging image while the second half of interval i is computed.
    // (self pairs have no use for Kl: its first six entries carry their internal-force sums, node i and node i+1)
    double Kd[9], Ku[9], Kl[9], Kn[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Kd[k] = 0.0; Ku[k] = 0.0; Kl[k] = 0.0; Kn[k] = 0.0; }
    // carried blocks -> dx = -1 pieces of the rows of node plane ip (asynchronous; completion: l3_wait_all)
    auto fetch_carry = [&](const int ip) {
      const double *cry = P.carry + (size_t)blockIdx.x * 18 * L3_THREADS + tid;
      if (geoA >= 0) {
        int rs;
        double *p = piece(geoA, ip, rs);
#pragma unroll
        for (int k = 0; k < 9; ++k) l3_cp8(p + (k / 3) * rs + k % 3, cry + k * L3_THREADS);
      }
      if (geoB >= 0) {
        int rs;
        double *p = piece(geoB, ip, rs);
#pragma unroll
        for (int k = 0; k < 9; ++k) l3_cp8(p + (k / 3) * rs + k % 3, cry + (9 + k) * L3_THREADS);
      }
    };

    auto load_plane = [&](int ip) {            // X, u of node plane ip of the cached lines -> ring slot ip & 1
      if (ip > nxm1) return;
      const int slot = ip & 1;
      if (oord < 0) return;                    // the one-column pair warps fetch: they have the slack (phase B is half as long for them)
      for (int k = oord * 32 + lane; k < 2 * L3_NLC; k += 5 * 32) {
        const int c = k >> 1, half = k & 1;
        const int y = y0 - 1 + c % L3_NY2, z = z0 - 1 + c / L3_NY2;
        if (y < 0 || y >= ny || z < 0 || z >= nzp) continue;
        const int64_t node = (int64_t)P.pbase[z] + (int64_t)y * nx + ip;
        const double *src = (half ? P.U : P.coords) + 3 * node;
        // component j of {X, u} of a line sits in double2 array j >> 1, half j & 1: phase A takes a node with three 16-byte loads
        double *dst = XU + 2 * (slot * 3 * L3_NLCP + c_l3_xrow[c / L3_NY2] + c % L3_NY2);
#pragma unroll
        for (int d = 0; d < 3; ++d) { const int j = 3 * half + d; l3_cp8(dst + 2 * (j >> 1) * L3_NLCP + (j & 1), src + d); }
      }
    };

    // the row of node plane ip of line `iline`: 16-byte aligned part by one bulk copy. A row that starts on an odd double
    // takes the last double of the previous row with it (parked in the slot's spare double 245 when that row was handed over);
    // the first and the last row of the unit store their odd end singly. Stateless: everything follows from the line's table.
    auto flush_row = [&](const int ip) {
      const L3Line *a = P.info + ((z0 + iline / L3_TY) * ny + y0 + (iline & (L3_TY - 1)));
      const int64_t alb = a->lb;
      const int anl = a->nl;
      const L3RowGeo g = l3_geo((int)(alb & 1), anl, ip, nxm1);
      const int64_t g0 = 9 * (alb + (int64_t)anl * (ip == 0 ? 0 : 3 * ip - 1));   // first double of the row in vals
      double *sm = STG + iline * L3_SLOT;                                            // the row sits at sm[par .. par + len)
      int lo = g.par, hi = g.par + 9 * anl * g.w;
      if (lo) {
        if (ip > i0) { sm[0] = sm[L3_SLOT - 1]; lo = 0; l3_fence_async(); }
        else { l3_stg1(P.vals + g0, sm[1]); lo = 2; }
      }
      if (hi & 1) {
        --hi;
        if (ip < ilast) sm[L3_SLOT - 1] = sm[hi];
        else l3_stg1(P.vals + (g0 - g.par + hi), sm[hi]);
      }
      if (hi > lo) l3_bulk_store(P.vals + (g0 - g.par + lo), sm + lo, (unsigned)(hi - lo) * 8u);
      l3_bulk_commit();
    };

    load_plane(ifirst); load_plane(ifirst + 1);
    l3_commit(); l3_wait_all();
    __syncthreads();

    for (int i = ifirst; i < i1; ++i) {
      const long long tp0 = PROF ? clock64() : 0;
      long long tA = 0, tB = 0, tAw = 0, tFc = 0, tMid = 0;
      const bool store = i >= i0 && !P.dbg_nostore;
      if (!LAP_BUILD && DO_K && on) {            // this interval's Laplacian parts -> shared memory, read when the blocks are finished
        const double *lq = P.lap + s_lapoff + tid + (int64_t)(i - ifirst) * 3 * L3_NPT;
        l3_cp8(LAPS + tid, lq); l3_cp8(LAPS + L3_THREADS + tid, lq + L3_NPT); l3_cp8(LAPS + 2 * L3_THREADS + tid, lq + 2 * L3_NPT);
      }
      l3_commit();
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const long long ta0 = PROF ? clock64() : 0;
        // ---------------- phase A: quadrature-point records 4 half .. 4 half + 3 of interval i ----------------
        if (half == 0 && iss && pend >= 0 && warp == 11) flush_row(pend);    // warp 11 first: it has no phase-A item
        if (tid < L3_QH * L3_NCOL) {
          const int slo = i & 1, shi = slo ^ 1;
          const int ql = tid / L3_NCOL, col = tid - ql * L3_NCOL, q = half * L3_QH + ql;
          const int cyi = col % L3_CY, czi = col / L3_CY;
          const int ey = y0 - 1 + cyi, ez = z0 - 1 + czi;
          if (ey >= 0 && ey < ny - 1 && ez >= 0 && ez < nzp - 1) {
            const double *tq = TAB + q * 24;
            const int xr[2] = {c_l3_xrow[czi] + cyi, c_l3_xrow[czi + 1] + cyi};
            double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              const int cl = xr[a >> 2] + ((a >> 1) & 1);
              const double2 *xs = reinterpret_cast<const double2 *>(XU) + ((a & 1) ? shi : slo) * 3 * L3_NLCP + cl;
              const double dn0 = tq[a * 3], dn1 = tq[a * 3 + 1], dn2 = tq[a * 3 + 2];
              const double2 v0 = xs[0], v1 = xs[L3_NLCP];
              const double x[3] = {v0.x, v0.y, v1.x};
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                J[d * 3 + 0] = fma(x[d], dn0, J[d * 3 + 0]); J[d * 3 + 1] = fma(x[d], dn1, J[d * 3 + 1]); J[d * 3 + 2] = fma(x[d], dn2, J[d * 3 + 2]);
              }
              if (!LAP_BUILD) {
                const double2 v2 = xs[2 * L3_NLCP];
                const double v[3] = {v1.y, v2.x, v2.y};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  H[d * 3 + 0] = fma(v[d], dn0, H[d * 3 + 0]); H[d * 3 + 1] = fma(v[d], dn1, H[d * 3 + 1]); H[d * 3 + 2] = fma(v[d], dn2, H[d * 3 + 2]);
                }
              }
            }
            // one division for both inverses: r = 1 / (det J det M), J^-1 = adj J (det M r), M^-1 = adj M (det J r), det F = det M^2 r.
            // Order chosen for register pressure (the pair accumulators stay live through this phase): H adj(J) is formed
            // before M overwrites J, so that J, H and adj J are dead when adj M is built.
            double Mi[9], Hp[9];
            double cs, cw, lnJ = 0.0, c1 = 0.0, c2 = 0.0, wdet;
            {
              double Ji[9];
              const double detJ = l3_adj(J, Ji);
              wdet = TAB[192 + q] * detJ;
              if (LAP_BUILD) {
                const double id = 1.0 / detJ;
#pragma unroll
                for (int k = 0; k < 9; ++k) Mi[k] = Ji[k] * id;
                cs = 0.5 * wdet; cw = 0.5 * wdet;
              } else {
                if (DO_R) {
#pragma unroll
                  for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Hp[r * 3 + c] = fma(H[r * 3], Ji[c], fma(H[r * 3 + 1], Ji[3 + c], H[r * 3 + 2] * Ji[6 + c]));
                }
#pragma unroll
                for (int k = 0; k < 9; ++k) J[k] += H[k];                    // J <- M = J + H
                const double detM = l3_adj(J, Mi);
                const double r = 1.0 / (detJ * detM);
                const double sj = detM * r, sm_ = detJ * r;
#pragma unroll
                for (int k = 0; k < 9; ++k) Mi[k] *= sm_;
                if (DO_R) {
#pragma unroll
                  for (int k = 0; k < 9; ++k) Hp[k] *= sj;                   // Hp = H J^-1 = F - I
                }
                lnJ = log(detM * sj);
                c1 = wdet * mu; c2 = wdet * lam;
                const double c3 = c1 - c2 * lnJ;
                cs = 0.5 * (c2 + c3); cw = 0.5 * (c2 - c3);
              }
            }
            double2 *rq = REC + ql * NARR * L3_NCOLP + col;
            rq[CC] = make_double2(cs, cw);
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
              double g[2][3];
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                const int a = 2 * pr + s;
                const double dn0 = tq[a * 3], dn1 = tq[a * 3 + 1], dn2 = tq[a * 3 + 2];
#pragma unroll
                for (int k = 0; k < 3; ++k) g[s][k] = fma(dn0, Mi[k], fma(dn1, Mi[3 + k], dn2 * Mi[6 + k]));
              }
              rq[(pr * 3 + 0) * L3_NCOLP] = make_double2(g[0][0], g[0][1]);
              rq[(pr * 3 + 1) * L3_NCOLP] = make_double2(g[0][2], g[1][0]);
              rq[(pr * 3 + 2) * L3_NCOLP] = make_double2(g[1][1], g[1][2]);
            }
            if (DO_R && !LAP_BUILD) {
              // Kirchhoff stress times w detJ: tau = mu (b - I) + lam lnJ I, b - I = Hp + Hp^T + Hp Hp^T, Hp = H J^-1 = F - I
              const double cl = c2 * lnJ;
              auto ee = [&](int r, int c) {
                return (Hp[r * 3 + c] + Hp[c * 3 + r]) + fma(Hp[r * 3], Hp[c * 3], fma(Hp[r * 3 + 1], Hp[c * 3 + 1], Hp[r * 3 + 2] * Hp[c * 3 + 2]));
              };
              rq[13 * L3_NCOLP] = make_double2(fma(c1, ee(0, 0), cl), fma(c1, ee(1, 1), cl));
              rq[14 * L3_NCOLP] = make_double2(fma(c1, ee(2, 2), cl), c1 * ee(0, 1));
              rq[15 * L3_NCOLP] = make_double2(c1 * ee(0, 2), c1 * ee(1, 2));
            }
          }
        }
        // last interval's bulk copies have long finished reading the staging image; make that a fact before anyone rewrites it
        if (half == 1 && isslot) l3_bulk_wait_read();
        if (PROF) tAw += clock64() - ta0;
        // The fence is required, not decoration: without it the phase-B loads of some warps returned the previous
        // interval's records although every STS above precedes this barrier in program order (found on a 3 x 2 x 3-node
        // mesh by the parity tests; profiles/r2_lines3d_notes.md). membar.cta costs nothing here.
        __threadfence_block();
        __syncthreads();
        if (PROF) tA += clock64() - ta0;
        // Fetches sit off the critical path: all warps leaving the barrier and issuing ~20 cp.async each cost 2.7 k cycles per
        // interval (tools/lines_probe.py). Node plane i+2 is fetched by the one-column warps only; the self-pair warps (the
        // long pole, 9 copies each) fetch their carried blocks here, the others after their pair loop, where the warps have
        // drifted apart and the copies of one overlap the arithmetic of the others.
        const bool fetch = half == 1 && STAGE && on && store && i > 0;
        if (half == 1) {
          load_plane(i + 2);       // both halves have read node plane i: its slot takes plane i+2
          if (fetch && ptype == 0) fetch_carry(i);
          l3_commit();
        }
        const long long tb0 = PROF ? clock64() : 0;
        if (PROF) tFc += tb0 - ta0;
        // ---------------- phase B: line pairs, four quadrature points ----------------
        if (on) {
          if (ptype == 0) {
            // two of the line's four element columns per loop trip (a missing column reads the all-zero record): twice the
            // independent loads and FMAs per trip -- the self pairs are the long pole of phase B and latency-bound
            auto self_one = [&](const double2 *rq, const int oa) {
              const double2 a0 = rq[oa], a1 = rq[oa + L3_NCOLP], a2 = rq[oa + 2 * L3_NCOLP];
              const double2 cc = rq[CC];
              const double glo[3] = {a0.x, a0.y, a1.x}, ghi[3] = {a1.y, a2.x, a2.y};
              if (DO_K) {
                const double plo[3] = {cc.x * glo[0], cc.x * glo[1], cc.x * glo[2]};
                const double mlo[3] = {cc.y * glo[0], cc.y * glo[1], cc.y * glo[2]};
                const double phi[3] = {cc.x * ghi[0], cc.x * ghi[1], cc.x * ghi[2]};
                // (lo, lo) and (hi, hi): symmetric, K_ik = 2 cs g_i g_k
                Kd[0] = fma(plo[0], glo[0], Kd[0]); Kd[1] = fma(plo[1], glo[1], Kd[1]); Kd[2] = fma(plo[2], glo[2], Kd[2]);
                Kd[3] = fma(plo[0], glo[1], Kd[3]); Kd[4] = fma(plo[0], glo[2], Kd[4]); Kd[5] = fma(plo[1], glo[2], Kd[5]);
                Kn[0] = fma(phi[0], ghi[0], Kn[0]); Kn[1] = fma(phi[1], ghi[1], Kn[1]); Kn[2] = fma(phi[2], ghi[2], Kn[2]);
                Kn[3] = fma(phi[0], ghi[1], Kn[3]); Kn[4] = fma(phi[0], ghi[2], Kn[4]); Kn[5] = fma(phi[1], ghi[2], Kn[5]);
                l3_acc(Ku, plo, mlo, ghi);
              }
              if (DO_R && !LAP_BUILD) {
                const double2 t0 = rq[13 * L3_NCOLP], t1 = rq[14 * L3_NCOLP], t2 = rq[15 * L3_NCOLP];
                // tau = {xx, yy | zz, xy | xz, yz}
                Kl[0] = fma(t0.x, glo[0], fma(t1.y, glo[1], fma(t2.x, glo[2], Kl[0])));
                Kl[1] = fma(t1.y, glo[0], fma(t0.y, glo[1], fma(t2.y, glo[2], Kl[1])));
                Kl[2] = fma(t2.x, glo[0], fma(t2.y, glo[1], fma(t1.x, glo[2], Kl[2])));
                Kl[3] = fma(t0.x, ghi[0], fma(t1.y, ghi[1], fma(t2.x, ghi[2], Kl[3])));
                Kl[4] = fma(t1.y, ghi[0], fma(t0.y, ghi[1], fma(t2.y, ghi[2], Kl[4])));
                Kl[5] = fma(t2.x, ghi[0], fma(t2.y, ghi[1], fma(t1.x, ghi[2], Kl[5])));
              }
            };
            auto self_two = [&](const int colA, const int oaA, const int colB, const int oaB) {
              const double2 *ba = REC + colA, *bb = REC + colB;
#pragma unroll kL3SelfUnroll
              for (int q = 0; q < L3_QH; ++q) {
                self_one(ba + q * NARR * L3_NCOLP, oaA);
                self_one(bb + q * NARR * L3_NCOLP, oaB);
              }
            };
            self_two(c0, 3 * PR, c1, 2 * PR); self_two(c2, PR, c3, 0);
          } else if (DO_K) {
            // one loop over the (column, point) pairs of the pair's one or two columns; the A-side operands and the scalars
            // of the next pair are loaded as soon as the products that consume the current ones have been formed
            const int n = two ? 2 * L3_QH : L3_QH;
            const double2 *rq = REC + c0;
            int oa = oa0, ob = ob0;
            double2 a0 = rq[oa], a1 = rq[oa + L3_NCOLP], a2 = rq[oa + 2 * L3_NCOLP], cc = rq[CC];
#pragma unroll kL3PairUnroll
            for (int it = 0; it < n; ++it) {
              const double2 b0 = rq[ob], b1 = rq[ob + L3_NCOLP], b2 = rq[ob + 2 * L3_NCOLP];
              const double plo[3] = {cc.x * a0.x, cc.x * a0.y, cc.x * a1.x}, mlo[3] = {cc.y * a0.x, cc.y * a0.y, cc.y * a1.x};
              const double phi[3] = {cc.x * a1.y, cc.x * a2.x, cc.x * a2.y}, mhi[3] = {cc.y * a1.y, cc.y * a2.x, cc.y * a2.y};
              if (it + 1 == L3_QH) { rq = REC + c1; oa = 0; ob = ob1; }       // second column (read past the end only when n == 4: the zero column)
              else rq += NARR * L3_NCOLP;
              if (it + 1 < n) { a0 = rq[oa]; a1 = rq[oa + L3_NCOLP]; a2 = rq[oa + 2 * L3_NCOLP]; cc = rq[CC]; }
              const double blo[3] = {b0.x, b0.y, b1.x}, bhi[3] = {b1.y, b2.x, b2.y};
              l3_acc(Kd, plo, mlo, blo); l3_acc(Ku, plo, mlo, bhi);
              l3_acc(Kl, phi, mhi, blo); l3_acc(Kn, phi, mhi, bhi);
            }
          }
          if (fetch && ptype != 0) { fetch_carry(i); l3_commit(); }
        }
        if (PROF) tB += clock64() - tb0;
        if (half == 0) {
          if (iss && pend >= 0 && warp != 11) flush_row(pend);
          pend = -1;
          __threadfence_block(); __syncthreads();      // the records are rewritten by the second half
          if (PROF) tMid += clock64() - tb0;
        }
      }
      // ---- interval i completes (A_i, B_i), (A_i, B_i+1), (A_i+1, B_i): with the two blocks carried from interval i-1 they
      //      fill the rows of node plane i, direct and transposed ----
      const long long tf0 = PROF ? clock64() : 0;
      l3_wait_all();      // node plane i+2 and the Laplacian parts are in; the carried blocks have left the scratch rewritten below
      const long long tf1 = PROF ? clock64() : 0;
      if (on) {
        const int tslot = (i - ifirst) * 3 * L3_NPT;
        double *lapu = P.lap + s_lapoff + tid;
        double *cry = P.carry + (size_t)blockIdx.x * 18 * L3_THREADS + tid;
        const int skip = i > 0 ? 3 : 0;      // the dx = -1 piece comes from the carry scratch
        if (ptype == 0) {
          if (DO_K) {
            double Fu[3][3];
            l3_full(Ku, 0.0, Fu);
            if (LAP_BUILD) {
              lapu[tslot] = 2.0 * (Kd[0] + Kd[1] + Kd[2]); lapu[tslot + L3_NPT] = Fu[0][0] + Fu[1][1] + Fu[2][2]; lapu[tslot + 2 * L3_NPT] = 0.0;
            } else {
              const double l0 = mu * LAPS[tid], l1 = mu * LAPS[L3_THREADS + tid];
              Fu[0][0] += l1; Fu[1][1] += l1; Fu[2][2] += l1;
              if (store) {
                const double D[3][3] = {{fma(2.0, Kd[0], l0), 2.0 * Kd[3], 2.0 * Kd[4]}, {2.0 * Kd[3], fma(2.0, Kd[1], l0), 2.0 * Kd[5]},
                                        {2.0 * Kd[4], 2.0 * Kd[5], fma(2.0, Kd[2], l0)}};
                int rs;
                double *p = piece(geoA, i, rs) + skip;
#pragma unroll
                for (int r = 0; r < 3; ++r, p += rs) l3_sts6(p, D[r][0], D[r][1], D[r][2], Fu[r][0], Fu[r][1], Fu[r][2]);
              }
              // (A_i+1, A_i) = (A_i, A_i+1)^T waits for the row of node i+1
#pragma unroll
              for (int k = 0; k < 9; ++k) cry[k * L3_THREADS] = Fu[k % 3][k / 3];
            }
          }
          if (DO_R && !LAP_BUILD && store) l3_st3(P.R + 3 * ((int64_t)nodeA0 + i), Kl[0], Kl[1], Kl[2]);
        } else if (DO_K) {
          double Fd[3][3], Fu[3][3], Fl[3][3];
          if (LAP_BUILD) {
            l3_full(Kd, 0.0, Fd); l3_full(Ku, 0.0, Fu); l3_full(Kl, 0.0, Fl);
            lapu[tslot] = Fd[0][0] + Fd[1][1] + Fd[2][2]; lapu[tslot + L3_NPT] = Fu[0][0] + Fu[1][1] + Fu[2][2];
            lapu[tslot + 2 * L3_NPT] = Fl[0][0] + Fl[1][1] + Fl[2][2];
          } else {
            l3_full(Kd, mu * LAPS[tid], Fd); l3_full(Ku, mu * LAPS[L3_THREADS + tid], Fu); l3_full(Kl, mu * LAPS[2 * L3_THREADS + tid], Fl);
            if (store) {
              if (geoA >= 0) {
                int rs;
                double *p = piece(geoA, i, rs) + skip;
#pragma unroll
                for (int r = 0; r < 3; ++r, p += rs) l3_sts6(p, Fd[r][0], Fd[r][1], Fd[r][2], Fu[r][0], Fu[r][1], Fu[r][2]);
              }
              if (geoB >= 0) {
                int rs;
                double *p = piece(geoB, i, rs) + skip;
#pragma unroll
                for (int r = 0; r < 3; ++r, p += rs) l3_sts6(p, Fd[0][r], Fd[1][r], Fd[2][r], Fl[0][r], Fl[1][r], Fl[2][r]);
              }
            }
            // (A_i+1, B_i) waits for the row of A's node i+1, (B_i+1, A_i) = (A_i, B_i+1)^T for the row of B's
            if (geoA >= 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) cry[k * L3_THREADS] = Fl[k / 3][k % 3];
            }
            if (geoB >= 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) cry[(9 + k) * L3_THREADS] = Fu[k % 3][k / 3];
            }
          }
        }
        // shift: node i+1 becomes node i
#pragma unroll
        for (int k = 0; k < 9; ++k) { Kd[k] = Kn[k]; Kn[k] = 0.0; Ku[k] = 0.0; }
        const bool self = ptype == 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { Kl[k] = self ? Kl[3 + k] : 0.0; Kl[3 + k] = 0.0; Kl[6 + k] = 0.0; }
      }
      const long long tf2 = PROF ? clock64() : 0;
      if (STAGE) l3_fence_async();
      __threadfence_block();
      __syncthreads();
      const long long tf3 = PROF ? clock64() : 0;
      // the rows are complete; their hand-off to the TMA unit waits for the idle time of the next interval (see the issuers above)
      pend = (STAGE && store) ? i : -1;
      if (PROF && lane == 0) {      // per warp: own phase-B work; warp 0 also: phase A to its barrier, the rest, cp.async wait, finalize, last barrier, row hand-off
        const long long tp3 = clock64();
        atomicAdd(P.prof + 8 + warp, (unsigned long long)tB);
        if (tid == 0) {
          atomicAdd(P.prof + 0, (unsigned long long)tA); atomicAdd(P.prof + 1, (unsigned long long)(tp3 - tp0 - tA)); atomicAdd(P.prof + 2, 1ull);
          atomicAdd(P.prof + 3, (unsigned long long)(tf1 - tf0)); atomicAdd(P.prof + 4, (unsigned long long)(tf2 - tf1));
          atomicAdd(P.prof + 5, (unsigned long long)(tf3 - tf2)); atomicAdd(P.prof + 6, (unsigned long long)(tp3 - tf3));
          // own phase-A work before its barrier; phase A + barrier + node-plane / carry fetch issue; first phase B + hand-off + its barrier
          atomicAdd(P.prof + 20, (unsigned long long)tAw); atomicAdd(P.prof + 21, (unsigned long long)tFc); atomicAdd(P.prof + 22, (unsigned long long)tMid);
        }
      }
    }
    if (iss && pend >= 0) flush_row(pend);       // the unit's last interval: hand its rows over now
    pend = -1;
    // ---- the last node of the line: its row has no further interval ----
    if (i1 == nxm1) {
      const int i = nxm1;
      if (STAGE) {
        if (isslot) l3_bulk_wait_read();
        __syncthreads();
      }
      if (on) {
        double lp0 = 0.0;
        const int tslot = (i1 - ifirst) * 3 * L3_NPT;
        double *lapu = P.lap + s_lapoff + tid;
        if (!LAP_BUILD && DO_K) lp0 = lapu[tslot];
        if (STAGE && !P.dbg_nostore) { fetch_carry(i); l3_commit(); }      // the dx = -1 pieces
        if (ptype == 0) {
          if (DO_K) {
            if (LAP_BUILD) {
              lapu[tslot] = 2.0 * (Kd[0] + Kd[1] + Kd[2]);
            } else if (!P.dbg_nostore) {
              const double l0 = mu * lp0;
              int rs;
              double *p = piece(geoA, i, rs) + 3;
              l3_sts3(p, fma(2.0, Kd[0], l0), 2.0 * Kd[3], 2.0 * Kd[4]);
              l3_sts3(p + rs, 2.0 * Kd[3], fma(2.0, Kd[1], l0), 2.0 * Kd[5]);
              l3_sts3(p + 2 * rs, 2.0 * Kd[4], 2.0 * Kd[5], fma(2.0, Kd[2], l0));
            }
          }
          if (DO_R && !LAP_BUILD && !P.dbg_nostore) l3_st3(P.R + 3 * ((int64_t)nodeA0 + i), Kl[0], Kl[1], Kl[2]);
        } else if (DO_K) {
          double Fd[3][3];
          l3_full(Kd, mu * lp0, Fd);
          if (LAP_BUILD) {
            lapu[tslot] = Fd[0][0] + Fd[1][1] + Fd[2][2];
          } else if (!P.dbg_nostore) {
            if (geoA >= 0) {
              int rs;
              double *p = piece(geoA, i, rs) + 3;
#pragma unroll
              for (int r = 0; r < 3; ++r, p += rs) l3_sts3(p, Fd[r][0], Fd[r][1], Fd[r][2]);
            }
            if (geoB >= 0) {
              int rs;
              double *p = piece(geoB, i, rs) + 3;
#pragma unroll
              for (int r = 0; r < 3; ++r, p += rs) l3_sts3(p, Fd[0][r], Fd[1][r], Fd[2][r]);
            }
          }
        }
        l3_wait_all();
      }
      if (STAGE) {
        l3_fence_async();
        __threadfence_block();
        __syncthreads();
        if (iss && !P.dbg_nostore) flush_row(i);
      }
    }
    // the staging image is next rewritten at the end of the next unit's first interval; its issuers wait for these copies
    // in the second half of that interval (above), and the barrier that follows holds everyone else back
  }
  if (STAGE) l3_bulk_wait_read();       // (a thread without outstanding bulk groups returns at once)
  // the last CTA to run out of units rearms the counters for the next launch
  if (tid == 0) {
    __threadfence();
    const unsigned d = atomicAdd(P.counter + 1, 1u);
    if (d == gridDim.x - 1) { P.counter[0] = 0u; P.counter[1] = 0u; __threadfence(); }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
void nls_lines3_free(nls_context *h) {
  nls_mg_free(h);
  Lines3Plan &p = h->lplan;
  if (p.d_pbase) cudaFree(p.d_pbase);
  if (p.d_pown) cudaFree(p.d_pown);
  if (p.d_info) cudaFree(p.d_info);
  if (p.d_units) cudaFree(p.d_units);
  if (p.d_lapoff) cudaFree(p.d_lapoff);
  if (p.d_lap) cudaFree(p.d_lap);
  if (p.d_counter) cudaFree(p.d_counter);
  if (p.d_carry) cudaFree(p.d_carry);
  p = Lines3Plan();
}

static size_t lines3_smem(bool) {   // one layout for every instantiation: staging image, records (16 arrays), node ring, tables
  return (size_t)L3_NL * L3_SLOT * 8 + (size_t)L3_QH * 16 * L3_NCOLP * 16 + (size_t)2 * 6 * L3_NLCP * 8 + (size_t)3 * L3_THREADS * 8 + 200 * 8;
}

// Detects grid topology (x-fastest lines, planes with arbitrary base ids: full grids and slab-local meshes), builds the
// per-line addressing tables, checks them against the CSR pattern entry by entry, and lists the work units. Leaves
// p.ok false (and the atomic scatter in use) for any other mesh.
int nls_lines3_build(nls_context *h) {
  nls_lines3_free(h);
  Lines3Plan &p = h->lplan;
  if (h->dim != 3 || h->nen != 8 || h->nel == 0 || h->n_owned == 0) return NLS_OK;
  if (lines3_smem(true) + 1024 > (size_t)h->max_optin_smem) return NLS_OK;
  const int32_t *c = h->h_conn.data();
  const int64_t nel = h->nel, nn = h->nn;
  if (c[1] != c[0] + 1) return NLS_OK;
  const int64_t nx = (int64_t)c[2] - c[0];
  if (nx < 2 || nx > 1000000 || nel % (nx - 1)) return NLS_OK;
  const int64_t epr = nx - 1, rows = nel / epr;
  int64_t nyr = 1;
  while (nyr < rows && c[nyr * epr * 8] == c[0] + nyr * nx && c[nyr * epr * 8 + 4] == c[4] + nyr * nx) ++nyr;
  const int64_t ny = nyr + 1, epl = epr * nyr;
  if (nel % epl) return NLS_OK;
  const int64_t nlay = nel / epl, nzp = nlay + 1, plane = nx * ny;
  if (nn != plane * nzp || nzp > 1000000 || ny > 1000000) return NLS_OK;
  std::vector<int32_t> pbase((size_t)nzp);
  for (int64_t k = 0; k < nlay; ++k) pbase[k] = c[k * epl * 8];
  pbase[nlay] = c[(nlay - 1) * epl * 8 + 4];
  for (int64_t k = 0; k < nzp; ++k) if (pbase[k] < 0 || (int64_t)pbase[k] + plane > nn) return NLS_OK;
  {
    std::vector<int32_t> s(pbase);
    std::sort(s.begin(), s.end());
    for (int64_t k = 1; k < nzp; ++k) if ((int64_t)s[k] - s[k - 1] < plane) return NLS_OK;
  }
  for (int64_t k = 0; k < nlay; ++k)
    for (int64_t j = 0; j < nyr; ++j)
      for (int64_t i = 0; i < epr; ++i) {
        const int32_t *ce = c + ((k * nyr + j) * epr + i) * 8;
        const int64_t n0 = pbase[k] + j * nx + i, n1 = pbase[k + 1] + j * nx + i;
        if (ce[0] != n0 || ce[1] != n0 + 1 || ce[2] != n0 + nx || ce[3] != n0 + nx + 1 ||
            ce[4] != n1 || ce[5] != n1 + 1 || ce[6] != n1 + nx || ce[7] != n1 + nx + 1) return NLS_OK;
      }
  std::vector<uint8_t> pown((size_t)nzp);
  for (int64_t k = 0; k < nzp; ++k) {
    pown[k] = pbase[k] < h->n_owned;
    if (pown[k] && (int64_t)pbase[k] + plane > h->n_owned) return NLS_OK;
  }
  // per-line tables, verified against the pattern
  std::vector<L3Line> info((size_t)(nzp * ny));
  const int64_t *bp = h->h_browptr.data();
  const int32_t *bc = h->h_bcol.data();
  for (int64_t z = 0; z < nzp; ++z)
    for (int64_t y = 0; y < ny; ++y) {
      L3Line &li = info[z * ny + y];
      std::memset(&li, 0, sizeof(li));
      for (int k = 0; k < 9; ++k) li.rank[k] = -1;
      if (!pown[z]) continue;
      std::pair<int64_t, int> nb[9];
      int nl = 0;
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy) {
          const int64_t z2 = z + dz, y2 = y + dy;
          if (z2 < 0 || z2 >= nzp || y2 < 0 || y2 >= ny) continue;
          nb[nl++] = {(int64_t)pbase[z2] + y2 * nx, (dz + 1) * 3 + (dy + 1)};
        }
      std::sort(nb, nb + nl);
      for (int k = 0; k < nl; ++k) li.rank[nb[k].second] = (int8_t)k;
      li.nl = nl;
      const int64_t first = (int64_t)pbase[z] + y * nx;
      li.lb = bp[first];
      for (int64_t i = 0; i < nx; ++i) {
        const int64_t a = first + i;
        const int w = (i == 0 || i == nx - 1) ? 2 : 3;
        const int64_t b0 = li.lb + (int64_t)nl * (i == 0 ? 0 : 3 * i - 1);
        if (bp[a] != b0 || bp[a + 1] - bp[a] != (int64_t)nl * w) return NLS_OK;
        for (int k = 0; k < nl; ++k)
          for (int dx = (i == 0 ? 0 : -1); dx <= (i == nx - 1 ? 0 : 1); ++dx)
            if (bc[b0 + k * w + dx + (i == 0 ? 0 : 1)] != nb[k].first + i + dx) return NLS_OK;
      }
    }
  // work units: tiles of lines x chunks of x-intervals, largest first
  const int64_t nxi = nx - 1;
  std::vector<std::pair<int64_t, int4>> units;
  int64_t ntiles = 0;
  for (int64_t z0 = 0; z0 < nzp; z0 += L3_TZ)
    for (int64_t y0 = 0; y0 < ny; y0 += L3_TY) {
      bool work = false;
      for (int64_t z = z0; z < std::min<int64_t>(z0 + L3_TZ, nzp); ++z) work = work || pown[z];   // a tile writes the rows of its own lines only
      if (work) ++ntiles;
    }
  if (ntiles == 0) return NLS_OK;
  int64_t nch = std::max<int64_t>(1, (8 * (int64_t)h->sm_count + ntiles - 1) / ntiles);
  int64_t clen = std::max<int64_t>(std::min<int64_t>(24, nxi), (nxi + nch - 1) / nch);
  nch = (nxi + clen - 1) / clen;
  for (int64_t z0 = 0; z0 < nzp; z0 += L3_TZ)
    for (int64_t y0 = 0; y0 < ny; y0 += L3_TY) {
      int64_t nlines = 0;
      for (int64_t z = z0; z < std::min<int64_t>(z0 + L3_TZ, nzp); ++z)
        if (pown[z]) nlines += std::min<int64_t>(L3_TY, ny - y0);
      if (nlines == 0) continue;
      for (int64_t ch = 0; ch < nch; ++ch) {
        const int64_t i0 = ch * clen, i1 = std::min(nxi, i0 + clen);
        units.push_back({nlines * (i1 - i0 + (i0 > 0 ? 1 : 0)), make_int4((int)y0, (int)z0, (int)i0, (int)i1)});
      }
    }
  std::stable_sort(units.begin(), units.end(), [](const std::pair<int64_t, int4> &a, const std::pair<int64_t, int4> &b) { return a.first > b.first; });
  std::vector<int4> uv(units.size());
  std::vector<int64_t> lapoff(units.size());
  int64_t lapn = 0;
  for (size_t k = 0; k < units.size(); ++k) {
    uv[k] = units[k].second;
    lapoff[k] = lapn;
    lapn += (int64_t)(uv[k].w - (uv[k].z > 0 ? uv[k].z - 1 : 0) + 1) * 3 * L3_NPT;   // warm-up interval included
  }
  p.nx = (int)nx; p.ny = (int)ny; p.nzp = (int)nzp; p.nunits = (int)uv.size(); p.lap_doubles = lapn;
#define UP(field, vec) do { NLS_CUDA(h, cudaMalloc((void **)&p.field, sizeof(vec[0]) * vec.size())); \
  NLS_CUDA(h, cudaMemcpyAsync(p.field, vec.data(), sizeof(vec[0]) * vec.size(), cudaMemcpyHostToDevice, h->stream)); } while (0)
  UP(d_pbase, pbase); UP(d_pown, pown); UP(d_info, info); UP(d_units, uv); UP(d_lapoff, lapoff);
#undef UP
  NLS_CUDA(h, cudaMalloc((void **)&p.d_carry, sizeof(double) * 18 * L3_THREADS * (size_t)h->sm_count));
  NLS_CUDA(h, cudaMalloc((void **)&p.d_counter, 2 * sizeof(unsigned)));
  NLS_CUDA(h, cudaMemsetAsync(p.d_counter, 0, 2 * sizeof(unsigned), h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  p.ok = true;
  p.tables = true;
  p.h_pown = pown; p.h_pbase = pbase;
  { const int rc = nls_mg_build(h); if (rc) return rc; }
  return nls_mg_prepare(h, false);
}

// 2-D meshes with grid topology (x-fastest rows; full grids and slab-local meshes): the same per-line addressing tables with
// one line per node row (ny = 1, the row index plays z). No assembly kernel uses them -- the 2-D default is the gather kernel --
// they give the multigrid preconditioner (kernels_mg.cu) the closed-form positions of the CSR blocks. p.ok stays false.
int nls_lines2_build(nls_context *h) {
  nls_lines3_free(h);
  Lines3Plan &p = h->lplan;
  if (h->dim != 2 || h->nen != 4 || h->nel == 0 || h->n_owned == 0) return NLS_OK;
  const int32_t *c = h->h_conn.data();
  const int64_t nel = h->nel, nn = h->nn;
  if (c[1] != c[0] + 1) return NLS_OK;
  // elements per row: the run of elements whose first node advances by one (the node above may sit in a ghost row with
  // an unrelated base id, so the row length cannot be read off one element)
  int64_t epr = 1;
  while (epr < nel && c[epr * 4] == c[0] + epr && c[epr * 4 + 1] == c[0] + epr + 1) ++epr;
  const int64_t nx = epr + 1;
  if (nx < 2 || nx > 1000000 || nel % epr) return NLS_OK;
  const int64_t nlay = nel / epr, nzp = nlay + 1;
  if (nn != nx * nzp || nzp > 1000000) return NLS_OK;
  std::vector<int32_t> pbase((size_t)nzp);
  for (int64_t k = 0; k < nlay; ++k) pbase[k] = c[k * epr * 4];
  pbase[nlay] = c[(nlay - 1) * epr * 4 + 2];
  for (int64_t k = 0; k < nzp; ++k) if (pbase[k] < 0 || (int64_t)pbase[k] + nx > nn) return NLS_OK;
  {
    std::vector<int32_t> s(pbase);
    std::sort(s.begin(), s.end());
    for (int64_t k = 1; k < nzp; ++k) if ((int64_t)s[k] - s[k - 1] < nx) return NLS_OK;
  }
  for (int64_t k = 0; k < nlay; ++k)
    for (int64_t i = 0; i < epr; ++i) {
      const int32_t *ce = c + (k * epr + i) * 4;
      const int64_t n0 = pbase[k] + i, n1 = pbase[k + 1] + i;
      if (ce[0] != n0 || ce[1] != n0 + 1 || ce[2] != n1 || ce[3] != n1 + 1) return NLS_OK;
    }
  std::vector<uint8_t> pown((size_t)nzp);
  for (int64_t k = 0; k < nzp; ++k) {
    pown[k] = pbase[k] < h->n_owned;
    if (pown[k] && (int64_t)pbase[k] + nx > h->n_owned) return NLS_OK;
  }
  std::vector<L3Line> info((size_t)nzp);
  const int64_t *bp = h->h_browptr.data();
  const int32_t *bc = h->h_bcol.data();
  for (int64_t z = 0; z < nzp; ++z) {
    L3Line &li = info[z];
    std::memset(&li, 0, sizeof(li));
    for (int k = 0; k < 9; ++k) li.rank[k] = -1;
    if (!pown[z]) continue;
    std::pair<int64_t, int> nb[3];
    int nl = 0;
    for (int dz = -1; dz <= 1; ++dz) {
      const int64_t z2 = z + dz;
      if (z2 < 0 || z2 >= nzp) continue;
      nb[nl++] = {(int64_t)pbase[z2], (dz + 1) * 3 + 1};
    }
    std::sort(nb, nb + nl);
    for (int k = 0; k < nl; ++k) li.rank[nb[k].second] = (int8_t)k;
    li.nl = nl;
    const int64_t first = pbase[z];
    li.lb = bp[first];
    for (int64_t i = 0; i < nx; ++i) {
      const int64_t a = first + i;
      const int w = (i == 0 || i == nx - 1) ? 2 : 3;
      const int64_t b0 = li.lb + (int64_t)nl * (i == 0 ? 0 : 3 * i - 1);
      if (bp[a] != b0 || bp[a + 1] - bp[a] != (int64_t)nl * w) return NLS_OK;
      for (int k = 0; k < nl; ++k)
        for (int dx = (i == 0 ? 0 : -1); dx <= (i == nx - 1 ? 0 : 1); ++dx)
          if (bc[b0 + k * w + dx + (i == 0 ? 0 : 1)] != nb[k].first + i + dx) return NLS_OK;
    }
  }
  p.nx = (int)nx; p.ny = 1; p.nzp = (int)nzp; p.nunits = 0;
  NLS_CUDA(h, cudaMalloc((void **)&p.d_pbase, sizeof(int32_t) * pbase.size()));
  NLS_CUDA(h, cudaMemcpyAsync(p.d_pbase, pbase.data(), sizeof(int32_t) * pbase.size(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaMalloc((void **)&p.d_info, sizeof(L3Line) * info.size()));
  NLS_CUDA(h, cudaMemcpyAsync(p.d_info, info.data(), sizeof(L3Line) * info.size(), cudaMemcpyHostToDevice, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  p.tables = true;
  p.h_pown = pown; p.h_pbase = pbase;
  { const int rc = nls_mg_build(h); if (rc) return rc; }
  return nls_mg_prepare(h, false);
}

template <bool R, bool K, bool LB, bool PF = false>
static void lines3_launch(nls_context *h, const L3Args &a, unsigned grid, size_t shm) {
  k_asm_q1_3d_lines<R, K, LB, PF><<<grid, L3_THREADS, shm, h->stream>>>(a, h->tabq, h->mat);
}

int nls_launch_lines3(nls_context *h, bool want_r, bool want_k) {
  Lines3Plan &p = h->lplan;
  if (!p.ok) NLS_FAIL(h, NLS_ERR_STATE, "lines assembly: no plan");
  if (!h->attr_lines3) {       // the kernel also holds a few bytes of static shared memory: ask for what it needs, not the opt-in maximum
    const int need = (int)lines3_smem(true);
    cudaError_t e1 = cudaFuncSetAttribute(k_asm_q1_3d_lines<true, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    cudaError_t e2 = cudaFuncSetAttribute(k_asm_q1_3d_lines<false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    cudaError_t e3 = cudaFuncSetAttribute(k_asm_q1_3d_lines<true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    cudaError_t e4 = cudaFuncSetAttribute(k_asm_q1_3d_lines<false, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    cudaError_t e5 = cudaFuncSetAttribute(k_asm_q1_3d_lines<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess) {
      cudaGetLastError();
      NLS_FAIL(h, NLS_ERR_CUDA, "lines assembly: cannot raise the dynamic shared-memory limit");
    }
    h->attr_lines3 = true;
  }
  L3Args a{};
  a.nx = p.nx; a.ny = p.ny; a.nzp = p.nzp; a.nunits = p.nunits;
  a.pbase = p.d_pbase; a.pown = p.d_pown; a.info = reinterpret_cast<const L3Line *>(p.d_info);
  a.units = reinterpret_cast<const int4 *>(p.d_units); a.lapoff = p.d_lapoff; a.lap = p.d_lap;
  a.coords = h->d_coords; a.U = h->d_U; a.R = h->d_R; a.vals = h->d_vals; a.counter = p.d_counter; a.carry = p.d_carry;
  a.dbg_nostore = h->opt_profile_phases == 2;
  a.prof = h->opt_profile_phases ? reinterpret_cast<unsigned long long *>(h->d_prof) : nullptr;
  const unsigned grid = (unsigned)std::min<int64_t>(p.nunits, h->sm_count);
  if (want_k && !p.lap_valid) {     // once per mesh: the reference-configuration Laplacian of every block this scheme stores
    if (!p.d_lap) {
      if (cudaMalloc((void **)&p.d_lap, sizeof(double) * (size_t)p.lap_doubles) != cudaSuccess) {
        cudaGetLastError(); p.d_lap = nullptr; p.ok = false;
        NLS_FAIL(h, NLS_ERR_CUDA, "lines assembly: no memory for the Laplacian table");
      }
      a.lap = p.d_lap;
    }
    lines3_launch<false, true, true>(h, a, grid, lines3_smem(false));
    h->launches++;
    p.lap_valid = true;
  }
  if (want_r && want_k && a.prof) lines3_launch<true, true, false, true>(h, a, grid, lines3_smem(true));   // instrumented build: tools/lines_probe.py
  else if (want_r && want_k) lines3_launch<true, true, false>(h, a, grid, lines3_smem(true));
  else if (want_k)      lines3_launch<false, true, false>(h, a, grid, lines3_smem(false));
  else                  lines3_launch<true, false, false>(h, a, grid, lines3_smem(true));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("lines kernel launch: ") + cudaGetErrorString(e); return NLS_ERR_CUDA; }
  h->launches++;
  return NLS_OK;
}
