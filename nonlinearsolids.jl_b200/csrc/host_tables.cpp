// host_tables.cpp -- host-side integer/tabulation work of libnls_b200.so:
//   * shape-function and quadrature tables, evaluated ONCE with the reference's own
//     formulas so the rounding of every tabulated value matches what the reference
//     recomputes per quadrature point (shapefunctions.jl:23-81, quadrature.jl:29-50);
//   * the node-level sparsity pattern (pattern precomputed on the host);
//   * the slab partitioner (nls_partition).
#include <algorithm>
#include <cmath>
#include <cstring>
#include "nls_internal.h"

// Julia's isapprox with default tolerances (rtol = sqrt(eps), atol = 0), used by `.≈`
// in shapefunctions.jl:51,66 to detect evaluation at a node.
static bool approx_eq(double a, double b) {
  if (a == b) return true;
  if (!std::isfinite(a) || !std::isfinite(b)) return false;
  return std::fabs(a - b) <= 1.4901161193847656e-8 * std::max(std::fabs(a), std::fabs(b));
}

// lagrange(order, t) node sets, shapefunctions.jl:38-47
void nls_host_lagrange_nodes(int p, int kind, double *nodes) {
  const int m = p + 1;
  for (int j = 0; j < m; ++j) {
    if (kind == 1) {  // :equidistant = LinRange(-1, 1, m)
      double t = m > 1 ? double(j) / double(m - 1) : 0.0;
      nodes[j] = (1.0 - t) * -1.0 + t * 1.0;
    } else {          // :chebyshev2 = cos(i*pi/order), i = order..0
      nodes[j] = std::cos(double(p - j) * M_PI / double(p));
    }
  }
}

// barycentric weights, shapefunctions.jl:23-30
void nls_host_lagrange_weights(int m, const double *x, double *w) {
  for (int i = 0; i < m; ++i) {
    double before = 1.0, after = 1.0;
    for (int k = 0; k < i; ++k) before *= x[i] - x[k];
    for (int k = i + 1; k < m; ++k) after *= x[i] - x[k];
    w[i] = 1.0 / (before * after);
  }
}

// N(shape, xi), shapefunctions.jl:50-61
void nls_host_shape_N(int m, const double *x, const double *w, double xi, double *N) {
  bool on_node = false;
  for (int i = 0; i < m; ++i) on_node = on_node || approx_eq(xi, x[i]);
  if (on_node) {
    for (int i = 0; i < m; ++i) N[i] = approx_eq(xi, x[i]) ? 1.0 : 0.0;
    return;
  }
  double ell = xi - x[0];
  for (int i = 1; i < m; ++i) ell *= xi - x[i];
  for (int i = 0; i < m; ++i) N[i] = ell * (w[i] / (xi - x[i]));
}

// gradN(shape, xi), shapefunctions.jl:64-81
void nls_host_shape_dN(int m, const double *x, const double *w, double xi, double *dN) {
  int k = -1;
  for (int i = 0; i < m && k < 0; ++i)
    if (approx_eq(xi, x[i])) k = i;
  if (k >= 0) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j) {
      dN[j] = (j == k) ? 0.0 : (w[j] / w[k]) / (x[k] - x[j]);
      acc += dN[j];
    }
    dN[k] = -acc;
    return;
  }
  double N[NLS_MAX_NEN1D];
  nls_host_shape_N(m, x, w, xi, N);
  for (int i = 0; i < m; ++i) {
    double recip_sum = 0.0;
    bool first = true;
    for (int j = 0; j < m; ++j) {
      if (j == i) continue;
      double t = 1.0 / (xi - x[j]);
      recip_sum = first ? t : recip_sum + t;
      first = false;
    }
    dN[i] = N[i] * recip_sum;
  }
}

// gauss_quadrature(order), quadrature.jl:29-50 -- the closed forms exactly as written there
int nls_host_gauss(int nq, double *x, double *w) {
  const double s3 = std::sqrt(3.0), s5 = std::sqrt(5.0), s6 = std::sqrt(6.0), s7 = std::sqrt(7.0);
  if (nq == 1) { x[0] = 0.0; w[0] = 2.0; }
  else if (nq == 2) { x[0] = -1.0 / s3; x[1] = 1.0 / s3; w[0] = w[1] = 1.0; }
  else if (nq == 3) {
    x[0] = -s3 / s5; x[1] = 0.0; x[2] = s3 / s5;
    w[0] = 5.0 / 9.0; w[1] = 8.0 / 9.0; w[2] = 5.0 / 9.0;
  } else if (nq == 4) {
    double a = std::sqrt(3.0 / 7.0 - 2.0 * s6 / (7.0 * s5)), b = std::sqrt(3.0 / 7.0 + 2.0 * s6 / (7.0 * s5));
    double wa = (18.0 + std::sqrt(30.0)) / 36.0, wb = (18.0 - std::sqrt(30.0)) / 36.0;
    x[0] = -b; x[1] = -a; x[2] = a; x[3] = b;
    w[0] = wb; w[1] = wa; w[2] = wa; w[3] = wb;
  } else if (nq == 5) {
    double a = 1.0 / 3.0 * std::sqrt(5.0 - 2.0 * std::sqrt(10.0) / s7);
    double b = 1.0 / 3.0 * std::sqrt(5.0 + 2.0 * std::sqrt(10.0) / s7);
    double wa = (322.0 + 13.0 * std::sqrt(70.0)) / 900.0, wb = (322.0 - 13.0 * std::sqrt(70.0)) / 900.0;
    x[0] = -b; x[1] = -a; x[2] = 0.0; x[3] = a; x[4] = b;
    w[0] = wb; w[1] = wa; w[2] = 128.0 / 225.0; w[3] = wa; w[4] = wb;
  } else return 1;
  return 0;
}

// FEMModel(mesh, q, p): Q = gauss_quadrature(q), P = lagrange(p, :chebyshev2) (fem.jl:18-23)
int nls_host_tab1d(int p, int nq, Tab1D *t) {
  if (p < 1 || p + 1 > NLS_MAX_NEN1D || nq < 1 || nq > NLS_MAX_NQ) return 1;
  double nodes[NLS_MAX_NEN1D], bw[NLS_MAX_NEN1D], xq[NLS_MAX_NQ];
  std::memset(t, 0, sizeof(*t));
  t->m = p + 1; t->nq = nq;
  nls_host_lagrange_nodes(p, 0, nodes);
  nls_host_lagrange_weights(p + 1, nodes, bw);
  nls_host_gauss(nq, xq, t->w);
  for (int q = 0; q < nq; ++q) {
    nls_host_shape_dN(p + 1, nodes, bw, xq[q], t->dN[q]);
    nls_host_shape_N(p + 1, nodes, bw, xq[q], t->N[q]);
  }
  return 0;
}

// Tensor-product Q1 tables from the same 1-D building blocks (p = 1, 2 Gauss points per
// direction, x fastest for both local nodes and quadrature points).
int nls_host_tabq1(int dim, TabQ1 *t) {
  if (dim != 2 && dim != 3) return 1;
  double nodes[2], bw[2], xq[2], wq[2], N1[2][2], D1[2][2];
  std::memset(t, 0, sizeof(*t));
  nls_host_lagrange_nodes(1, 0, nodes);
  nls_host_lagrange_weights(2, nodes, bw);
  nls_host_gauss(2, xq, wq);
  for (int q = 0; q < 2; ++q) {
    nls_host_shape_N(2, nodes, bw, xq[q], N1[q]);
    nls_host_shape_dN(2, nodes, bw, xq[q], D1[q]);
  }
  t->dim = dim; t->nen = 1 << dim; t->nqp = 1 << dim;
  for (int q = 0; q < t->nqp; ++q) {
    int qd[3] = {q & 1, (q >> 1) & 1, (q >> 2) & 1};
    double w = wq[qd[0]];
    for (int d = 1; d < dim; ++d) w *= wq[qd[d]];
    t->w[q] = w;
    for (int a = 0; a < t->nen; ++a) {
      int ad[3] = {a & 1, (a >> 1) & 1, (a >> 2) & 1};
      for (int D = 0; D < dim; ++D) {
        double v = (D == 0) ? D1[qd[0]][ad[0]] : N1[qd[0]][ad[0]];
        for (int d = 1; d < dim; ++d) v *= (d == D) ? D1[qd[d]][ad[d]] : N1[qd[d]][ad[d]];
        t->dN[q][a][D] = v;
      }
      double nv = N1[qd[0]][ad[0]];
      for (int d = 1; d < dim; ++d) nv *= N1[qd[d]][ad[d]];
      t->N[q][a] = nv;
    }
  }
  return 0;
}

// Node-level pattern: block row a (owned nodes only) lists, ascending, every node that shares
// an element with a. The scalar CSR pattern is its dim x dim expansion, which is exactly
// {(i,j): exists e with i,j in dofs(e)} -- the entries assemble! can touch (element.jl:70).
void nls_host_block_pattern(int nen, int64_t nel, const int32_t *conn, int64_t nn, int64_t n_owned,
                            std::vector<int64_t> &browptr, std::vector<int32_t> &bcol) {
  // node -> incident elements (counting sort)
  std::vector<int64_t> nptr(nn + 1, 0);
  for (int64_t k = 0; k < nel * nen; ++k) nptr[conn[k] + 1]++;
  for (int64_t i = 0; i < nn; ++i) nptr[i + 1] += nptr[i];
  std::vector<int64_t> nidx(nptr[nn]);
  {
    std::vector<int64_t> cur(nptr.begin(), nptr.end() - 1);
    for (int64_t e = 0; e < nel; ++e)
      for (int a = 0; a < nen; ++a) nidx[cur[conn[e * nen + a]]++] = e;
  }
  browptr.assign(n_owned + 1, 0);
  bcol.clear();
  bcol.reserve((size_t)n_owned * (nen == 8 ? 27 : (nen == 4 ? 9 : 2 * nen - 1)));
  std::vector<int32_t> scratch;
  for (int64_t a = 0; a < n_owned; ++a) {
    scratch.clear();
    for (int64_t k = nptr[a]; k < nptr[a + 1]; ++k) {
      const int32_t *c = conn + nidx[k] * nen;
      scratch.insert(scratch.end(), c, c + nen);
    }
    std::sort(scratch.begin(), scratch.end());
    scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
    bcol.insert(bcol.end(), scratch.begin(), scratch.end());
    browptr[a + 1] = (int64_t)bcol.size();
  }
}

// ---- slab partitioner ------------------------------------------------------------------
namespace {
struct RankView {
  std::vector<int64_t> elems, ghosts;
};
RankView rank_view(int nen, int64_t nel, const int32_t *conn, int64_t lo, int64_t hi, bool want_elems) {
  RankView v;
  for (int64_t e = 0; e < nel; ++e) {
    const int32_t *c = conn + e * nen;
    bool touches = false;
    for (int a = 0; a < nen; ++a) touches = touches || (c[a] >= lo && c[a] < hi);
    if (!touches) continue;
    if (want_elems) v.elems.push_back(e);
    for (int a = 0; a < nen; ++a)
      if (c[a] < lo || c[a] >= hi) v.ghosts.push_back(c[a]);
  }
  std::sort(v.ghosts.begin(), v.ghosts.end());
  v.ghosts.erase(std::unique(v.ghosts.begin(), v.ghosts.end()), v.ghosts.end());
  return v;
}
}  // namespace

extern "C" int32_t nls_partition(int32_t nen, int64_t nelem, const int32_t *conn0, int32_t nranks,
                                 const int64_t *node_off, int32_t rank, int64_t *n_loc_elems,
                                 int64_t *loc_elems, int64_t *n_ghost, int64_t *ghost_nodes,
                                 int64_t *send_ptr, int64_t *send_nodes) {
  if (!conn0 || !node_off || !n_loc_elems || !n_ghost || !send_ptr || rank < 0 || rank >= nranks || nen < 1)
    return NLS_ERR_ARG;
  RankView mine = rank_view(nen, nelem, conn0, node_off[rank], node_off[rank + 1], true);
  *n_loc_elems = (int64_t)mine.elems.size();
  *n_ghost = (int64_t)mine.ghosts.size();
  if (loc_elems) std::copy(mine.elems.begin(), mine.elems.end(), loc_elems);
  if (ghost_nodes) std::copy(mine.ghosts.begin(), mine.ghosts.end(), ghost_nodes);
  int64_t pos = 0;
  send_ptr[0] = 0;
  for (int q = 0; q < nranks; ++q) {
    if (q != rank) {
      RankView other = rank_view(nen, nelem, conn0, node_off[q], node_off[q + 1], false);
      for (int64_t g : other.ghosts)
        if (g >= node_off[rank] && g < node_off[rank + 1]) {
          if (send_nodes) send_nodes[pos] = g;
          ++pos;
        }
    }
    send_ptr[q + 1] = pos;
  }
  return NLS_OK;
}

// ---- host-only building blocks exported through the C ABI ---------------------------------
extern "C" int32_t nls_lagrange(int32_t p_order, int32_t kind, double *nodes, double *weights) {
  if (p_order < 1 || p_order + 1 > NLS_MAX_NEN1D || !nodes || !weights || (kind != 0 && kind != 1)) return NLS_ERR_ARG;
  nls_host_lagrange_nodes(p_order, kind, nodes);
  nls_host_lagrange_weights(p_order + 1, nodes, weights);
  return NLS_OK;
}
extern "C" int32_t nls_shape_eval(int32_t m, const double *nodes, const double *weights, double xi, double *N, double *dN) {
  if (m < 2 || m > NLS_MAX_NEN1D || !nodes || !weights) return NLS_ERR_ARG;
  if (N) nls_host_shape_N(m, nodes, weights, xi, N);
  if (dN) nls_host_shape_dN(m, nodes, weights, xi, dN);
  return NLS_OK;
}
extern "C" int32_t nls_gauss_quadrature(int32_t nq, double *points, double *weights) {
  if (!points || !weights) return NLS_ERR_ARG;
  return nls_host_gauss(nq, points, weights) ? NLS_ERR_ARG : NLS_OK;
}
extern "C" int32_t nls_pattern_host(int32_t dim, int32_t nen, int64_t nelem, const int32_t *conn0, int64_t nnodes,
                                    int64_t n_owned_nodes, int64_t *rowptr, int32_t *colind, int64_t *nnz) {
  if (dim < 1 || dim > 3 || nen < 1 || nelem < 0 || nnodes < 0 || n_owned_nodes < 0 || n_owned_nodes > nnodes || !rowptr || !nnz ||
      (nelem > 0 && !conn0)) return NLS_ERR_ARG;
  for (int64_t k = 0; k < nelem * nen; ++k) if (conn0[k] < 0 || conn0[k] >= nnodes) return NLS_ERR_ARG;
  std::vector<int64_t> browptr; std::vector<int32_t> bcol;
  nls_host_block_pattern(nen, nelem, conn0, nnodes, n_owned_nodes, browptr, bcol);
  int64_t pos = 0;
  for (int64_t a = 0; a < n_owned_nodes; ++a) {
    const int64_t b0 = browptr[a], deg = browptr[a + 1] - b0;
    for (int i = 0; i < dim; ++i) {
      rowptr[a * dim + i] = pos;
      if (colind)
        for (int64_t j = 0; j < deg; ++j)
          for (int k = 0; k < dim; ++k) colind[pos + j * dim + k] = bcol[b0 + j] * dim + k;
      pos += deg * dim;
    }
  }
  rowptr[n_owned_nodes * dim] = pos;
  *nnz = pos;
  return NLS_OK;
}

// ---- cluster plan for the gather (row-owner) assembly ----------------------------------------
// Owned nodes are grouped into spatially compact clusters of <= nc nodes (Morton order of the
// quantised coordinates, or rectangular tiles of bins when the caller asks for them; works for any mesh). A CTA
// computes every element touching its cluster once into shared memory, then each cluster node
// gathers its rows from there: no atomics, no zero-fill, rows written once, contributions summed
// in ascending element order (the order of the reference's serial element loop, defaults.jl:77-81).
static inline uint64_t spread_bits(uint64_t v, int dim) {
  uint64_t r = 0;
  for (int b = 0; b < 21; ++b) r |= ((v >> b) & 1ull) << (b * dim);
  return r;
}

void nls_host_cluster_plan(int dim, int nen, int64_t nel, const int32_t *conn, int64_t nn, int64_t n_owned,
                           const double *coords, int nc, AsmPlan &plan, const int *tile) {
  plan = AsmPlan();
  plan.nc = nc;
  if (n_owned == 0) { plan.eptr.assign(1, 0); return; }
  // Keys of the owned nodes. Coordinates are binned on a grid with about one node per bin (bins per
  // axis = nodes^(1/dim) scaled by the aspect ratio), so on a structured mesh the bin index is the grid
  // index. Without `tile`, keys are Morton codes and a tile (key >> tile_bits) is an 8x8 (4x4x4) node
  // tile for nc = 64. With `tile` = {tx, ty, tz} the tiles are tx x ty (x tz) bins, ordered
  // lexicographically: the 2-D kernel asks for 15 x 7 nodes, whose 16 x 8 = 128 touching elements fill
  // its 128 threads exactly.
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t a = 0; a < n_owned; ++a)
    for (int d = 0; d < dim; ++d) { lo[d] = std::min(lo[d], coords[a * dim + d]); hi[d] = std::max(hi[d], coords[a * dim + d]); }
  double vol = 1.0;
  for (int d = 0; d < dim; ++d) vol *= std::max(hi[d] - lo[d], 1e-300);
  const double hbin = std::pow(vol / (double)n_owned, 1.0 / dim);   // ~ node spacing
  int tile_bits = 0;
  while ((1 << tile_bits) < nc) ++tile_bits;                        // nc = 64 -> 6 bits: 8x8 or 4x4x4
  uint64_t ntile[3] = {1, 1, 1};
  if (tile) {
    tile_bits = 24;                                                 // key = tile id << 24 | position inside the tile
    for (int d = 0; d < dim; ++d) ntile[d] = (uint64_t)std::floor((hi[d] - lo[d]) / hbin + 0.5) / (uint64_t)tile[d] + 1;
  }
  std::vector<std::pair<uint64_t, int32_t>> keyed((size_t)n_owned);
  for (int64_t a = 0; a < n_owned; ++a) {
    uint64_t key = 0, tid_ = 0, within = 0, tmul = 1, wmul = 1;
    for (int d = 0; d < dim; ++d) {
      double b = std::floor((coords[a * dim + d] - lo[d]) / hbin + 0.5);
      if (b < 0) b = 0;
      if (b > 2097151.0) b = 2097151.0;
      if (tile) {
        const uint64_t bi = (uint64_t)b, t = bi / (uint64_t)tile[d];
        tid_ += std::min<uint64_t>(t, ntile[d] - 1) * tmul; tmul *= ntile[d];
        within += (bi % (uint64_t)tile[d]) * wmul; wmul *= (uint64_t)tile[d];
      } else {
        key |= spread_bits((uint64_t)b, dim) << d;
      }
    }
    if (tile) key = (tid_ << 24) | std::min<uint64_t>(within, (1u << 24) - 1);
    keyed[a] = {key, (int32_t)a};
  }
  std::sort(keyed.begin(), keyed.end());
  // clusters: split the Morton order at tile boundaries, and every nc nodes inside a tile
  std::vector<int64_t> cstart;
  for (int64_t s = 0, in_tile = 0; s < n_owned; ++s) {
    const bool new_tile = (s == 0) || ((keyed[s].first >> tile_bits) != (keyed[s - 1].first >> tile_bits));
    if (new_tile || in_tile == nc) { cstart.push_back(s); in_tile = 0; }
    ++in_tile;
  }
  cstart.push_back(n_owned);
  // node -> incident elements
  std::vector<int64_t> nptr(nn + 1, 0);
  for (int64_t k = 0; k < nel * nen; ++k) nptr[conn[k] + 1]++;
  for (int64_t i = 0; i < nn; ++i) nptr[i + 1] += nptr[i];
  std::vector<int32_t> nidx(nptr[nn]);
  {
    std::vector<int64_t> cur(nptr.begin(), nptr.end() - 1);
    for (int64_t e = 0; e < nel; ++e)
      for (int a = 0; a < nen; ++a) nidx[cur[conn[e * nen + a]]++] = (int32_t)e;   // ascending e per node
  }
  const int64_t ncl = (int64_t)cstart.size() - 1;
  plan.ncl = ncl;
  plan.nodes.assign((size_t)(ncl * nc), -1);
  plan.eptr.assign((size_t)ncl + 1, 0);
  std::vector<int32_t> el;
  for (int64_t c = 0; c < ncl; ++c) {
    el.clear();
    const int64_t n0 = cstart[c], n1 = cstart[c + 1];
    for (int64_t s = n0; s < n1; ++s) {
      const int32_t a = keyed[s].second;
      plan.nodes[c * nc + (s - n0)] = a;
      plan.max_adj = std::max<int>(plan.max_adj, (int)(nptr[a + 1] - nptr[a]));
      el.insert(el.end(), nidx.begin() + nptr[a], nidx.begin() + nptr[a + 1]);
    }
    std::sort(el.begin(), el.end());
    el.erase(std::unique(el.begin(), el.end()), el.end());
    plan.max_elems = std::max<int>(plan.max_elems, (int)el.size());
    plan.elems.insert(plan.elems.end(), el.begin(), el.end());
    plan.eptr[c + 1] = (int32_t)plan.elems.size();
  }
}

extern "C" int32_t nls_plan_stats(int32_t dim, int32_t nen, int64_t nelem, const int32_t *conn0, int64_t nnodes,
                                  int64_t n_owned_nodes, const double *coords, int32_t nc, int64_t *out4) {
  if (dim < 1 || dim > 3 || nen < 1 || nelem < 0 || nnodes < 1 || n_owned_nodes < 0 || n_owned_nodes > nnodes || !coords ||
      !out4 || nc < 1 || (nelem > 0 && !conn0)) return NLS_ERR_ARG;
  AsmPlan plan;
  const int tile2d[3] = {15, 7, 1};
  nls_host_cluster_plan(dim, nen, nelem, conn0, nnodes, n_owned_nodes, coords, nc, plan, (dim == 2 && nc == 105) ? tile2d : nullptr);
  out4[0] = plan.ncl; out4[1] = plan.max_elems; out4[2] = (int64_t)plan.elems.size(); out4[3] = plan.max_adj;
  return NLS_OK;
}
