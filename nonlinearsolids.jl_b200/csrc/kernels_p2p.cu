// kernels_p2p.cu -- halo exchange and Krylov reductions over NVLink peer memory (one process per GPU).
//
// The slab partition needs, per operator application, one forward exchange of the ghost DoFs
// (<= 2 neighbours, one node plane each) and, per PCG iteration, two sums of 1-2 doubles over all ranks.
// Both are latency-bound: through ncclSend/ncclRecv and ncclAllReduce they cost 25-45 us each, more than the
// SpMV of a strong-scaled slab. Here every rank exports a small window and a flag array with CUDA IPC (the
// handles travel once over the NCCL communicator); after that the data path is plain kernels:
//   halo      k_p2p_halo: packs the send planes and STORES them straight into the neighbours' windows (posted
//             NVLink writes), fences, the last block raises an epoch flag in each neighbour's memory; then every
//             block waits for the neighbours' flags (ld.acquire.sys) and unpacks its own window into the ghost
//             entries. One launch per exchange.
//   allreduce k_p2p_allreduce: every rank writes its partial sums into all windows, raises flags, waits for all
//             flags, and adds the nranks slots in rank order -- the same bits on every rank, run to run.
// Windows are double buffered by epoch parity: a rank can only be one exchange ahead of a neighbour (its own
// receive waits for the neighbour's send of the same epoch, which follows the neighbour's previous receive in
// stream order), so the slot being overwritten has always been consumed.
// Spins are bounded (about 60 s): a lost peer raises an error flag instead of hanging the GPU.
#include <cstdlib>
#include <cstring>
#include "nls_internal.h"
#include "nccl_dyn.h"

#define P2P_SPIN_CYCLES (120000000000ll)   // ~60 s at 2 GHz: ranks may reach an exchange seconds apart (host-side setup)

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// `err` lives in device memory and is sticky: once set, every later exchange returns at once instead of spinning its
// full minute. `herr` is its host-visible twin in mapped pinned memory, which the solvers read wherever they already
// synchronise with the stream -- no extra copy.
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long epoch, int *err, volatile int *herr) {
  if (*reinterpret_cast<volatile int *>(err)) return false;
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < epoch) {
    if (clock64() - t0 > P2P_SPIN_CYCLES || *reinterpret_cast<volatile int *>(err)) {
      atomicExch(err, 1); *herr = 1; __threadfence_system();
      return false;
    }
    __nanosleep(40);
  }
  return true;
}

struct P2PHaloArgs {
  int nneigh, dim;
  int64_t send_ptr[9];             // nodes, per neighbour (NLS_P2P_MAXNEIGH = 8)
  double *dst[8];                  // neighbour q's window, at the slot my data goes to (this epoch's parity)
  unsigned long long *dst_flag[8]; // neighbour q's halo flag for me
  const unsigned long long *my_flag[8];   // my halo flag for neighbour q
};

// One launch per exchange: pack + push + flag, then wait for the neighbours' flags and unpack. All blocks are
// co-resident (grid <= 64), so waiting inside the kernel cannot starve the sending part of another block.
__global__ void __launch_bounds__(256)
k_p2p_halo(const P2PHaloArgs a, const int32_t *__restrict__ send_nodes, int64_t nrecv, const int32_t *__restrict__ recv_nodes,
           const double *win, double *v, unsigned long long epoch, unsigned int *done, int *err, volatile int *herr) {
  const int64_t n = a.send_ptr[a.nneigh] * a.dim;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node_idx = t / a.dim;
    int q = 0;
    while (q + 1 < a.nneigh && node_idx >= a.send_ptr[q + 1]) ++q;
    a.dst[q][t - a.send_ptr[q] * a.dim] = v[(int64_t)send_nodes[node_idx] * a.dim + (t - node_idx * a.dim)];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {           // last block: every block's stores are fenced -> raise the flags
      *done = 0;
      __threadfence_system();
      for (int q = 0; q < a.nneigh; ++q) st_release_sys(a.dst_flag[q], epoch);
    }
    ok = 1;
    for (int q = 0; q < a.nneigh; ++q) if (!spin_until(a.my_flag[q], epoch, err, herr)) ok = 0;
  }
  __syncthreads();
  if (!ok) return;
  const int64_t m = nrecv * a.dim;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < m; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node_idx = t / a.dim;
    v[(int64_t)recv_nodes[node_idx] * a.dim + (t - node_idx * a.dim)] = __ldcv(win + t);
  }
}

// one CTA; thread r talks to rank r
__global__ void __launch_bounds__(64)
k_p2p_allreduce(double *dptr, int count, int me, int nranks, double *const *peer_red, unsigned long long *const *peer_flag,
                const double *my_red, const unsigned long long *my_flag, unsigned long long epoch, int *err, volatile int *herr) {
  __shared__ int ok;
  const int r = threadIdx.x;
  if (r == 0) ok = 1;
  __syncthreads();
  if (r < nranks) {
    double *dst = peer_red[r] + (size_t)me * 4;
    for (int c = 0; c < count; ++c) dst[c] = dptr[c];
    __threadfence_system();
    st_release_sys(peer_flag[r] + me, epoch);
    if (!spin_until(my_flag + r, epoch, err, herr)) ok = 0;
  }
  __syncthreads();
  if (r == 0 && ok) {
    for (int c = 0; c < count; ++c) {
      double s = 0.0;
      for (int q = 0; q < nranks; ++q) s += __ldcv(my_red + (size_t)q * 4 + c);
      dptr[c] = s;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------
void nls_p2p_free(nls_context *h) {
  P2PState &p = h->p2p;
  for (void *q : p.opened) cudaIpcCloseMemHandle(q);
  if (p.win) cudaFree(p.win);
  if (p.flags) cudaFree(p.flags);
  if (p.d_peer_red[0]) cudaFree(p.d_peer_red[0]);
  if (p.d_peer_red[1]) cudaFree(p.d_peer_red[1]);
  if (p.d_peer_rflag) cudaFree(p.d_peer_rflag);
  if (p.d_done) cudaFree(p.d_done);
  if (p.d_err) cudaFree(p.d_err);
  if (p.h_err) cudaFreeHost((void *)p.h_err);
  p = P2PState();
}

// Collective over the communicator (every rank calls it from nls_set_halo). Leaves p2p.on false -- the NCCL
// path stays in use -- when IPC is unavailable on ANY rank or NLS_B200_P2P=0.
int nls_p2p_setup(nls_context *h) {
  nls_p2p_free(h);
  // nranks is the same on every rank, so this early return is taken by all of them or by none. Per-rank conditions
  // (too many neighbours, a bad neighbour list) must NOT return here: the exchange below is collective, so such a rank
  // publishes ok = 0 and all ranks fall back to NCCL together.
  if (!h->distributed || !h->comm || h->nranks < 2 || h->nranks > 64) return NLS_OK;
  const char *env = std::getenv("NLS_B200_P2P");
  bool neigh_ok = h->nneigh <= 8;
  for (int q = 0; q < h->nneigh && neigh_ok; ++q) {
    if (h->neigh[q] < 0 || h->neigh[q] >= h->nranks || h->neigh[q] == h->rank) neigh_ok = false;
    for (int q2 = 0; q2 < q && neigh_ok; ++q2) if (h->neigh[q2] == h->neigh[q]) neigh_ok = false;
  }
  const bool bad_ranks = h->nneigh <= 8 && !neigh_ok;     // an argument error, reported after the collective part
  const bool want = !(env && env[0] == '0') && neigh_ok;
  P2PState &p = h->p2p;
  NcclApi *nc = h->nccl;
  const int nr = h->nranks, me = h->rank, dim = h->dim;
  const int64_t H = h->recv_ptr[h->nneigh] * dim;                      // doubles of one halo area
  const size_t win_doubles = (size_t)2 * H + (size_t)2 * nr * 4;
  const size_t nflags = (size_t)2 * nr;                                // [0][r] halo, [1][r] reduction
  NLS_CUDA(h, cudaMalloc((void **)&p.win, std::max<size_t>(1, win_doubles) * sizeof(double)));
  NLS_CUDA(h, cudaMalloc((void **)&p.flags, nflags * sizeof(unsigned long long)));
  NLS_CUDA(h, cudaMemset(p.win, 0, std::max<size_t>(1, win_doubles) * sizeof(double)));
  NLS_CUDA(h, cudaMemset(p.flags, 0, nflags * sizeof(unsigned long long)));
  NLS_CUDA(h, cudaMalloc((void **)&p.d_done, sizeof(unsigned int)));
  NLS_CUDA(h, cudaMemset(p.d_done, 0, sizeof(unsigned int)));
  NLS_CUDA(h, cudaMalloc((void **)&p.d_err, sizeof(int)));
  NLS_CUDA(h, cudaMemset(p.d_err, 0, sizeof(int)));
  NLS_CUDA(h, cudaHostAlloc((void **)&p.h_err, sizeof(int), cudaHostAllocMapped));
  *p.h_err = 0;
  NLS_CUDA(h, cudaHostGetDevicePointer((void **)&p.d_herr, (void *)p.h_err, 0));
  // record: {ok, H, recv offset (doubles) for data from rank s or -1}[nr], win handle, flag handle
  const size_t rec_i64 = (size_t)2 + nr, rec_bytes = rec_i64 * 8 + 2 * sizeof(cudaIpcMemHandle_t);
  std::vector<unsigned char> mine(rec_bytes, 0), all(rec_bytes * nr, 0);
  int64_t *mi = reinterpret_cast<int64_t *>(mine.data());
  cudaIpcMemHandle_t hw, hf;
  bool ok = want && cudaIpcGetMemHandle(&hw, p.win) == cudaSuccess && cudaIpcGetMemHandle(&hf, p.flags) == cudaSuccess;
  cudaGetLastError();
  mi[0] = ok ? 1 : 0; mi[1] = H;
  for (int s = 0; s < nr; ++s) mi[2 + s] = -1;
  for (int q = 0; q < h->nneigh && neigh_ok; ++q) mi[2 + h->neigh[q]] = h->recv_ptr[q] * dim;
  std::memcpy(mine.data() + rec_i64 * 8, &hw, sizeof(hw));
  std::memcpy(mine.data() + rec_i64 * 8 + sizeof(hw), &hf, sizeof(hf));
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  NLS_CUDA(h, cudaMalloc((void **)&d_send, rec_bytes));
  NLS_CUDA(h, cudaMalloc((void **)&d_recv, rec_bytes * nr));
  NLS_CUDA(h, cudaMemcpyAsync(d_send, mine.data(), rec_bytes, cudaMemcpyHostToDevice, h->stream));
  if (nc->AllGather(d_send, d_recv, rec_bytes, 1 /* ncclUint8 */, h->comm, h->stream) != 0) {
    cudaFree(d_send); cudaFree(d_recv);
    NLS_FAIL(h, NLS_ERR_NCCL, "ncclAllGather of the IPC handles failed");
  }
  NLS_CUDA(h, cudaMemcpyAsync(all.data(), d_recv, rec_bytes * nr, cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_send); cudaFree(d_recv);
  std::vector<double *> pwin(nr, nullptr);
  std::vector<unsigned long long *> pflag(nr, nullptr);
  std::vector<int64_t> pH(nr, 0);
  for (int r = 0; r < nr && ok; ++r) {
    const unsigned char *rec = all.data() + rec_bytes * r;
    const int64_t *ri = reinterpret_cast<const int64_t *>(rec);
    if (!ri[0]) { ok = false; break; }
    pH[r] = ri[1];
    if (r == me) { pwin[r] = p.win; pflag[r] = p.flags; continue; }
    cudaIpcMemHandle_t a, b;
    std::memcpy(&a, rec + rec_i64 * 8, sizeof(a));
    std::memcpy(&b, rec + rec_i64 * 8 + sizeof(a), sizeof(b));
    void *pa = nullptr, *pb = nullptr;
    if (cudaIpcOpenMemHandle(&pa, a, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
    p.opened.push_back(pa);
    if (cudaIpcOpenMemHandle(&pb, b, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
    p.opened.push_back(pb);
    pwin[r] = (double *)pa; pflag[r] = (unsigned long long *)pb;
  }
  // every rank must take the same path: agree on the minimum
  double *d_ok = h->d_scal;
  double notok = ok ? 0.0 : 1.0;             // sum over ranks is zero iff every rank opened every window
  NLS_CUDA(h, cudaMemcpy(d_ok, &notok, sizeof(double), cudaMemcpyHostToDevice));
  if (nc->AllReduce(d_ok, d_ok, 1, NCCL_DYN_FLOAT64, NCCL_DYN_SUM, h->comm, h->stream) != 0)
    NLS_FAIL(h, NLS_ERR_NCCL, "ncclAllReduce (p2p agreement) failed");
  NLS_CUDA(h, cudaMemcpyAsync(&notok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  if (bad_ranks) { nls_p2p_free(h); NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: neighbour rank out of range, equal to this rank, or listed twice"); }
  if (notok != 0.0) { nls_p2p_free(h); return NLS_OK; }
  // per-neighbour destinations (both parities) and flags
  p.H = H;
  for (int q = 0; q < h->nneigh; ++q) {
    const int r = h->neigh[q];
    const int64_t *ri = reinterpret_cast<const int64_t *>(all.data() + rec_bytes * r);
    const int64_t off = ri[2 + me];
    if (off < 0) { nls_p2p_free(h); NLS_FAIL(h, NLS_ERR_ARG, "nls_set_halo: neighbour lists are not symmetric across ranks"); }
    p.dst[0][q] = pwin[r] + off;
    p.dst[1][q] = pwin[r] + pH[r] + off;
    p.dst_flag[q] = pflag[r] + me;            // halo flags: [0][src]
    p.my_flag[q] = p.flags + r;
  }
  // reduction areas: after the two halo areas of each window
  std::vector<double *> red0(nr), red1(nr);
  std::vector<unsigned long long *> rfl(nr);
  for (int r = 0; r < nr; ++r) {
    red0[r] = pwin[r] + 2 * pH[r];
    red1[r] = red0[r] + (size_t)nr * 4;
    rfl[r] = pflag[r] + nr;                   // reduction flags: [1][src]
  }
  NLS_CUDA(h, cudaMalloc((void **)&p.d_peer_red[0], sizeof(double *) * nr));
  NLS_CUDA(h, cudaMalloc((void **)&p.d_peer_red[1], sizeof(double *) * nr));
  NLS_CUDA(h, cudaMalloc((void **)&p.d_peer_rflag, sizeof(unsigned long long *) * nr));
  NLS_CUDA(h, cudaMemcpy(p.d_peer_red[0], red0.data(), sizeof(double *) * nr, cudaMemcpyHostToDevice));
  NLS_CUDA(h, cudaMemcpy(p.d_peer_red[1], red1.data(), sizeof(double *) * nr, cudaMemcpyHostToDevice));
  NLS_CUDA(h, cudaMemcpy(p.d_peer_rflag, rfl.data(), sizeof(unsigned long long *) * nr, cudaMemcpyHostToDevice));
  p.my_red[0] = p.win + 2 * H;
  p.my_red[1] = p.my_red[0] + (size_t)nr * 4;
  p.on = true;
  return NLS_OK;
}

#define P2P_KCHECK(h) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  (h)->err = std::string("kernel launch: ") + cudaGetErrorString(e_); return NLS_ERR_CUDA; } (h)->launches++; } while (0)

int nls_p2p_halo_exchange(nls_context *h, double *v) {
  P2PState &p = h->p2p;
  const int dim = h->dim;
  const unsigned long long epoch = ++p.halo_epoch;
  const int par = (int)(epoch & 1);
  P2PHaloArgs a{};
  a.nneigh = h->nneigh; a.dim = dim;
  for (int q = 0; q <= h->nneigh; ++q) a.send_ptr[q] = h->send_ptr[q];
  for (int q = 0; q < h->nneigh; ++q) { a.dst[q] = p.dst[par][q]; a.dst_flag[q] = p.dst_flag[q]; a.my_flag[q] = p.my_flag[q]; }
  const int64_t ns = h->send_ptr[h->nneigh] * dim, nr = h->recv_ptr[h->nneigh] * dim;
  const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((std::max(ns, nr) + 255) / 256, 64));
  k_p2p_halo<<<g, 256, 0, h->halo_stream ? h->halo_stream : h->stream>>>(a, h->d_send_nodes, h->recv_ptr[h->nneigh], h->d_recv_nodes, p.win + (size_t)par * p.H, v,
                                       epoch, p.d_done, p.d_err, p.d_herr);
  P2P_KCHECK(h);
  return NLS_OK;
}

int nls_p2p_allreduce(nls_context *h, double *dptr, int count) {
  P2PState &p = h->p2p;
  if (count > 4) NLS_FAIL(h, NLS_ERR_ARG, "p2p allreduce holds at most 4 doubles");
  const unsigned long long epoch = ++p.red_epoch;
  const int par = (int)(epoch & 1);
  k_p2p_allreduce<<<1, 64, 0, h->stream>>>(dptr, count, h->rank, h->nranks, p.d_peer_red[par], p.d_peer_rflag, p.my_red[par],
                                           p.flags + h->nranks, epoch, p.d_err, p.d_herr);
  P2P_KCHECK(h);
  return NLS_OK;
}

// After a stream synchronisation: did any exchange so far time out? Reads mapped pinned memory, no copy. Sticky.
int nls_p2p_failed(nls_context *h) {
  if (!h->p2p.on || !h->p2p.h_err) return NLS_OK;
  if (*h->p2p.h_err) NLS_FAIL(h, NLS_ERR_NCCL, "peer-memory exchange timed out waiting for a neighbour's flag (results since then are invalid)");
  return NLS_OK;
}

int nls_p2p_check(nls_context *h) {
  if (!h->p2p.on) return NLS_OK;
  NLS_CUDA(h, cudaStreamSynchronize(h->stream));
  return nls_p2p_failed(h);
}
