// ubench.cu -- B200 micro-benchmarks that size the design choices of the assembly kernels:
// HBM copy/store bandwidth, FP64 DFMA peak, and FP64 RED (atomicAdd) vs plain RMW scatter
// throughput for the access shapes a CSR scatter produces. Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_store(double2 *__restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_double2(1.0, 2.0);
}
__global__ void k_dfma(double *out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// scatter of `run` consecutive doubles per group; groups placed pseudo-randomly (stride hash) or contiguously.
// mode 0: RED (atomicAdd), 1: plain RMW, 2: plain store
template <int MODE>
__global__ void k_scatter(double *a, size_t n_groups, int run, size_t total_ops, int contiguous, size_t mask) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total_ops; t += (size_t)gridDim.x * blockDim.x) {
    size_t g = t / run; int r = (int)(t % run);
    size_t gg = contiguous ? g : ((g * 2654435761ull) & mask);
    if (!contiguous) gg = ((g * 2654435761ull) & mask) % n_groups;
    double *p = a + gg * run + r;
    if (MODE == 0) atomicAdd(p, 1.0);
    else if (MODE == 1) *p += 1.0;
    else *p = 1.0;
  }
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  const size_t N = (size_t)1 << 28;  // 2 GiB of doubles
  double *a, *b; CK(cudaMalloc(&a, N * 8)); CK(cudaMalloc(&b, N * 8));
  CK(cudaMemset(a, 0, N * 8)); CK(cudaMemset(b, 0, N * 8));
  float ms = timeit([&] { k_copy<<<sms * 16, 512>>>((double2 *)a, (double2 *)b, N / 2); });
  printf(", \"copy_GBs\": %.1f", 2.0 * N * 8 / ms / 1e6);
  ms = timeit([&] { k_store<<<sms * 16, 512>>>((double2 *)b, N / 2); });
  printf(", \"store_GBs\": %.1f", 1.0 * N * 8 / ms / 1e6);
  ms = timeit([&] { cudaMemsetAsync(b, 0, N * 8); });
  printf(", \"memset_GBs\": %.1f", 1.0 * N * 8 / ms / 1e6);
  {
    const int iters = 20000, blocks = sms * 8, threads = 256;
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(b, iters); }, 3);
    printf(", \"dfma_TFLOPs\": %.2f", 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9);
  }
  // scatter shapes: run in {1,2,3,6,9,18,32}; footprint large (2 GiB, DRAM) or small (64 MiB, L2)
  const int runs[] = {1, 2, 3, 4, 6, 9, 18, 32};
  for (int fp = 0; fp < 2; ++fp) {
    const size_t words = fp == 0 ? N : ((size_t)1 << 23);
    for (int contiguous = 0; contiguous < 2; ++contiguous)
      for (int ri = 0; ri < 8; ++ri) {
        const int run = runs[ri];
        if (contiguous && ri > 0) continue;
        size_t n_groups = words / run;
        size_t mask = 1; while (mask < n_groups) mask <<= 1; mask -= 1;
        const size_t ops = (size_t)1 << 27;
        float t0 = timeit([&] { k_scatter<0><<<sms * 16, 256>>>(a, n_groups, run, ops, contiguous, mask); }, 3);
        float t1 = timeit([&] { k_scatter<1><<<sms * 16, 256>>>(a, n_groups, run, ops, contiguous, mask); }, 3);
        float t2 = timeit([&] { k_scatter<2><<<sms * 16, 256>>>(a, n_groups, run, ops, contiguous, mask); }, 3);
        printf(", \"scatter_%s_%s_run%d_Gops\": {\"red\": %.2f, \"rmw\": %.2f, \"store\": %.2f}",
               fp == 0 ? "2GiB" : "64MiB", contiguous ? "contig" : "rand", run,
               ops / t0 / 1e6, ops / t1 / 1e6, ops / t2 / 1e6);
      }
  }
  printf("}\n");
  return 0;
}
