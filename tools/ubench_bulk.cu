// ubench_bulk.cu -- how fast does one SM push a shared-memory image to global memory?
//   mode 0: cp.async.bulk.global.shared::cta, one issuing thread
//   mode 1: cp.async.bulk, 32 lanes of one warp issue (compiler waterfall)
//   mode 2: plain 16-byte stores by `nw` warps (LDS.128 + STG.128), coalesced
// per SM: `rows` rows of `bytes` bytes per round, `rounds` rounds, every row to its own global address (footprint >> L2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_bulk tools/ubench_bulk.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void bulk_store(void *g, const void *s, unsigned bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(384, 1) k(char *out, int mode, int rows, int bytes, int rounds, int nw, unsigned long long *clk) {
  extern __shared__ __align__(128) char sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < rows * bytes / 8; k += blockDim.x) reinterpret_cast<double *>(sm)[k] = k;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  const size_t per_round = (size_t)rows * bytes * gridDim.x;
  long long t_issue = 0;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    char *base = out + (size_t)r * per_round + (size_t)blockIdx.x * rows * bytes;
    if (mode == 0) {
      if (tid == 0) {
        const long long a = clock64();
        for (int j = 0; j < rows; ++j) bulk_store(base + (size_t)j * bytes, sm + (size_t)j * bytes, bytes);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        t_issue += clock64() - a;
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      }
    } else if (mode == 1) {
      if (warp == 0) {
        const long long a = clock64();
        for (int j = lane; j < rows; j += 32) bulk_store(base + (size_t)j * bytes, sm + (size_t)j * bytes, bytes);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        t_issue += clock64() - a;
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      }
    } else {
      if (warp < nw) {
        const long long a = clock64();
        const int n16 = rows * bytes / 16;
        for (int k = warp * 32 + lane; k < n16; k += nw * 32) {
          const double2 v = reinterpret_cast<const double2 *>(sm)[k];
          reinterpret_cast<double2 *>(base)[k] = v;
        }
        t_issue += clock64() - a;
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) { atomicAdd(clk, (unsigned long long)(t1 - t0)); atomicAdd(clk + 1, (unsigned long long)t_issue); }
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t cap = (size_t)8 << 30;
  char *out; cudaMalloc(&out, cap);
  unsigned long long *clk; cudaMalloc(&clk, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int sizes[] = {512, 1024, 1936, 3888, 7776, 15552};
  for (int mode = 0; mode < 3; ++mode)
    for (int si = 0; si < 6; ++si)
      for (int nw = (mode == 2 ? 2 : 1); nw <= (mode == 2 ? 12 : 1); nw *= (mode == 2 ? 6 : 2)) {
        const int bytes = sizes[si] / 16 * 16;
        const int rows = (124 * 1024) / bytes;
        const int rounds = (int)(cap / ((size_t)rows * bytes * sms));
        cudaMemset(clk, 0, 16);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<sms, 384, 128 * 1024>>>(out, mode, rows, bytes, rounds, nw, clk);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[2]; cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
        const double tot = (double)rows * bytes * rounds * sms;
        printf("mode %d bytes %5d rows %3d nw %2d: %.2f ms, %.0f GB/s, per round per SM: %.0f clk total, %.0f clk issuing (%.1f per row)  %s\n", mode, bytes, rows, nw, ms,
               tot / ms * 1e-6, (double)h[0] / sms / rounds, (double)h[1] / sms / rounds, (double)h[1] / sms / rounds / rows, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
