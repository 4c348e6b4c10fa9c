// Does an L2 prefetch make a later read of a weight matrix faster on B200, and which instruction actually does it?
//   1. flush L2 (stream 512 MB of other data)   2. prefetch a 32 MB region with method M   3. wait 30 us
//   4. time a full read of the region by 148 CTAs (ld.global.nc v4)  ->  GB/s
// methods: 0 none (cold, HBM)  1 cp.async.bulk.prefetch.L2 (2 KB pieces)  2 prefetch.global.L2 per 128-byte line
//          3 ld.global.cg.L2::128B.u32 one word per line, result dropped  4 the timed read itself run twice (warm)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/bin/l2_prefetch_test scripts/l2_prefetch_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void __launch_bounds__(512) read_kernel(const uint4* w, unsigned long long n16, unsigned long long* sink) {
  uint32_t acc = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * 512;
  for (unsigned long long i = (unsigned long long)blockIdx.x * 512 + threadIdx.x; i + 7 * stride < n16; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(w + i + j * stride));
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(32) prefetch_kernel(const char* w, unsigned long long bytes, int method, unsigned long long* sink) {
  // one warp per CTA (like the decode kernel's prefetch warp), each CTA its contiguous share
  const unsigned long long per = (bytes / gridDim.x) & ~4095ull;      // (the tail of the region stays cold: < 1 %)
  const char* p = w + (unsigned long long)blockIdx.x * per;
  const int lane = threadIdx.x;
  uint32_t acc = 0;
  if (method == 1) {
    for (unsigned long long o = (unsigned long long)lane * 2048; o < per; o += 32ull * 2048)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + o), "r"(2048u) : "memory");
  } else if (method == 2) {
    for (unsigned long long o = (unsigned long long)lane * 128; o < per; o += 32ull * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o) : "memory");
  } else if (method == 3) {
    for (unsigned long long o = (unsigned long long)lane * 128; o < per; o += 32ull * 128) {
      uint32_t v;
      asm volatile("ld.global.cg.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p + o) : "memory");
      acc ^= v;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void spin_kernel(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) { } }

int main() {
  const unsigned long long region = 32ull << 20, flush = 512ull << 20;
  char *w, *f; unsigned long long* sink;
  cudaMalloc(&w, region); cudaMalloc(&f, flush); cudaMalloc(&sink, 64);
  cudaMemset(w, 1, region); cudaMemset(f, 2, flush);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"none (cold)", "cp.async.bulk.prefetch.L2 2KB", "prefetch.global.L2 per line", "ld.global.cg.L2::128B one word per line", "previous read (warm)"};
  for (int method = 0; method < 5; ++method) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      read_kernel<<<148, 512>>>(reinterpret_cast<const uint4*>(f), flush / 16, sink);          // flush L2
      if (method == 4) read_kernel<<<148, 512>>>(reinterpret_cast<const uint4*>(w), region / 16, sink);
      else if (method > 0) prefetch_kernel<<<148, 32>>>(w, region, method, sink);
      spin_kernel<<<1, 1>>>(60000);                                                             // ~30 us for the prefetches to land
      cudaEventRecord(e0);
      read_kernel<<<148, 512>>>(reinterpret_cast<const uint4*>(w), region / 16, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-44s read of 32 MB: %7.2f us = %8.1f GB/s  (%s)\n", names[method], best * 1000, region / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
