// Microbenchmark of the decode kernels' weight ring (sv_ring.cuh): how fast can 148 CTAs pull a > L2 buffer through a
// shared-memory ring with cp.async.bulk, as a function of copy granularity, slot size, ring depth and consumer work?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/bin/ring_stream scripts/ring_stream.cu
//   scripts/bin/ring_stream            (prints one line per variant: GB/s over a 2 GB sweep, best of 3)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define DEVINL __device__ __forceinline__
DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
DEVINL void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (it > (1u << 24)) __trap();
  }
}
DEVINL void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
DEVINL uint4 lds16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

struct P {
  const uint8_t* w;
  unsigned long long bytes_per_cta;   // each CTA streams its own contiguous region
  unsigned long long window;          // != 0: the CTA re-reads only the first `window` bytes of its region (L2 resident)
  int rows, row_bytes, ksplit, pitch, stages, slot_bytes, consume, one_lane;   // ksplit: slots per tile (row length = ksplit * row_bytes)
  unsigned long long* sink;
};

// slot = `rows` pieces of `row_bytes` (source stride `row_stride`, smem pitch `pitch`); rows == 1: one contiguous copy
__global__ void __launch_bounds__(288, 1) ring_kernel(const P p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem), full0 = base + p.stages * p.slot_bytes, empty0 = full0 + 256;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8u * s, 1); mbar_init(empty0 + 8u * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long slot_payload = (unsigned long long)p.rows * p.row_bytes;
  const unsigned long long nslots = p.bytes_per_cta / slot_payload;
  const uint8_t* src0 = p.w + (unsigned long long)blockIdx.x * p.bytes_per_cta;
  uint32_t slot = 0, phase = 0;
  if (warp == 8) {
    for (unsigned long long i = 0; i < nslots; ++i) {
      const uint32_t fb = full0 + 8u * slot;
      if (lane == 0) { mbar_wait(empty0 + 8u * slot, phase ^ 1u); mbar_expect_tx(fb, (uint32_t)slot_payload); }
      __syncwarp();
      // the source walks the CTA's region in slot order; inside a slot, rows are row_stride apart (wrapping in the region)
      const unsigned long long tile = i / p.ksplit, kh = i % p.ksplit, row_stride = (unsigned long long)p.ksplit * p.row_bytes;
      const unsigned long long off = tile * p.rows * row_stride + kh * p.row_bytes;
      const uint8_t* s = src0 + (p.window ? off % p.window : off);
      if (p.one_lane) {
        if (lane == 0) for (int r = 0; r < p.rows; ++r) bulk_g2s(base + slot * p.slot_bytes + r * p.pitch, s + (unsigned long long)r * row_stride, p.row_bytes, fb);
      } else if (lane < p.rows) {
        bulk_g2s(base + slot * p.slot_bytes + lane * p.pitch, s + (unsigned long long)lane * row_stride, p.row_bytes, fb);
      }
      if (++slot == (uint32_t)p.stages) { slot = 0; phase ^= 1u; }
    }
    return;
  }
  uint32_t acc = 0;
  for (unsigned long long i = 0; i < nslots; ++i) {
    mbar_wait(full0 + 8u * slot, phase);
    if (p.consume) {      // the GEMV's fragment reads: every warp reads 1/8 of the slot with 16-byte loads
      const uint32_t sb = base + slot * p.slot_bytes;
      const int per_warp = (int)(slot_payload / 8);
      for (int o = lane * 16; o < per_warp; o += 512) { const uint4 v = lds16(sb + warp * per_warp + o); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8u * slot);
    if (++slot == (uint32_t)p.stages) { slot = 0; phase ^= 1u; }
  }
  if (acc == 0x12345678u) p.sink[0] = acc;
}

// ceiling: plain 16-byte streaming loads, 8 in flight per thread
__global__ void __launch_bounds__(512, 1) ldg_kernel(const uint4* w, unsigned long long n16_per_cta, unsigned long long* sink, unsigned long long win16 = 0) {
  const uint4* p = w + (unsigned long long)blockIdx.x * n16_per_cta;
  uint32_t acc = 0;
  for (unsigned long long i = threadIdx.x; i + 7 * 512 < n16_per_cta; i += 8 * 512) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(p + (win16 ? (i + j * 512) % win16 : (i + j * 512))));
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  const int ncta = 148;
  const unsigned long long total = 2ull << 30;
  uint8_t* w; unsigned long long* sink;
  cudaMalloc(&w, total + (1 << 20)); cudaMalloc(&sink, 64);
  cudaMemset(w, 1, total);
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct V { const char* name; int rows, row_bytes, pitch, stages, consume, one_lane, ksplit; };
  std::vector<V> vs = {
      {"16x2KB pitch2112 5 stages, consume (the decode ring today)", 16, 2048, 2112, 5, 1, 0, 1},
      {"14x2KB pitch2112 5 stages, consume (14-row tiles)", 14, 2048, 2112, 5, 1, 0, 1},
      {"16x2KB pitch2112 5 stages, no consumer reads", 16, 2048, 2112, 5, 0, 0, 1},
      {"16x2KB issued by one lane, 5 stages, consume", 16, 2048, 2112, 5, 1, 1, 1},
      {"1x32KB contiguous, 5 stages, consume", 1, 32768, 32768, 5, 1, 0, 1},
      {"1x32KB contiguous, 6 stages, consume", 1, 32768, 32768, 6, 1, 0, 1},
      {"1x16KB contiguous, 12 stages, consume", 1, 16384, 16384, 12, 1, 0, 1},
      {"1x8KB contiguous, 24 stages, consume", 1, 8192, 8192, 24, 1, 0, 1},
      {"8x4KB pitch4160, 6 stages, consume", 8, 4096, 4160, 6, 1, 0, 1},
      {"16x2KB pitch2112 3 stages, consume", 16, 2048, 2112, 3, 1, 0, 1},
      {"16x2KB pitch2112 6 stages, consume", 16, 2048, 2112, 6, 1, 0, 1},
      {"32x1KB pitch1088 6 stages, consume", 32, 1024, 1088, 6, 1, 0, 1},
      {"16x2KB, rows 4 KB apart (K=2048 weights), 5 stages", 16, 2048, 2112, 5, 1, 0, 2},
      {"14x2KB, rows 16 KB apart (K=8192 weights), 5 stages", 14, 2048, 2112, 5, 1, 0, 8},
      {"14x2KB, rows 16 KB apart, one lane issues", 14, 2048, 2112, 5, 1, 1, 8},
  };
  for (const V& v : vs) {
    P p{};
    p.w = w; p.rows = v.rows; p.row_bytes = v.row_bytes; p.pitch = v.pitch; p.stages = v.stages; p.consume = v.consume; p.one_lane = v.one_lane; p.ksplit = v.ksplit;
    p.slot_bytes = v.rows * v.pitch; p.sink = sink;
    const unsigned long long payload = (unsigned long long)v.rows * v.row_bytes * v.ksplit;
    p.bytes_per_cta = total / ncta / payload * payload;
    const int smem = v.stages * p.slot_bytes + 512 + 128;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      ring_kernel<<<ncta, 288, smem>>>(p);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("%-62s smem %6d  %8.1f GB/s  (%s)\n", v.name, smem, p.bytes_per_cta * ncta / best * 1e-6, cudaGetErrorString(err));
  }
  {
    float best = 1e30f;
    const unsigned long long n16 = total / 16 / ncta;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      ldg_kernel<<<ncta, 512>>>(reinterpret_cast<const uint4*>(w), n16, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-62s              %8.1f GB/s\n", "ld.global.nc v4, 512 threads x 8 in flight per SM", n16 * 16 * ncta / best * 1e-6);
  }
  // L2-resident source: how fast can one SM's ring be filled when HBM is not involved?
  {
    float best = 1e30f;
    const unsigned long long n16 = total / 16 / ncta;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      ldg_kernel<<<ncta, 512>>>(reinterpret_cast<const uint4*>(w), n16, sink, 131072 / 16);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("ld.global.nc v4, 512 threads x 8 in flight, L2-RESIDENT (128 KB window per CTA): %8.1f GB/s = %.1f B/clk/SM\n", n16 * 16 * ncta / best * 1e-6,
           n16 * 16 / best * 1e-6 / 1.9);
  }
  struct R { int rows, row_bytes, pitch, stages; };
  for (const R& r : {R{8, 4096, 4160, 5}, R{4, 8192, 8256, 5}, R{2, 16384, 16448, 5}, R{1, 32768, 32768, 5}, R{1, 16384, 16384, 10}, R{16, 2048, 2112, 5}}) {
    P p{};
    p.w = w; p.rows = r.rows; p.row_bytes = r.row_bytes; p.pitch = r.pitch; p.stages = r.stages; p.consume = 1; p.slot_bytes = r.rows * r.pitch; p.sink = sink; p.ksplit = 1;
    p.bytes_per_cta = (total / ncta) / 32768 * 32768; p.window = 131072;
    const int smem = r.stages * p.slot_bytes + 640;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      ring_kernel<<<ncta, 288, smem>>>(p);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%dx%dKB, %d stages, L2-RESIDENT source: %8.1f GB/s = %.1f B/clk/SM (%s)\n", r.rows, r.row_bytes / 1024, r.stages,
           p.bytes_per_cta * ncta / best * 1e-6, p.bytes_per_cta / best * 1e-6 / 1.9, cudaGetErrorString(cudaGetLastError()));
  }
  for (int stages : {5, 3}) {
    P p{};
    p.w = w; p.rows = 16; p.row_bytes = 2048; p.pitch = 2112; p.stages = stages; p.consume = 1; p.slot_bytes = 16 * 2112; p.sink = sink; p.ksplit = 1;
    p.bytes_per_cta = (total / ncta) / 32768 * 32768; p.window = 131072;          // 148 x 128 KB = 19 MB working set
    const int smem = stages * p.slot_bytes + 640;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      ring_kernel<<<ncta, 288, smem>>>(p);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("16x2KB, %d stages, L2-RESIDENT source (128 KB window per CTA): %8.1f GB/s = %.1f B/clk/SM at 1.9 GHz (%s)\n", stages,
           p.bytes_per_cta * ncta / best * 1e-6, p.bytes_per_cta / best * 1e-6 / 1.9, cudaGetErrorString(cudaGetLastError()));
  }
  // phase-sized bursts: the same ring run for only ~226 KB per CTA (one fc phase), launched back to back
  {
    P p{};
    p.w = w; p.rows = 14; p.row_bytes = 2048; p.pitch = 2112; p.stages = 5; p.consume = 1; p.slot_bytes = 14 * 2112; p.sink = sink; p.ksplit = 1;
    p.bytes_per_cta = 8ull * 14 * 2048;
    const int smem = 5 * p.slot_bytes + 640;
    cudaEventRecord(e0);
    for (int i = 0; i < 200; ++i) { p.w = w + (unsigned long long)(i % 50) * (40ull << 20); ring_kernel<<<ncta, 288, smem>>>(p); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("fc-sized launches (33.9 MB each, 200 back to back): %.2f us per launch = %.1f GB/s\n", ms * 1000 / 200, p.bytes_per_cta * ncta * 200 / ms * 1e-6);
  }
  return 0;
}
