// peaks.cu — measured on-chip bandwidth peaks of this B200, for the denominators of the render kernel's memory-side
// numbers (north_star: "achieved L2/HBM GB/s on BVH node fetches against the chip's peaks"):
//   smem_stream   conflict-free LDS.128, every lane its own 16 B column             -> shared-memory peak
//   smem_random   LDS.128 x 3 + LDS.64 of a random 64 B record per lane             -> the node-fetch pattern (bank conflicts included)
//   l1_hit        LDG.128 (.ca) over a 32 KB per-CTA window                          -> L1 hit bandwidth
//   l2_hit        LDG.128 (.cg, bypasses L1) over a 48 MB buffer, resident in L2     -> L2 bandwidth
//   dram          LDG.128 (.cg) over 4 GB                                            -> HBM (cross-check of MEASURED_PEAKS.json)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/peaks tools/peaks/peaks.cu      Run: build/peaks > peaks.json
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  // ld.volatile: the stream kernel's addresses are loop-invariant and ptxas hoists a plain ld.shared out of the loop
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

// mode 0: conflict-free stream; mode 1: random 64-byte records (4 loads per record: 16+16+16+8 bytes)
// consumes all four words of a 128-bit load with one ALU-pipe and one FMA-pipe instruction (ptxas narrows a vector load
// whose upper words are unused)
__device__ __forceinline__ void eat(float4 v, uint32_t& a, uint32_t& m) {
  a ^= __float_as_uint(v.x) ^ __float_as_uint(v.y);
  m = __float_as_uint(v.z) * __float_as_uint(v.w) + m;
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) smem_kernel(int iters, int n_records, float* sink) {
  extern __shared__ float4 s[];
  for (int i = threadIdx.x; i < n_records * 4; i += blockDim.x) s[i] = make_float4(float(i), 1.f, 2.f, 3.f);
  __syncthreads();
  const uint32_t base = uint32_t(__cvta_generic_to_shared(s));
  float acc = 0.f;
  uint32_t ea = 0u, em = 0u;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  if (MODE == 0) {
    // 8 loads per iteration at immediate offsets from one address register, one FADD per load: nothing but the LDS
    // pipe can be the bound (an earlier version with per-load address arithmetic measured the ALU pipe: 64 B/clk/SM)
    const uint32_t a = base + 16u * threadIdx.x;  // 1024 threads x 16 B = 16 KB per "row"; 8 rows = the 128 KB
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) eat(lds128(a + uint32_t(u) * 16384u), ea, em);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        x = x * 1664525u + 1013904223u;
        const uint32_t p = base + ((x >> 2) & (uint32_t(n_records - 1) << 6));  // n_records is a power of two
        float4 a = lds128(p), b = lds128(p + 16u), c = lds128(p + 32u);
        float2 d = lds64(p + 48u);
        eat(a, ea, em), eat(b, ea, em), eat(c, ea, em);
        ea ^= __float_as_uint(d.x) ^ __float_as_uint(d.y);
      }
    }
  }
  if (acc == 123.456f || (ea ^ em) == 0x12345u) sink[0] = acc;
}

// CG: bypass L1 (ld.global.cg) or not (.ca); each CTA walks `window` bytes starting at its own offset, `passes` times
template <bool CG>
__global__ void __launch_bounds__(512) gmem_kernel(const float4* __restrict__ buf, size_t window_vec, size_t cta_stride_vec, int passes, float* sink) {
  const float4* p = buf + size_t(blockIdx.x) * cta_stride_vec;
  float acc = 0.f;
  uint32_t ea = 0u, em = 0u;
  for (int k = 0; k < passes; k++)
    for (size_t i = threadIdx.x; i + 3 * blockDim.x < window_vec; i += 4 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float4 v;
        if (CG) asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i + u * blockDim.x));
        else asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i + u * blockDim.x));
        eat(v, ea, em);
      }
    }
  if (acc == 123.456f || (ea ^ em) == 0x12345u) sink[0] = acc;
}

__global__ void __launch_bounds__(512) l2_kernel(const float4* __restrict__ buf, size_t total_vec, int passes, float* sink) {
  uint32_t ea = 0u, em = 0u;
  const size_t start = (size_t(blockIdx.x) * 7919u * 512u) % total_vec;
  for (int k = 0; k < passes; k++)
    for (size_t j = threadIdx.x; j + 3 * blockDim.x < total_vec; j += 4 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        size_t i = start + j + u * blockDim.x;
        if (i >= total_vec) i -= total_vec;
        float4 v;
        asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
        eat(v, ea, em);
      }
    }
  if ((ea ^ em) == 0x12345u) sink[0] = 1.f;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  float* sink;
  CK(cudaMalloc(&sink, 4));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d", pr.name, sms, clock_khz / 1000);
  {  // shared memory
    const int n_records = 2048;  // 128 KB of 64-byte records (book2_final stages 1,095 nodes = 70 KB)
    const size_t smem = size_t(n_records) * 64;
    CK(cudaFuncSetAttribute(smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    CK(cudaFuncSetAttribute(smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int iters = 4000;
    double ms0 = time_ms([&] { smem_kernel<0><<<sms, 1024, smem>>>(iters, n_records, sink); }, 5);
    double bytes0 = double(sms) * 1024 * double(iters) * 8 * 16;
    double ms1 = time_ms([&] { smem_kernel<1><<<sms, 1024, smem>>>(iters, n_records, sink); }, 5);
    double bytes1 = double(sms) * 1024 * double(iters) * 4 * 56;
    printf(", \"smem_stream_ms\": %.3f, \"smem_random_ms\": %.3f", ms0, ms1);
    printf(", \"smem_stream_GBps\": %.0f, \"smem_random_record_GBps\": %.0f, \"smem_random_records_per_s\": %.4g", bytes0 / ms0 / 1e6, bytes1 / ms1 / 1e6,
           double(sms) * 1024 * double(iters) * 4 / (ms1 * 1e-3));
  }
  {  // L1, L2, DRAM
    const size_t big = size_t(4) << 30;
    float4* buf;
    CK(cudaMalloc(&buf, big));
    CK(cudaMemset(buf, 1, big));
    // L1: 32 KB per CTA, one CTA per SM, many passes
    {
      const size_t win = (32 << 10) / 16;
      const int passes = 2000;
      double ms = time_ms([&] { gmem_kernel<false><<<sms, 512>>>(buf, win, win, passes, sink); }, 5);
      printf(", \"l1_hit_GBps\": %.0f", double(sms) * 32768.0 * passes / ms / 1e6);
    }
    // L2: a 48 MB buffer, every CTA reads ALL of it (from its own starting offset, wrapping) a few times: the footprint per
    // SM is far beyond L1, the whole of it stays in the 126 MB L2
    {
      const int ctas = 4 * sms;
      const size_t total_vec = (size_t(48) << 20) / 16;
      const int passes = 2;
      double ms = time_ms([&] { l2_kernel<<<ctas, 512>>>(buf, total_vec, passes, sink); }, 5);
      printf(", \"l2_hit_GBps\": %.0f, \"l2_footprint_MB\": %.1f, \"l2_ms\": %.3f", double(ctas) * double(total_vec) * 16 * passes / ms / 1e6, double(total_vec) * 16 / 1048576.0, ms);
    }
    // DRAM: 4 GB once
    {
      const int ctas = 8 * sms;
      const size_t slice = big / 16 / ctas;
      double ms = time_ms([&] { gmem_kernel<true><<<ctas, 512>>>(buf, slice, slice, 1, sink); }, 5);
      printf(", \"dram_read_GBps\": %.0f", double(ctas) * double(slice) * 16 / ms / 1e6);
    }
    CK(cudaFree(buf));
  }
  printf("}\n");
  return 0;
}
