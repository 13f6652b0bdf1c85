// Host stand-in for <cuda_runtime.h>: lets tools/simt_sim compile csrc/rt_device.cuh with g++ so that the SIMT
// scheduling simulator runs the SAME per-lane code (node_step, hit_*, scatter_ray ...) as the kernels.  Tool only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <limits>
#include <map>
#include <string>
#include <vector>
// (every std header the tool uses is pulled in BEFORE the keyword macros below)
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define asm (void)
#define volatile(...) 0
struct float3 { float x, y, z; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
#define RTB200_SIM_HOST 1
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct float4;
struct float2;
float4 make_float4(float, float, float, float);
float2 make_float2(float, float);
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline unsigned __umulhi(unsigned a, unsigned b) { return unsigned((uint64_t(a) * b) >> 32); }
inline float __sinf(float x) { return std::sin(x); }
inline float __cosf(float x) { return std::cos(x); }
inline float __logf(float x) { return std::log(x); }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline size_t __cvta_generic_to_shared(const void* p) { return size_t(p); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline bool __any_sync(unsigned, bool p) { return p; }
inline bool __all_sync(unsigned, bool p) { return p; }
inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
// one-lane "warp": enough for the device routines the simulator calls per lane (the warp-cooperative ones only have to compile)
template <typename T> inline T __shfl_sync(unsigned, T v, int) { return v; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline float __fsub_ru(float a, float b) { return std::nextafter(a - b, INFINITY); }
struct shim_dim3 { unsigned x = 0, y = 0, z = 0; };
static shim_dim3 threadIdx, blockIdx;
static shim_dim3 blockDim{1, 1, 1};
using std::max;
using std::min;
