// Shared host/device helpers for the dmme_b200 kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dmme_b200.h"

namespace dmme {

// ---- host side: error reporting and launch accounting ------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DMME_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ::dmme::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  count_launch();
  return 0;
}

inline int check_launch_err(cudaError_t e, const char* what) {
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  count_launch();
  return 0;
}

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while the previous kernel of the
// stream drains; it MUST call pdl_wait() (ptx_sm100.cuh) before touching memory.  Captured by CUDA graphs as a
// programmatic dependency edge.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// the same for a kernel whose CTAs form clusters of two (cta_group::2 tensor-core kernels)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 2;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// one-time per-DEVICE setup (cudaFuncSetAttribute applies to the current device only): `static DeviceOnce once_;`
struct DeviceOnce {
  bool done[64] = {};
  bool& here() {
    int d = 0;
    cudaGetDevice(&d);
    return done[d & 63];
  }
};
// SM count of the current device (cached per device ordinal)
inline int device_sm_count() {
  static int cache[64] = {};
  int d = 0;
  cudaGetDevice(&d);
  int& c = cache[d & 63];
  if (c == 0) {
    cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, d);
    if (c <= 0) c = 148;
  }
  return c;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its predecessor drains.
//   pdl_wait()    -- returns once every prerequisite grid has completed and its memory is visible; a no-op when the
//                    kernel was launched without the attribute.  Nothing the predecessor wrote (or still reads, for
//                    buffers this kernel overwrites) may be touched before it.
//   pdl_trigger() -- lets the dependent grid's CTAs be scheduled once every CTA of this grid has called it (or exited).
//                    Kernels that own TMEM call it only AFTER tcgen05.alloc: a waiting dependent CTA that had taken the
//                    columns first would block this CTA's allocation forever.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- device side: storage-type access ----------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld_act(const T* p);
template <>
__device__ __forceinline__ float ld_act<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_act<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__device__ __forceinline__ void st_act(T* p, float v);
template <>
__device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_act<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}

// bf16 paths: silu(y) = y * sigmoid(y) = h + h * tanh(h), h = y/2 -- one MUFU (tanh.approx, rel. error 2^-11, far below the
// 2^-9 rounding of the bf16 result) and two FMA-pipe instructions; exp + full-precision division cost ~14 and made the
// streaming GroupNorm pass instruction-bound instead of HBM-bound
__device__ __forceinline__ float silu_f(float y) {
  const float h = 0.5f * y;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float silu_precise(float y) { return y / (1.0f + expf(-y)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace dmme
