// Epilogue pieces shared by the tcgen05 convolution kernels: the per-row post-accumulation math
// (bias + timestep embedding + residual addend -> bf16 store) and the GroupNorm statistics of the stored tensor.
#pragma once
#include "common.cuh"

namespace dmme {

// f[0..32) += bias[col..] + temb_row[col..] + addend_row[col..] (each optional)
__device__ __forceinline__ void epi_add_terms(float (&f)[32], const float* __restrict__ bias,
                                              const float* __restrict__ trow, const __nv_bfloat16* __restrict__ arow,
                                              int col) {
  if (bias) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col + i));
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
  if (trow) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + col + i));
      f[i] += t4.x; f[i + 1] += t4.y; f[i + 2] += t4.z; f[i + 3] += t4.w;
    }
  }
  if (arow) {
    const uint4* ap = reinterpret_cast<const uint4*>(arow + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 a4 = __ldg(ap + i);
      float lo, hi;
      unpack_bf16x2(a4.x, lo, hi); f[8 * i + 0] += lo; f[8 * i + 1] += hi;
      unpack_bf16x2(a4.y, lo, hi); f[8 * i + 2] += lo; f[8 * i + 3] += hi;
      unpack_bf16x2(a4.z, lo, hi); f[8 * i + 4] += lo; f[8 * i + 5] += hi;
      unpack_bf16x2(a4.w, lo, hi); f[8 * i + 6] += lo; f[8 * i + 7] += hi;
    }
  }
}

// round f to bf16, store 32 channels (64 bytes) at dst, and leave the ROUNDED values in f
__device__ __forceinline__ void epi_store_bf16(float (&f)[32], __nv_bfloat16* dst) {
  uint4* dp = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o;
    o.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
    o.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
    o.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
    o.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
    dp[i] = o;
    unpack_bf16x2(o.x, f[8 * i + 0], f[8 * i + 1]);
    unpack_bf16x2(o.y, f[8 * i + 2], f[8 * i + 3]);
    unpack_bf16x2(o.z, f[8 * i + 4], f[8 * i + 5]);
    unpack_bf16x2(o.w, f[8 * i + 6], f[8 * i + 7]);
  }
}

// GroupNorm statistics of one warp's 32 rows x 32 columns: for each of the 8 four-channel micro-groups the sum
// and the sum of squares over the rows with `mine` set, added to st[(col/4 + g) * 2 + {0,1}] as 2^-20 fixed point.
// Recursive-halving butterfly: 16 shuffles instead of 80, and 16 lanes issue one atomic each.
__device__ __forceinline__ void epi_stats_warp32(const float (&f)[32], bool mine, int lane, unsigned long long* st) {
  float v[16];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float a = f[4 * g], b = f[4 * g + 1], c = f[4 * g + 2], d = f[4 * g + 3];
    v[2 * g] = mine ? (a + b) + (c + d) : 0.f;
    v[2 * g + 1] = mine ? (a * a + b * b) + (c * c + d * d) : 0.f;
  }
  // after the step with offset `off`, a lane keeps the half of its values selected by its `off` bit
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
  float w8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b16 ? v[i] : v[i + 8];
    const float keep = b16 ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b8 ? w8[i] : w8[i + 4];
    const float keep = b8 ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float w2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b4 ? w4[i] : w4[i + 2];
    const float keep = b4 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = b2 ? w2[0] : w2[1];
    const float keep = b2 ? w2[1] : w2[0];
    float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    if ((lane & 1) == 0) {
      const int idx = (b16 ? 8 : 0) + (b8 ? 4 : 0) + (b4 ? 2 : 0) + (b2 ? 1 : 0);  // = 2 * micro-group + {sum, sumsq}
      atomicAdd(st + idx, static_cast<unsigned long long>(
                              __float2ll_rn(w1 * static_cast<float>(1 << DMME_STATS_FRAC_BITS))));
    }
  }
}

// same statistics when a warp's rows split into segments of `seg` (< 32) lanes that belong to different images
__device__ __forceinline__ void epi_stats_segmented(const float (&f)[32], bool valid, int lane, int seg, bool seg_ok,
                                                    unsigned long long* st) {
  float s1[8], s2[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float a = f[4 * g], b = f[4 * g + 1], c = f[4 * g + 2], d = f[4 * g + 3];
    s1[g] = valid ? (a + b) + (c + d) : 0.f;
    s2[g] = valid ? (a * a + b * b) + (c * c + d * d) : 0.f;
  }
  for (int off = seg >> 1; off > 0; off >>= 1) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      s1[g] += __shfl_xor_sync(0xffffffffu, s1[g], off);
      s2[g] += __shfl_xor_sync(0xffffffffu, s2[g], off);
    }
  }
  if ((lane & (seg - 1)) == 0 && seg_ok) {
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      atomicAdd(st + 2 * g, static_cast<unsigned long long>(__float2ll_rn(s1[g] * kFix)));
      atomicAdd(st + 2 * g + 1, static_cast<unsigned long long>(__float2ll_rn(s2[g] * kFix)));
    }
  }
}

}  // namespace dmme
