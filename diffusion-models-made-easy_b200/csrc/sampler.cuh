// Device-side pieces of the per-step sampler updates shared by sampler.cu (stand-alone elementwise kernels) and
// conv_out_tc.cu (the same updates applied in the output conv's epilogue, where eps is still in registers).
// The arithmetic follows the reference operation by operation (separately rounded mul / sub / add, IEEE sqrt and div):
// given the same eps and the same noise both paths produce the same bits.
#pragma once
#include "common.cuh"

namespace dmme {

// ---- Philox4x32-10 -----------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// schedule-table index: negative values wrap like torch indexing (table[-1] is the last entry), anything still outside
// [0, len) is clamped -- a replayed CUDA graph cannot raise; the eager wrappers raise IndexError on the host instead
__device__ __forceinline__ long long table_index(long long t, int len) {
  if (t < 0) t += len;
  return t < 0 ? 0 : (t >= len ? len - 1 : t);
}
__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x) + 0.5f) * 2.3283064365386963e-10f; }
// four standard normals for element group g of stream sid
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long sid, unsigned long long g) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32),
                                           static_cast<uint32_t>(sid), static_cast<uint32_t>(sid >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  float4 z;
  float s, c;
  float rad = sqrtf(-2.0f * logf(u01(r.x)));
  sincospif(2.0f * u01(r.y), &s, &c);
  z.x = rad * c; z.y = rad * s;
  rad = sqrtf(-2.0f * logf(u01(r.z)));
  sincospif(2.0f * u01(r.w), &s, &c);
  z.z = rad * c; z.w = rad * s;
  return z;
}


// ---- per-step scalars (read once per kernel from the device tables at *t_ptr) and per-element updates ---------------
struct DdpmScalars { float c1, c2, sd; bool last; long long t; };
__device__ __forceinline__ DdpmScalars ddpm_scalars(const float* beta, const float* alpha, const float* alpha_bar,
                                                    const int64_t* t_ptr, int table_len) {
  DdpmScalars s;
  s.t = *t_ptr;
  const long long ti = table_index(s.t, table_len);
  const float b = beta[ti], a = alpha[ti], ab = alpha_bar[ti];
  s.c1 = __fdiv_rn(1.0f, __fsqrt_rn(a));
  s.c2 = __fdiv_rn(b, __fsqrt_rn(__fsub_rn(1.0f, ab)));
  s.sd = __fsqrt_rn(b);
  s.last = (s.t == 1);
  return s;
}
// x_{t-1} = where(t == 1, mean, mean + sqrt(beta_t) z), mean = 1/sqrt(alpha_t) (x - beta_t/sqrt(1-abar_t) eps)
__device__ __forceinline__ float ddpm_update(float x, float eps, float z, const DdpmScalars& s) {
  const float mean = __fmul_rn(s.c1, __fsub_rn(x, __fmul_rn(s.c2, eps)));
  return s.last ? mean : __fadd_rn(__fmul_rn(z, s.sd), mean);
}

struct DdimScalars { float s1, sp; };
__device__ __forceinline__ DdimScalars ddim_scalars(const float* alpha_bar, const int64_t* tau, const int64_t* i_ptr,
                                                    int table_len, int tau_len) {
  const long long i = *i_ptr;
  const float ab_i = alpha_bar[table_index(tau[table_index(i, tau_len)], table_len)];
  const float ab_p = alpha_bar[table_index(tau[table_index(i - 1, tau_len)], table_len)];
  DdimScalars s;
  s.s1 = __fsqrt_rn(__fsub_rn(1.0f, ab_i));
  s.sp = __fsqrt_rn(ab_p);
  return s;
}
// as written in equations/ddim/ddim.py:52-57: x0 = (x - sqrt(1-abar_i) eps) / sqrt(abar_prev); x <- sqrt(abar_prev) x0
__device__ __forceinline__ float ddim_update(float x, float eps, const DdimScalars& s) {
  return __fmul_rn(s.sp, __fdiv_rn(__fsub_rn(x, __fmul_rn(s.s1, eps)), s.sp));
}

struct IddpmScalars { float c1, c2, log_b, log_bt; bool last; long long t; };
__device__ __forceinline__ IddpmScalars iddpm_scalars(const float* beta, const float* alpha, const float* alpha_bar,
                                                      const int64_t* t_ptr, int table_len) {
  IddpmScalars s;
  s.t = *t_ptr;
  const long long ti = table_index(s.t, table_len), tp = table_index(s.t - 1, table_len);
  const float b = beta[ti], a = alpha[ti], ab = alpha_bar[ti], abp = alpha_bar[tp];
  s.c1 = __fdiv_rn(1.0f, __fsqrt_rn(a));
  s.c2 = __fdiv_rn(b, __fsqrt_rn(__fsub_rn(1.0f, ab)));
  const float bt = __fmul_rn(__fdiv_rn(__fsub_rn(1.0f, abp), __fsub_rn(1.0f, ab)), b);
  s.log_b = logf(b);
  s.log_bt = logf(fmaxf(bt, 1e-12f));
  s.last = (s.t == 1);
  return s;
}
__device__ __forceinline__ float iddpm_update(float x, float e, float v, float z, const IddpmScalars& s) {
  const float var = expf(__fadd_rn(__fmul_rn(v, s.log_b), __fmul_rn(__fsub_rn(1.0f, v), s.log_bt)));
  const float mean = __fmul_rn(s.c1, __fsub_rn(x, __fmul_rn(s.c2, e)));
  return s.last ? mean : __fadd_rn(__fmul_rn(z, __fsqrt_rn(var)), mean);
}

}  // namespace dmme
