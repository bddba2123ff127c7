// Per-step sampler updates in image space (NCHW fp32), one elementwise pass each.
// The arithmetic follows the reference operation by operation (separately rounded mul / sub / add,
// IEEE sqrt and div) so that, given the same eps and the same noise, DDPM and DDIM updates are
// bit-identical to torch on the CPU.  Schedule scalars are read from device tables at index *t_ptr,
// so a captured CUDA graph of one step can be replayed for every t with no host work.
#include "common.cuh"
#include "sampler.cuh"

namespace dmme {

__global__ void philox_normal_kernel(float* __restrict__ out, long long numel, unsigned long long seed,
                                     unsigned long long sid, unsigned long long goff) {
  const long long groups = (numel + 3) / 4;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 z = philox_normal4(seed, sid, goff + g);
    const float zz[4] = {z.x, z.y, z.z, z.w};
    for (int j = 0; j < 4; ++j)
      if (g * 4 + j < numel) out[g * 4 + j] = zz[j];
  }
}

// ---- DDPM ancestral step ---------------------------------------------------------------------
__global__ void ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                 const float* __restrict__ beta, const float* __restrict__ alpha,
                                 const float* __restrict__ alpha_bar, const int64_t* __restrict__ t_ptr, int table_len,
                                 long long numel, unsigned long long seed, unsigned long long goff) {
  pdl_trigger();
  pdl_wait();
  const DdpmScalars sc = ddpm_scalars(beta, alpha, alpha_bar, t_ptr, table_len);
  const long long t = sc.t;
  const bool last = sc.last;
  const long long groups = (numel + 3) / 4;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    if (!noise && !last) {
      const float4 z = philox_normal4(seed, static_cast<unsigned long long>(t), goff + g);
      zz[0] = z.x; zz[1] = z.y; zz[2] = z.z; zz[3] = z.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = g * 4 + j;
      if (i < numel) {
        x[i] = ddpm_update(x[i], eps[i], noise ? noise[i] : zz[j], sc);
      }
    }
  }
}

// ---- DDIM step, as written in the reference --------------------------------------------------
__global__ void ddim_step_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                 const float* __restrict__ alpha_bar, const int64_t* __restrict__ tau,
                                 const int64_t* __restrict__ i_ptr, int table_len, int tau_len, long long numel) {
  const DdimScalars sc = ddim_scalars(alpha_bar, tau, i_ptr, table_len, tau_len);
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < numel;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    x[e] = ddim_update(x[e], eps[e], sc);
  }
}

// ---- IDDPM learned-variance step -------------------------------------------------------------
__global__ void iddpm_step_kernel(float* __restrict__ x, const float* __restrict__ mo, const float* __restrict__ noise,
                                  const float* __restrict__ beta, const float* __restrict__ alpha,
                                  const float* __restrict__ alpha_bar, const int64_t* __restrict__ t_ptr, int table_len,
                                  int n, int c, int hw, unsigned long long seed, unsigned long long goff) {
  const IddpmScalars sc = iddpm_scalars(beta, alpha, alpha_bar, t_ptr, table_len);
  const long long t = sc.t;
  const bool last = sc.last;
  const long long numel = static_cast<long long>(n) * c * hw;
  const long long groups = (numel + 3) / 4;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    if (!noise && !last) {
      const float4 z = philox_normal4(seed, static_cast<unsigned long long>(t), goff + g);
      zz[0] = z.x; zz[1] = z.y; zz[2] = z.z; zz[3] = z.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = g * 4 + j;
      if (i < numel) {
        const long long l = i % hw, nc = i / hw;
        const long long ch = nc % c, ni = nc / c;
        const float e = mo[(ni * 2 * c + ch) * hw + l];
        const float v = mo[(ni * 2 * c + c + ch) * hw + l];
        x[i] = iddpm_update(x[i], e, v, noise ? noise[i] : zz[j], sc);
      }
    }
  }
}

__global__ void gather_i64_kernel(const int64_t* table, const int64_t* idx, int64_t* out) { *out = table[*idx]; }
__global__ void add_i64_kernel(int64_t* p, int64_t d) {
  pdl_trigger();
  pdl_wait();
  *p += d;
}

static int ew_grid(long long work, int threads) {
  long long b = ceil_div_ll(work, threads);
  const long long cap = 148LL * 8;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_ddpm_step(float* x, const float* eps, const float* noise, const float* beta, const float* alpha,
                              const float* alpha_bar, const int64_t* t_ptr, int table_len, long long numel,
                              unsigned long long seed, unsigned long long noise_offset, void* stream) {
  DMME_REQUIRE(x && eps && beta && alpha && alpha_bar && t_ptr && numel > 0 && table_len > 0, DMME_E_BADARG,
               "ddpm_step: bad arguments");
  DMME_REQUIRE(noise_offset % 4 == 0, DMME_E_BADARG, "ddpm_step: noise_offset must be a multiple of 4");
  return check_launch_err(launch_pdl(ddpm_step_kernel, dim3(ew_grid((numel + 3) / 4, 256)), dim3(256), 0,
                                     static_cast<cudaStream_t>(stream), x, eps, noise, beta, alpha, alpha_bar, t_ptr,
                                     table_len, numel, seed, noise_offset / 4),
                          "ddpm_step_kernel");
}

extern "C" int dmme_ddim_step(float* x, const float* eps, const float* alpha_bar, const int64_t* tau,
                              const int64_t* i_ptr, int table_len, int tau_len, long long numel, void* stream) {
  DMME_REQUIRE(x && eps && alpha_bar && tau && i_ptr && numel > 0 && table_len > 0 && tau_len > 0, DMME_E_BADARG,
               "ddim_step: bad arguments");
  ddim_step_kernel<<<ew_grid(numel, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, alpha_bar, tau, i_ptr, table_len,
                                                                                     tau_len, numel);
  return check_launch("ddim_step_kernel");
}

extern "C" int dmme_iddpm_step(float* x, const float* model_out, const float* noise, const float* beta,
                               const float* alpha, const float* alpha_bar, const int64_t* t_ptr, int table_len, int n, int c,
                               int hw, unsigned long long seed, unsigned long long noise_offset, void* stream) {
  DMME_REQUIRE(x && model_out && beta && alpha && alpha_bar && t_ptr && table_len > 0 && n > 0 && c > 0 && hw > 0, DMME_E_BADARG,
               "iddpm_step: bad arguments");
  DMME_REQUIRE(noise_offset % 4 == 0, DMME_E_BADARG, "iddpm_step: noise_offset must be a multiple of 4");
  const long long numel = static_cast<long long>(n) * c * hw;
  iddpm_step_kernel<<<ew_grid((numel + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, model_out, noise, beta, alpha, alpha_bar, t_ptr, table_len, n, c, hw, seed, noise_offset / 4);
  return check_launch("iddpm_step_kernel");
}

extern "C" int dmme_gather_i64(const int64_t* table, const int64_t* idx_ptr, int64_t* out, void* stream) {
  DMME_REQUIRE(table && idx_ptr && out, DMME_E_BADARG, "gather_i64: null pointer");
  gather_i64_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(table, idx_ptr, out);
  return check_launch("gather_i64_kernel");
}

extern "C" int dmme_add_i64(int64_t* value, int64_t delta, void* stream) {
  DMME_REQUIRE(value, DMME_E_BADARG, "add_i64: null pointer");
  return check_launch_err(launch_pdl(add_i64_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), value, delta),
                          "add_i64_kernel");
}

extern "C" int dmme_philox_normal(float* out, long long numel, unsigned long long seed, unsigned long long stream_id,
                                  unsigned long long noise_offset, void* stream) {
  DMME_REQUIRE(out && numel > 0 && noise_offset % 4 == 0, DMME_E_BADARG, "philox_normal: bad arguments");
  philox_normal_kernel<<<ew_grid((numel + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, numel, seed, stream_id, noise_offset / 4);
  return check_launch("philox_normal_kernel");
}

// ------------------------------------------------------------------------------------------------
// denorm (common/norm.py:9-11): clip((x + 1) / 2, 0, 1), the image-space tail of GenerateImage.generate_img
// (callbacks/generate.py:64-90) and of LitDDPM.test_step (lit_modules/ddpm.py:97-100); optionally also quantised to
// uint8 (round(255 y)) for image logging / FID feature extraction.  One pass: 4 B read, 4 and / or 1 B written.
// ------------------------------------------------------------------------------------------------
namespace dmme {
__global__ void denorm_kernel(const float* __restrict__ x, float* __restrict__ out_f32, uint8_t* __restrict__ out_u8,
                              long long numel) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    // torch: (x + 1) / 2 then clip(0, 1); clip propagates NaN
    float y = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
    y = (y != y) ? y : fminf(fmaxf(y, 0.0f), 1.0f);
    if (out_f32) out_f32[i] = y;
    if (out_u8) out_u8[i] = static_cast<uint8_t>(__float2int_rn(y * 255.0f));
  }
}
}  // namespace dmme

extern "C" int dmme_denorm(const float* x, float* out_f32, uint8_t* out_u8, long long numel, void* stream) {
  DMME_REQUIRE(x && (out_f32 || out_u8) && numel > 0, DMME_E_BADARG, "denorm: bad arguments");
  return check_launch_err(launch_pdl(dmme::denorm_kernel, dim3(ew_grid(numel, 256)), dim3(256), 0,
                                     static_cast<cudaStream_t>(stream), x, out_f32, out_u8, numel),
                          "denorm_kernel");
}
