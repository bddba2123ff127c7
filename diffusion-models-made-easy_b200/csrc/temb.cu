// Timestep conditioning, all fp32 (this part of the network must not be rounded to bf16).
//   temb_mlp:  sinusoidal embedding -> Linear -> SiLU -> Linear -> SiLU   (models/ddpm.py:211-217, 338-349)
//   temb_proj: every ResBlock.condition Linear batched into one [total][emb] GEMV/GEMM (models/ddpm.py:101-104)
#include "common.cuh"

namespace dmme {

constexpr int kProjRows = 8;

// out[r][o] = act( b[o] + sum_i in[r][i] * W[o][i] ); in = emb rows, or (t != null) the sinusoidal embedding
// [sin(t f_j), cos(t f_j)] computed on the fly.  One warp per output column, 8 rows per CTA.
__global__ void __launch_bounds__(256) temb_linear_kernel(const float* __restrict__ in, const int64_t* __restrict__ t,
                                                          const float* __restrict__ freq, int rows, int in_dim,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          int out_dim, int act, float* __restrict__ out) {
  extern __shared__ float es[];  // [kProjRows][in_dim]
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.y * kProjRows;
  for (int idx = tid; idx < kProjRows * in_dim; idx += 256) {
    const int r = idx / in_dim, i = idx - r * in_dim;
    float v = 0.f;
    if (r0 + r < rows) {
      if (t) {
        // int64 * float32 promotes to float32 in the reference (models/ddpm.py:347)
        const int half = in_dim / 2;
        const float a = static_cast<float>(t[r0 + r]) * freq[i < half ? i : i - half];
        v = i < half ? sinf(a) : cosf(a);
      } else {
        v = in[static_cast<long long>(r0 + r) * in_dim + i];
      }
    }
    es[idx] = v;
  }
  __syncthreads();
  const int o = blockIdx.x * 8 + warp;
  if (o >= out_dim) return;
  const float* wr = w + static_cast<long long>(o) * in_dim;
  float acc[kProjRows];
#pragma unroll
  for (int r = 0; r < kProjRows; ++r) acc[r] = 0.f;
  for (int i = lane; i < in_dim; i += 32) {
    const float wv = wr[i];
#pragma unroll
    for (int r = 0; r < kProjRows; ++r) acc[r] = fmaf(wv, es[r * in_dim + i], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kProjRows; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float bo = b ? b[o] : 0.f;
#pragma unroll
    for (int r = 0; r < kProjRows; ++r)
      if (r0 + r < rows) {
        const float y = acc[r] + bo;
        out[static_cast<long long>(r0 + r) * out_dim + o] = act ? silu_precise(y) : y;
      }
  }
}

int launch_linear(const float* in, const int64_t* t, const float* freq, int rows, int in_dim, const float* w,
                         const float* b, int out_dim, int act, float* out, cudaStream_t st, const char* what) {
  const size_t smem = sizeof(float) * kProjRows * static_cast<size_t>(in_dim);
  DMME_REQUIRE(smem <= 48 * 1024, DMME_E_SHAPE, "%s: input width %d too large", what, in_dim);
  dim3 grid(ceil_div(out_dim, 8), ceil_div(rows, kProjRows));
  cudaError_t le = launch_pdl(temb_linear_kernel, grid, dim3(256), smem, st, in, t, freq, rows, in_dim, w, b, out_dim, act, out);
  if (le != cudaSuccess) return check_launch_err(le, "temb_linear_kernel");
  return check_launch(what);
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_temb_mlp_fwd(const int64_t* t, int rows, const float* freq, int half, const float* w1,
                                 const float* b1, const float* w2, const float* b2, int emb_dim, float* scratch,
                                 float* emb_out, void* stream) {
  DMME_REQUIRE(t && freq && w1 && b1 && w2 && b2 && scratch && emb_out, DMME_E_BADARG, "temb_mlp: null pointer");
  DMME_REQUIRE(rows > 0 && half > 0 && emb_dim > 0, DMME_E_BADARG, "temb_mlp: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = launch_linear(nullptr, t, freq, rows, 2 * half, w1, b1, emb_dim, 1, scratch, st, "temb_mlp(layer 1)");
  if (rc) return rc;
  return launch_linear(scratch, nullptr, nullptr, rows, emb_dim, w2, b2, emb_dim, 1, emb_out, st, "temb_mlp(layer 2)");
}

extern "C" int dmme_temb_proj_fwd(const float* emb, int rows, int emb_dim, const float* wcat, const float* bcat,
                                  int total, float* out, void* stream) {
  DMME_REQUIRE(emb && wcat && out, DMME_E_BADARG, "temb_proj: null pointer");
  DMME_REQUIRE(rows > 0 && emb_dim > 0 && total > 0, DMME_E_BADARG, "temb_proj: bad sizes");
  return launch_linear(emb, nullptr, nullptr, rows, emb_dim, wcat, bcat, total, 0, out, static_cast<cudaStream_t>(stream),
                       "temb_proj");
}
