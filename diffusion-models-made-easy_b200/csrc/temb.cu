// Timestep conditioning, all fp32 (this part of the network must not be rounded to bf16).
//   temb_mlp:  sinusoidal embedding -> Linear -> SiLU -> Linear -> SiLU   (models/ddpm.py:211-217, 338-349)
//   temb_proj: every ResBlock.condition Linear batched into one [total][emb] GEMV/GEMM (models/ddpm.py:101-104)
#include "common.cuh"

namespace dmme {

__global__ void __launch_bounds__(256) temb_mlp_kernel(const int64_t* __restrict__ t, const float* __restrict__ freq,
                                                       int half, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, int emb, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* e0 = sm;             // [2 * half]
  float* h1 = sm + 2 * half;  // [emb]
  const int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float tv = static_cast<float>(t[row]);  // int64 * float32 promotes to float32 in the reference
  for (int j = tid; j < half; j += blockDim.x) {
    const float a = tv * freq[j];
    e0[j] = sinf(a);
    e0[half + j] = cosf(a);
  }
  __syncthreads();
  const int pos = 2 * half;
  for (int o = warp; o < emb; o += nw) {
    const float* wr = w1 + static_cast<long long>(o) * pos;
    float s = 0.f;
    for (int i = lane; i < pos; i += 32) s = fmaf(wr[i], e0[i], s);
    s = warp_sum(s);
    if (lane == 0) h1[o] = silu_precise(s + b1[o]);
  }
  __syncthreads();
  for (int o = warp; o < emb; o += nw) {
    const float* wr = w2 + static_cast<long long>(o) * emb;
    float s = 0.f;
    for (int i = lane; i < emb; i += 32) s = fmaf(wr[i], h1[i], s);
    s = warp_sum(s);
    if (lane == 0) out[static_cast<long long>(row) * emb + o] = silu_precise(s + b2[o]);
  }
}

constexpr int kProjRows = 8;

__global__ void __launch_bounds__(256) temb_proj_kernel(const float* __restrict__ emb, int rows, int emb_dim,
                                                        const float* __restrict__ wcat,
                                                        const float* __restrict__ bcat, int total,
                                                        float* __restrict__ out) {
  extern __shared__ float es[];  // [kProjRows][emb_dim]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.y * kProjRows;
  for (int idx = tid; idx < kProjRows * emb_dim; idx += 256) {
    const int r = idx / emb_dim, i = idx - r * emb_dim;
    es[idx] = (r0 + r < rows) ? emb[static_cast<long long>(r0 + r) * emb_dim + i] : 0.f;
  }
  __syncthreads();
  const int o = blockIdx.x * 8 + warp;
  if (o >= total) return;
  const float* wr = wcat + static_cast<long long>(o) * emb_dim;
  float acc[kProjRows];
#pragma unroll
  for (int r = 0; r < kProjRows; ++r) acc[r] = 0.f;
  for (int i = lane; i < emb_dim; i += 32) {
    const float w = wr[i];
#pragma unroll
    for (int r = 0; r < kProjRows; ++r) acc[r] = fmaf(w, es[r * emb_dim + i], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kProjRows; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float bo = bcat ? bcat[o] : 0.f;
#pragma unroll
    for (int r = 0; r < kProjRows; ++r)
      if (r0 + r < rows) out[static_cast<long long>(r0 + r) * total + o] = acc[r] + bo;
  }
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_temb_mlp_fwd(const int64_t* t, int rows, const float* freq, int half, const float* w1,
                                 const float* b1, const float* w2, const float* b2, int emb_dim, float* emb_out,
                                 void* stream) {
  DMME_REQUIRE(t && freq && w1 && b1 && w2 && b2 && emb_out, DMME_E_BADARG, "temb_mlp: null pointer");
  DMME_REQUIRE(rows > 0 && half > 0 && emb_dim > 0, DMME_E_BADARG, "temb_mlp: bad sizes");
  const size_t smem = sizeof(float) * (2 * static_cast<size_t>(half) + emb_dim);
  DMME_REQUIRE(smem <= 48 * 1024, DMME_E_SHAPE, "temb_mlp: pos_dim + emb_dim too large");
  temb_mlp_kernel<<<rows, 256, smem, static_cast<cudaStream_t>(stream)>>>(t, freq, half, w1, b1, w2, b2, emb_dim, emb_out);
  return check_launch("temb_mlp_kernel");
}

extern "C" int dmme_temb_proj_fwd(const float* emb, int rows, int emb_dim, const float* wcat, const float* bcat,
                                  int total, float* out, void* stream) {
  DMME_REQUIRE(emb && wcat && out, DMME_E_BADARG, "temb_proj: null pointer");
  DMME_REQUIRE(rows > 0 && emb_dim > 0 && total > 0, DMME_E_BADARG, "temb_proj: bad sizes");
  const size_t smem = sizeof(float) * kProjRows * static_cast<size_t>(emb_dim);
  DMME_REQUIRE(smem <= 48 * 1024, DMME_E_SHAPE, "temb_proj: emb_dim too large");
  dim3 grid(ceil_div(total, 8), ceil_div(rows, kProjRows));
  temb_proj_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(emb, rows, emb_dim, wcat, bcat, total, out);
  return check_launch("temb_proj_kernel");
}
