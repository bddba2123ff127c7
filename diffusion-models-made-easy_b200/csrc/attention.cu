// Self-attention core  out = softmax(scale * Q K^T) V  for the UNet attention blocks
// (Attention.forward_attention models/ddpm.py:54-63, MultiHeadAttention models/iddpm.py:36-47).
//
// Generic CUDA-core kernel: any head count / head dim / sequence length / storage type, fp32 math.
// One CTA handles 8 query rows of one (batch, head); K and V stream through shared memory in
// 32-key tiles, the full score row lives in shared memory so softmax is exact (no online rescale).
#include "common.cuh"

namespace dmme {

struct AttnParams {
  const void* q; const void* k; const void* v;
  long long batch_stride; int row_stride, head_stride;
  int v_transposed; long long v_batch_stride;
  int n, heads, L, dh;
  float scale;
  int swap;
  void* out;
};

constexpr int kAttnRows = 8;
constexpr int kAttnTile = 32;

template <typename T>
__global__ void __launch_bounds__(256) attn_generic_kernel(const AttnParams p) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  const int dh = p.dh, L = p.L;
  float* qs = sm;                                // [8][dh]
  float* tile = qs + kAttnRows * dh;             // [32][dh + 1]
  float* sc = tile + kAttnTile * (dh + 1);       // [8][L]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int row = blockIdx.x * kAttnRows + warp;
  const bool row_ok = row < L;

  const T* qb = static_cast<const T*>(p.q) + b * p.batch_stride + static_cast<long long>(h) * p.head_stride;
  const T* kb = static_cast<const T*>(p.k) + b * p.batch_stride + static_cast<long long>(h) * p.head_stride;
  const T* vb = static_cast<const T*>(p.v);

  for (int c = lane; c < dh; c += 32)
    qs[warp * dh + c] = row_ok ? ld_act<T>(qb + static_cast<long long>(row) * p.row_stride + c) : 0.f;

  // ---- scores ----
  for (int j0 = 0; j0 < L; j0 += kAttnTile) {
    __syncthreads();
    for (int idx = tid; idx < kAttnTile * dh; idx += 256) {
      const int j = idx / dh, c = idx - j * dh;
      tile[j * (dh + 1) + c] = (j0 + j < L) ? ld_act<T>(kb + static_cast<long long>(j0 + j) * p.row_stride + c) : 0.f;
    }
    __syncthreads();
    float s = 0.f;
    const float* kr = tile + lane * (dh + 1);
    const float* qr = qs + warp * dh;
    for (int c = 0; c < dh; ++c) s = fmaf(qr[c], kr[c], s);
    if (j0 + lane < L) sc[warp * L + j0 + lane] = s * p.scale;
  }
  __syncwarp();
  // ---- softmax over the row (one warp per row) ----
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) mx = fmaxf(mx, sc[warp * L + j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = expf(sc[warp * L + j] - mx);
    sc[warp * L + j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;

  // ---- P V ----
  constexpr int kMaxCPerLane = 16;  // dh <= 512 (the LSUN-256 UNet attends over 512 channels, configs/ddpm/lsun_bedroom.yaml:82-90)
  float acc[kMaxCPerLane];
#pragma unroll
  for (int i = 0; i < kMaxCPerLane; ++i) acc[i] = 0.f;
  for (int j0 = 0; j0 < L; j0 += kAttnTile) {
    __syncthreads();
    if (!p.v_transposed) {
      const T* vh = vb + b * p.batch_stride + static_cast<long long>(h) * p.head_stride;
      for (int idx = tid; idx < kAttnTile * dh; idx += 256) {
        const int j = idx / dh, c = idx - j * dh;
        tile[j * (dh + 1) + c] = (j0 + j < L) ? ld_act<T>(vh + static_cast<long long>(j0 + j) * p.row_stride + c) : 0.f;
      }
    } else {
      const T* vh = vb + b * p.v_batch_stride + static_cast<long long>(h) * dh * L;
      for (int idx = tid; idx < kAttnTile * dh; idx += 256) {
        const int c = idx / kAttnTile, j = idx - c * kAttnTile;
        tile[j * (dh + 1) + c] = (j0 + j < L) ? ld_act<T>(vh + static_cast<long long>(c) * L + j0 + j) : 0.f;
      }
    }
    __syncthreads();
    const int jn = (L - j0) < kAttnTile ? (L - j0) : kAttnTile;
    for (int j = 0; j < jn; ++j) {
      const float pj = sc[warp * L + j0 + j];
#pragma unroll
      for (int i = 0; i < kMaxCPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < dh) acc[i] = fmaf(pj, tile[j * (dh + 1) + c], acc[i]);
      }
    }
  }
  if (row_ok) {
    int bo = b, ho = h;
    if (p.swap) {
      const int flat = b * p.heads + h;  // "(b head)" index reinterpreted as "(head b)"
      bo = flat % p.n;
      ho = flat / p.n;
    }
    T* o = static_cast<T*>(p.out) + (static_cast<long long>(bo) * L + row) * (p.heads * dh) + ho * dh;
#pragma unroll
    for (int i = 0; i < kMaxCPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < dh) st_act<T>(o + c, acc[i] * inv);
    }
  }
}

bool attn_tc_supported(int act_dtype, int heads, int L, int dh, int row_stride, long long batch_stride,
                       int v_transposed, long long v_batch_stride, int swap);
int attn_tc_forward(const void* q, const void* k, const void* vt, int n, int d, float scale, void* out,
                    cudaStream_t stream);
bool attn_mma_supported(int act_dtype, int heads, int L, int dh, int row_stride, int head_stride, long long batch_stride,
                        int v_transposed, const void* q, const void* k, const void* v, const void* out);
bool attn_tensor_core_multi_head_enabled();
bool attn_tc_mh_supported(int act_dtype, int heads, int L, int dh, int row_stride, int head_stride, long long batch_stride,
                          int v_transposed, const void* q, const void* k, const void* v, const void* out);
int attn_tc_mh_forward(const void* qkv, int n, int heads, int L, int dh, float scale, int swap, void* out, float* p_out,
                       cudaStream_t stream);
int attn_mma_forward(const void* q, const void* k, const void* v, long long batch_stride, int row_stride, int head_stride,
                     int n, int heads, int L, int dh, float scale, int swap, void* out, float* p_out, cudaStream_t stream);

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_attention_uses_tc(long long batch_stride, int row_stride, int v_transposed,
                                      long long v_batch_stride, int heads, int L, int dh, int head_batch_swap,
                                      int act_dtype) {
  return attn_tc_supported(act_dtype, heads, L, dh, row_stride, batch_stride, v_transposed, v_batch_stride,
                           head_batch_swap) ? 1 : 0;
}

extern "C" int dmme_attention_fwd(const void* q, const void* k, const void* v, long long batch_stride,
                                  int row_stride, int head_stride, int v_transposed, long long v_batch_stride, int n,
                                  int heads, int L, int dh, float scale, int head_batch_swap, void* out,
                                  int act_dtype, int kernel, void* stream) {
  DMME_REQUIRE(q && k && v && out, DMME_E_BADARG, "attention: null pointer");
  DMME_REQUIRE(n > 0 && heads > 0 && L > 0 && dh > 0, DMME_E_BADARG, "attention: bad sizes");
  const bool tc_ok = attn_tc_supported(act_dtype, heads, L, dh, row_stride, batch_stride, v_transposed,
                                       v_batch_stride, head_batch_swap);
  DMME_REQUIRE(kernel != DMME_CONV_TC || tc_ok, DMME_E_SHAPE, "attention: shape/layout not eligible for the tcgen05 kernel");
  if (tc_ok && kernel != DMME_CONV_GENERIC)
    return attn_tc_forward(q, k, v, n, dh, scale, out, static_cast<cudaStream_t>(stream));
  // packed multi-head qkv at 256 tokens / 64- or 32-channel heads (the IDDPM UNet's 16x16 attention sites): tcgen05
  if (kernel != DMME_CONV_GENERIC && attn_tensor_core_multi_head_enabled() &&
      attn_tc_mh_supported(act_dtype, heads, L, dh, row_stride, head_stride, batch_stride, v_transposed, q, k, v, out))
    return attn_tc_mh_forward(q, n, heads, L, dh, scale, head_batch_swap, out, nullptr, static_cast<cudaStream_t>(stream));
  if (kernel != DMME_CONV_GENERIC &&
      attn_mma_supported(act_dtype, heads, L, dh, row_stride, head_stride, batch_stride, v_transposed, q, k, v, out))
    return attn_mma_forward(q, k, v, batch_stride, row_stride, head_stride, n, heads, L, dh, scale, head_batch_swap, out,
                            nullptr, static_cast<cudaStream_t>(stream));
  DMME_REQUIRE(dh <= 512, DMME_E_SHAPE, "attention: head dim %d > 512 not supported", dh);
  const size_t smem = sizeof(float) * (static_cast<size_t>(kAttnRows) * dh + kAttnTile * (dh + 1) +
                                        static_cast<size_t>(kAttnRows) * L);
  DMME_REQUIRE(smem <= 200 * 1024, DMME_E_SHAPE, "attention: sequence length %d too long for the generic kernel", L);
  AttnParams p;
  p.q = q; p.k = k; p.v = v; p.batch_stride = batch_stride; p.row_stride = row_stride; p.head_stride = head_stride;
  p.v_transposed = v_transposed; p.v_batch_stride = v_batch_stride;
  p.n = n; p.heads = heads; p.L = L; p.dh = dh; p.scale = scale; p.swap = head_batch_swap; p.out = out;
  dim3 grid(ceil_div(L, kAttnRows), heads, n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (act_dtype == DMME_BF16) {
    e = cudaFuncSetAttribute(attn_generic_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    e = launch_pdl(attn_generic_kernel<__nv_bfloat16>, grid, dim3(256), smem, st, p);
    if (e != cudaSuccess) return check_launch_err(e, "attn_generic_kernel");
  } else {
    e = cudaFuncSetAttribute(attn_generic_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    e = launch_pdl(attn_generic_kernel<float>, grid, dim3(256), smem, st, p);
    if (e != cudaSuccess) return check_launch_err(e, "attn_generic_kernel");
  }
  return check_launch("attn_generic_kernel");
}
