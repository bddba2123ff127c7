// Multi-head self-attention core on the tensor cores for the IDDPM flavour (MultiHeadAttention, models/iddpm.py:16-59:
// 4 heads of 64 or 32 channels over 256 / 64 / 16 tokens, channels laid out [head][q | k | v][dh], and the reference's
// "(b head)" -> "(head b)" regrouping of the output) and for any other bf16 shape with dh % 16 == 0, L % 16 == 0, L <= 256.
//
// One CTA = 64 query rows of one (image, head), four warps of 16 rows each:
//   S = Q K^T     mma.sync (wmma 16x16x16, bf16 operands, fp32 accumulate) over 64-key chunks -> fp32 rows in shared memory
//   P = softmax   exact two-pass row softmax by the owning warp, P rounded to bf16 IN PLACE over the fp32 row
//   O = P V       mma.sync over 64-key chunks, accumulators in registers, scaled by 1 / rowsum on the way out
// The CUDA-core kernel this replaces (attn_generic_kernel: 8 rows per CTA, scalar FMAs) ran the IDDPM UNet's 11 attention
// sites at 3.6 TFLOP/s -- 10 ms of a batch-128 forward.  The single-head 256-token blocks of the DDPM flavour keep their
// tcgen05 kernel (attention_tc.cu); head dims of 32 / 64 are below that kernel's 64-channel K chunks and 128-row tiles.
#include <mma.h>

#include "common.cuh"

namespace dmme {

struct AttnMmaParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;
  long long batch_stride; int row_stride, head_stride;
  int n, heads, L, dh;
  float scale;
  int swap;
  __nv_bfloat16* out;
  float* p_out;  // optional fp32 [n * heads][L][L]: the normalised softmax matrix, kept for the backward pass
};

constexpr int kAmRows = 64;   // query rows per CTA
constexpr int kAmKeys = 64;   // keys per chunk

__global__ void __launch_bounds__(128) attn_mma_kernel(const AttnMmaParams p) {
  using namespace nvcuda;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  pdl_trigger();
  pdl_wait();
  const int dh = p.dh, L = p.L;
  const int ldq = dh + 8;           // bf16 elements
  const int lds = L + 8;            // fp32 elements per score row
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* KVs = Qs + kAmRows * ldq;
  float* S = reinterpret_cast<float*>(KVs + kAmKeys * ldq);
  float* rowinv = S + kAmRows * lds;
  float* Os = reinterpret_cast<float*>(sm_raw);  // [64][dh + 4], reuses Qs | KVs after the last product

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int row0 = blockIdx.x * kAmRows;
  const long long base = b * p.batch_stride + static_cast<long long>(h) * p.head_stride;
  const __nv_bfloat16* qb = p.q + base;
  const __nv_bfloat16* kb = p.k + base;
  const __nv_bfloat16* vb = p.v + base;
  const int vec = dh >> 3;  // 16-byte vectors per row

  // rows [r0, r0 + 64) of a [L][dh] operand (row stride p.row_stride) -> dst [64][ldq], zero beyond L
  auto load_rows = [&](const __nv_bfloat16* src, int r0, __nv_bfloat16* dst) {
    for (int idx = tid; idx < kAmRows * vec; idx += 128) {
      const int r = idx / vec, c = (idx - r * vec) << 3;
      uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
      if (r0 + r < L) v4 = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r0 + r) * p.row_stride + c));
      *reinterpret_cast<uint4*>(dst + r * ldq + c) = v4;
    }
  };

  load_rows(qb, row0, Qs);
  const int wr = warp * 16;  // this warp's rows inside the CTA tile
  const bool warp_live = row0 + wr < L;

  // ---- S = Q K^T ----
  for (int k0 = 0; k0 < L; k0 += kAmKeys) {
    __syncthreads();  // previous chunk consumed (and, first time, Qs visible)
    load_rows(kb, k0, KVs);
    __syncthreads();
    if (warp_live) {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) wmma::fill_fragment(acc[f], 0.f);
      for (int kk = 0; kk < dh; kk += 16) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fa;
        wmma::load_matrix_sync(fa, Qs + wr * ldq + kk, ldq);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (k0 + f * 16 < L) {
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> fb;  // K^T: (k, key) = KVs[key][k]
            wmma::load_matrix_sync(fb, KVs + f * 16 * ldq + kk, ldq);
            wmma::mma_sync(acc[f], fa, fb, acc[f]);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < 4; ++f)
        if (k0 + f * 16 < L) wmma::store_matrix_sync(S + wr * lds + k0 + f * 16, acc[f], lds, wmma::mem_row_major);
    }
  }
  __syncwarp();

  // ---- exact row softmax by the owning warp; P (bf16) overwrites the start of its own fp32 row ----
  if (warp_live) {
    for (int r = 0; r < 16; ++r) {
      float* srow = S + (wr + r) * lds;
      float t[8];  // L <= 256
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = lane + 32 * i;
        t[i] = j < L ? srow[j] * p.scale : -INFINITY;
        mx = fmaxf(mx, t[i]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t[i] = (lane + 32 * i < L) ? __expf(t[i] - mx) : 0.f;
        sum += t[i];
      }
      sum = warp_sum(sum);
      __syncwarp();  // every lane has read its fp32 values before the row is overwritten
      __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(srow);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = lane + 32 * i;
        if (j < L) prow[j] = __float2bfloat16_rn(t[i]);
      }
      if (lane == 0) rowinv[wr + r] = 1.0f / sum;
      if (p.p_out) {
        const float inv = 1.0f / sum;
        float* prow_out = p.p_out + ((static_cast<long long>(b) * p.heads + h) * L + row0 + wr + r) * L;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = lane + 32 * i;
          if (j < L) prow_out[j] = t[i] * inv;
        }
      }
    }
  }

  // ---- O = P V ----
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> oacc[8];  // dh <= 128
#pragma unroll
  for (int f = 0; f < 8; ++f) wmma::fill_fragment(oacc[f], 0.f);
  const int nf = dh >> 4;
  for (int k0 = 0; k0 < L; k0 += kAmKeys) {
    __syncthreads();
    load_rows(vb, k0, KVs);
    __syncthreads();
    if (warp_live) {
      for (int kk = 0; kk < kAmKeys && k0 + kk < L; kk += 16) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fa;
        wmma::load_matrix_sync(fa, reinterpret_cast<const __nv_bfloat16*>(S + wr * lds) + k0 + kk, 2 * lds);
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          if (f < nf) {
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;
            wmma::load_matrix_sync(fb, KVs + kk * ldq + f * 16, ldq);
            wmma::mma_sync(oacc[f], fa, fb, oacc[f]);
          }
        }
      }
    }
  }
  __syncthreads();  // Qs / KVs are dead: Os may overwrite them
  const int ldo = dh + 4;
  if (warp_live) {
#pragma unroll
    for (int f = 0; f < 8; ++f)
      if (f < nf) wmma::store_matrix_sync(Os + wr * ldo + f * 16, oacc[f], ldo, wmma::mem_row_major);
    __syncwarp();
    int bo = b, ho = h;
    if (p.swap) {
      const int flat = b * p.heads + h;  // "(b head)" index reinterpreted as "(head b)" (models/iddpm.py:44-46)
      bo = flat % p.n;
      ho = flat / p.n;
    }
    for (int idx = lane; idx < 16 * vec; idx += 32) {
      const int r = idx / vec, c = (idx - r * vec) << 3;
      const int row = row0 + wr + r;
      if (row >= L) continue;
      const float inv = rowinv[wr + r];
      const float* o = Os + (wr + r) * ldo + c;
      uint4 v4;
      v4.x = pack_bf16x2(o[0] * inv, o[1] * inv); v4.y = pack_bf16x2(o[2] * inv, o[3] * inv);
      v4.z = pack_bf16x2(o[4] * inv, o[5] * inv); v4.w = pack_bf16x2(o[6] * inv, o[7] * inv);
      *reinterpret_cast<uint4*>(p.out + (static_cast<long long>(bo) * L + row) * (p.heads * dh) + ho * dh + c) = v4;
    }
  }
}

static int g_attn_mma_mode = 1;
bool attn_tensor_core_multi_head_enabled() { return g_attn_mma_mode != 0; }

bool attn_mma_supported(int act_dtype, int heads, int L, int dh, int row_stride, int head_stride, long long batch_stride,
                        int v_transposed, const void* q, const void* k, const void* v, const void* out) {
  if (g_attn_mma_mode == 0) return false;
  if (act_dtype != DMME_BF16 || v_transposed) return false;
  if (dh % 16 || dh < 16 || dh > 128 || L % 16 || L < 16 || L > 256) return false;
  if (row_stride % 8 || head_stride % 8 || batch_stride % 8 || (heads * dh) % 8) return false;
  auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  return al(q) && al(k) && al(v) && al(out);
}

int attn_mma_forward(const void* q, const void* k, const void* v, long long batch_stride, int row_stride, int head_stride,
                     int n, int heads, int L, int dh, float scale, int swap, void* out, float* p_out, cudaStream_t stream) {
  AttnMmaParams p;
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.batch_stride = batch_stride; p.row_stride = row_stride; p.head_stride = head_stride;
  p.n = n; p.heads = heads; p.L = L; p.dh = dh; p.scale = scale; p.swap = swap;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.p_out = p_out;
  const size_t smem = static_cast<size_t>(2) * kAmRows * (dh + 8) * 2 + static_cast<size_t>(kAmRows) * (L + 8) * 4 +
                      kAmRows * 4 + 128;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) { set_error("attn_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid(ceil_div(L, kAmRows), heads, n);
  return check_launch_err(launch_pdl(attn_mma_kernel, grid, dim3(128), smem, stream, p), "attn_mma_kernel");
}

}  // namespace dmme

// A/B measurement switch: 0 = multi-head attention stays on the CUDA-core kernel, 1 = mma.sync kernel (default)
extern "C" void dmme_set_attn_mma_mode(int mode) { dmme::g_attn_mma_mode = mode; }
