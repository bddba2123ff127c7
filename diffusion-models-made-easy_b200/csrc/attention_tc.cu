// Fused single-head self-attention core on tcgen05 for the 16x16-resolution blocks (L = 256 tokens):
//
//   S = Q K^T   (128 query rows x 256 keys, fp32 in TMEM columns [0,256))
//   P = exp2((S - rowmax) * scale * log2e)      (fp32 row statistics, P rounded to bf16 into shared memory)
//   O = P V     (128 x d, fp32 in TMEM columns [256, 256 + d)),   out = O / rowsum
//
// One CTA per (image, 128-query block).  Q/K chunks of 64 channels and V^T chunks of 64 keys stream through
// a two-slot TMA ring; the score matrix never leaves the SM.  Operands are all K-major SWIZZLE_128B tiles:
// Q,K from the [n][L][C] tensors and V from the transposed [n][C][L] tensor the qkv conv epilogue writes.
// Replaces torch.bmm / F.softmax / torch.bmm of Attention.forward_attention (models/ddpm.py:58-61).
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct AttnTcParams {
  CUtensorMap q, k, vt;
  int n, d;
  float scale_log2e;
  __nv_bfloat16* out;  // [n][256][d]
};

constexpr int kSeq = 256;
constexpr int kAttnSmWarps = 8;            // softmax / epilogue warps: two per TMEM lane quarter, each half of the columns
constexpr int kAttnThreads = (2 + kAttnSmWarps) * 32;
constexpr int kStageA = 128 * 128;        // 128 rows x 64 bf16
constexpr int kStageB = 256 * 128;        // up to 256 rows x 64 bf16
constexpr int kAttnStage = kStageA + kStageB;
constexpr int kAttnStages = 3;           // 3 x 48 KB ring + 64 KB P = 209 KB: the Q/K phase is load-latency bound with two
constexpr int kPBytes = 128 * kSeq * 2;   // P: 4 chunks of [128][64] bf16
constexpr int kAttnSmem = kAttnStages * kAttnStage + kPBytes + 1024;

__global__ void __launch_bounds__(kAttnThreads, 1) attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kAttnStages];
  __shared__ __align__(8) uint64_t empty_bar[kAttnStages];
  __shared__ __align__(8) uint64_t s_full, p_ready, o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float row_part[2][128];  // per-row partial max, then partial sum, of the two column halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* ring = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* pbuf = ring + kAttnStages * kAttnStage;

  const int q0 = blockIdx.x * 128;
  const int img = blockIdx.y;
  const int d = p.d;
  const int kch = d / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&s_full, 1);
    mbar_init(&p_ready, kAttnSmWarps * 32);
    mbar_init(&o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q);
    tma_prefetch_desc(&p.k);
    tma_prefetch_desc(&p.vt);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;
  const uint32_t tmem_o = tmem_slot + 256;
  pdl_trigger();  // after the TMEM allocation (see ptx_sm100.cuh)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int kc = 0; kc < kch; ++kc, ++it) {
        const int s = it % kAttnStages;
        mbar_wait(&empty_bar[s], ((it / kAttnStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], kStageA + kStageB);
        tma_load_3d(ring + s * kAttnStage, &p.q, &full_bar[s], kc * 64, q0, img);
        tma_load_3d(ring + s * kAttnStage + kStageA, &p.k, &full_bar[s], kc * 64, 0, img);
      }
      for (int jc = 0; jc < kSeq / 64; ++jc, ++it) {
        const int s = it % kAttnStages;
        mbar_wait(&empty_bar[s], ((it / kAttnStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], d * 128);
        tma_load_3d(ring + s * kAttnStage + kStageA, &p.vt, &full_bar[s], jc * 64, 0, img);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, kSeq);
      const uint32_t idesc_o = umma_idesc_bf16(128, d);
      int it = 0;
      for (int kc = 0; kc < kch; ++kc, ++it) {
        const int s = it % kAttnStages;
        mbar_wait(&full_bar[s], (it / kAttnStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + s * kAttnStage);
        const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + kStageA);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, (kc | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&s_full);
      mbar_wait(&p_ready, 0);
      tc_fence_after();
      for (int jc = 0; jc < kSeq / 64; ++jc, ++it) {
        const int s = it % kAttnStages;
        mbar_wait(&full_bar[s], (it / kAttnStages) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(pbuf + jc * kStageA));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + s * kAttnStage + kStageA));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, adesc + 2 * k, bdesc + 2 * k, idesc_o, (jc | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&o_full);
    }
  } else {
    // thread = (query row, half of the key columns): the two warps of a TMEM lane quarter split the 256 scores of a row,
    // exchange their partial row maximum / sum through shared memory, and later split the d output columns the same way
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    constexpr int kHalfSeq = kSeq / 2;
    const int c_lo = half * kHalfSeq;
    mbar_wait(&s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = c_lo; c < c_lo + kHalfSeq; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_off + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    row_part[half][row] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
    mx = fmaxf(mx, row_part[half ^ 1][row]);
    asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");  // both halves have read before the sums overwrite
    float sum = 0.f;
    const float sl = p.scale_log2e;
    const float mxs = mx * sl;
    uint8_t* prow = pbuf + row * 128;
#pragma unroll 1
    for (int c = c_lo; c < c_lo + kHalfSeq; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_off + c, v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        e[j] = exp2f(fmaf(__uint_as_float(v[j]), sl, -mxs));
        sum += e[j];
      }
      uint8_t* pc = prow + (c >> 6) * kStageA;
      const int u0 = (c & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(e[8 * jj + 0], e[8 * jj + 1]);
        o.y = pack_bf16x2(e[8 * jj + 2], e[8 * jj + 3]);
        o.z = pack_bf16x2(e[8 * jj + 4], e[8 * jj + 5]);
        o.w = pack_bf16x2(e[8 * jj + 6], e[8 * jj + 7]);
        *reinterpret_cast<uint4*>(pc + (((u0 + jj) ^ (row & 7)) << 4)) = o;
      }
    }
    row_part[half][row] = sum;
    tc_fence_before();
    fence_proxy_async();  // P was written through the generic proxy; the MMA reads it through the async proxy
    mbar_arrive(&p_ready);
    asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
    // the same summation order in both halves: the two threads of a row scale by the same bits
    sum = row_part[0][row] + row_part[1][row];

    mbar_wait(&o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(img) * kSeq + q0 + row) * d;
    const int dh2 = d >> 1;
#pragma unroll 1
    for (int c = half * dh2; c < (half + 1) * dh2; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_off + c, v);
      tmem_ld_wait();
      uint4* dp = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[8 * jj + 0]) * inv, __uint_as_float(v[8 * jj + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[8 * jj + 2]) * inv, __uint_as_float(v[8 * jj + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[8 * jj + 4]) * inv, __uint_as_float(v[8 * jj + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[8 * jj + 6]) * inv, __uint_as_float(v[8 * jj + 7]) * inv);
        dp[jj] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

// ----------------------------------------------------------------------------------------------
// Multi-head variant for the IDDPM flavour (MultiHeadAttention, models/iddpm.py:16-59) at 256 tokens and 64-channel heads:
// q | k | v of head h are the channel ranges [3 dh h, +dh), [+dh, +2 dh), [+2 dh, +3 dh) of one packed [n][L][3C] tensor, so
// all three operands come from the same tensor map with different boxes.  One CTA per (image, head, 128 queries):
//   S = Q K^T   one 64-channel k-block, four M=128 N=256 K=16 MMAs, both operands K-major
//   P           the single-head kernel's softmax (eight warps), bf16 into shared memory
//   O = P V     V is NOT transposed here: a [64 keys][64 channels] box is the canonical MN-major SWIZZLE_128B layout of the
//               B operand (N = channels contiguous, K = keys in rows), as in conv_wgrad_tc.cu -- K step = 16 rows = 2 KB
// and the output goes to the reference's "(b head) -> (head b)" position (models/iddpm.py:44-46).
// ----------------------------------------------------------------------------------------------
struct AttnTcMhParams {
  CUtensorMap q, k, v;  // the packed qkv tensor, boxes [64 ch][128 rows], [64][256], [64][64]
  int n, heads;
  float scale_log2e;
  int swap;
  __nv_bfloat16* out;   // [n][256][heads * 64]
  float* p_out;         // optional fp32 [n * heads][256][256]: the normalised softmax matrix, kept for the backward pass
};

// P V instruction descriptor: A (P) K-major, B (V) MN-major, M = 128 queries, N = head channels
__host__ __device__ constexpr uint32_t attn_mh_idesc_o(int dh) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | (uint32_t(dh >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}

__device__ __forceinline__ uint64_t attn_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t((8192u >> 4) & 0x3FFFu) << 16;   // LBO: next 64-channel block (N = 64: a single block, never used)
  d |= uint64_t(1024 >> 4) << 32;                // SBO: 8 key rows * 128 B
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// DH = 64: boxes q [64][128 rows], k [64][256], v [64][64 keys].  DH = 32: a 64-channel box at the head's first channel holds
// [q | k] and one at +32 holds [k | v]: Q and K are the two 64-byte halves of the rows of ONE [64][256] box (K-major
// descriptors step 32 bytes inside a 128-byte row anyway), V is the second half of the rows of the [k | v] boxes.
template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_tc_mh_kernel(const __grid_constant__ AttnTcMhParams p) {
  constexpr int kMhDh = DH;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qk_full, v_full[kSeq / 64];
  __shared__ __align__(8) uint64_t s_full, p_ready, o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float row_part[2][128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* qbuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // [128][64] bf16
  uint8_t* kbuf = qbuf + kStageA;                                       // [256][64]
  // Two CTAs share an SM (their serial phases overlap): 97 KB of shared memory and 256 TMEM columns each.  P (64 KB)
  // overwrites Q | K (48 KB, dead once S is complete) and 16 KB more; O reuses the first columns of S (dead once every
  // softmax thread has read its row, which is what p_ready says).
  uint8_t* pbuf = qbuf;                                                 // P: 4 chunks of [128][64]
  uint8_t* vbuf = qbuf + kPBytes;                                       // 4 x [64 keys][64 ch]

  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y, img = blockIdx.z;
  const int ch0 = head * 3 * kMhDh;

  if (threadIdx.x == 0) {
    mbar_init(&qk_full, 1);
    for (int j = 0; j < kSeq / 64; ++j) mbar_init(&v_full[j], 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_ready, kAttnSmWarps * 32);
    mbar_init(&o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q);
    tma_prefetch_desc(&p.k);
    tma_prefetch_desc(&p.v);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;
  const uint32_t tmem_o = tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // every operand of the tile fits shared memory at once: all loads are issued up front
      if constexpr (DH == 64) {
        mbar_expect_tx(&qk_full, kStageA + kStageB);
        tma_load_3d(qbuf, &p.q, &qk_full, ch0, q0, img);
        tma_load_3d(kbuf, &p.k, &qk_full, ch0 + kMhDh, 0, img);
      } else {
        mbar_expect_tx(&qk_full, kStageB);
        tma_load_3d(kbuf, &p.k, &qk_full, ch0, 0, img);  // [q | k] of all 256 tokens
      }
      for (int jc = 0; jc < kSeq / 64; ++jc) {
        mbar_expect_tx(&v_full[jc], 64 * 128);
        tma_load_3d(vbuf + jc * 64 * 128, &p.v, &v_full[jc], DH == 64 ? ch0 + 2 * kMhDh : ch0 + kMhDh, jc * 64, img);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, kSeq);
      mbar_wait(&qk_full, 0);
      tc_fence_after();
      // DH = 32: Q = first 64 bytes of rows q0.. of the [q | k] box, K = second 64 bytes of all its rows
      const uint64_t adesc = umma_desc_sw128(DH == 64 ? smem_u32(qbuf) : smem_u32(kbuf) + static_cast<uint32_t>(q0) * 128u);
      const uint64_t bdesc = umma_desc_sw128(DH == 64 ? smem_u32(kbuf) : smem_u32(kbuf) + 64u);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit(&s_full);
      mbar_wait(&p_ready, 0);
      tc_fence_after();
      for (int jc = 0; jc < kSeq / 64; ++jc) {
        mbar_wait(&v_full[jc], 0);
        tc_fence_after();
        const uint64_t pdesc = umma_desc_sw128(smem_u32(pbuf + jc * kStageA));
        const uint32_t vaddr = smem_u32(vbuf + jc * 64 * 128) + (DH == 64 ? 0u : 64u);  // DH = 32: V = second half of [k | v]
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_o, pdesc + 2 * k, attn_desc_mn_sw128(vaddr + k * 2048), attn_mh_idesc_o(DH), (jc | k) != 0 ? 1u : 0u);
      }
      umma_commit(&o_full);
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    constexpr int kHalfSeq = kSeq / 2;
    const int c_lo = half * kHalfSeq;
    mbar_wait(&s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = c_lo; c < c_lo + kHalfSeq; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_off + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    row_part[half][row] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
    mx = fmaxf(mx, row_part[half ^ 1][row]);
    asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
    float sum = 0.f;
    const float sl = p.scale_log2e;
    const float mxs = mx * sl;
    uint8_t* prow = pbuf + row * 128;
#pragma unroll 1
    for (int c = c_lo; c < c_lo + kHalfSeq; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_off + c, v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        e[j] = exp2f(fmaf(__uint_as_float(v[j]), sl, -mxs));
        sum += e[j];
      }
      uint8_t* pc = prow + (c >> 6) * kStageA;
      const int u0 = (c & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(e[8 * jj + 0], e[8 * jj + 1]);
        o.y = pack_bf16x2(e[8 * jj + 2], e[8 * jj + 3]);
        o.z = pack_bf16x2(e[8 * jj + 4], e[8 * jj + 5]);
        o.w = pack_bf16x2(e[8 * jj + 6], e[8 * jj + 7]);
        *reinterpret_cast<uint4*>(pc + (((u0 + jj) ^ (row & 7)) << 4)) = o;
      }
    }
    row_part[half][row] = sum;
    if (p.p_out == nullptr) {
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&p_ready);
      asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
      sum = row_part[0][row] + row_part[1][row];
    } else {
      // training forward: the normalised probabilities go to global memory before the scores are released (O reuses their
      // TMEM columns): a third pass over this thread's half of the row, now that the whole row's sum is known
      asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
      sum = row_part[0][row] + row_part[1][row];
      const float inv_p = 1.0f / sum;
      float* gp = p.p_out + ((static_cast<long long>(img) * p.heads + head) * kSeq + q0 + row) * kSeq;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + kHalfSeq; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = exp2f(fmaf(__uint_as_float(v[j + 0]), sl, -mxs)) * inv_p;
          o.y = exp2f(fmaf(__uint_as_float(v[j + 1]), sl, -mxs)) * inv_p;
          o.z = exp2f(fmaf(__uint_as_float(v[j + 2]), sl, -mxs)) * inv_p;
          o.w = exp2f(fmaf(__uint_as_float(v[j + 3]), sl, -mxs)) * inv_p;
          *reinterpret_cast<float4*>(gp + c + j) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&p_ready);
    }

    mbar_wait(&o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    int bo = img, ho = head;
    if (p.swap) {
      const int flat = img * p.heads + head;  // "(b head)" index reinterpreted as "(head b)" (models/iddpm.py:44-46)
      bo = flat % p.n;
      ho = flat / p.n;
    }
    constexpr int kCols = DH / 2;  // output channels per thread
    __nv_bfloat16* orow = p.out + (static_cast<long long>(bo) * kSeq + q0 + row) * (p.heads * kMhDh) + ho * kMhDh + half * kCols;
    uint32_t v[kCols];
    if constexpr (DH == 64) tmem_ld32(tmem_o + lane_off + half * kCols, v);
    else tmem_ld16(tmem_o + lane_off + half * kCols, v);
    tmem_ld_wait();
    uint4* dp = reinterpret_cast<uint4*>(orow);
#pragma unroll
    for (int jj = 0; jj < kCols / 8; ++jj) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v[8 * jj + 0]) * inv, __uint_as_float(v[8 * jj + 1]) * inv);
      o.y = pack_bf16x2(__uint_as_float(v[8 * jj + 2]) * inv, __uint_as_float(v[8 * jj + 3]) * inv);
      o.z = pack_bf16x2(__uint_as_float(v[8 * jj + 4]) * inv, __uint_as_float(v[8 * jj + 5]) * inv);
      o.w = pack_bf16x2(__uint_as_float(v[8 * jj + 6]) * inv, __uint_as_float(v[8 * jj + 7]) * inv);
      dp[jj] = o;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 256);
  }
}

// ----------------------------------------------------------------------------------------------
// The same for the 64-token sites (8x8 maps, 64-channel heads): one CTA takes TWO images of one head -- 128 query rows and
// 128 key rows that are contiguous in the packed tensor -- computes the 128 x 128 score tile and keeps its two diagonal
// 64 x 64 blocks: the thread of (row, half) owns the keys of image `half`, and where that is not the row's own image it
// contributes nothing to the row maximum / sum and writes zeros into P, so the off-diagonal products vanish in P V.
// ----------------------------------------------------------------------------------------------
struct AttnTcMh64Params {
  CUtensorMap qk, v;   // the packed qkv tensor as [n * 64 rows][3C]: boxes [64 ch][128 rows] and [64 ch][64 rows]
  int n, heads;
  float scale_log2e;
  int swap;
  __nv_bfloat16* out;  // [n][64][heads * 64]
  float* p_out;        // optional fp32 [n * heads][64][64]: the normalised softmax matrix, kept for the backward pass
};

constexpr int kMh64Smem = 2 * kStageA + 2 * 64 * 128 + 1024;  // P (2 chunks, over Q | K) + V (2 chunks)

__global__ void __launch_bounds__(kAttnThreads, 2) attn_tc_mh64_kernel(const __grid_constant__ AttnTcMh64Params p) {
  constexpr int kDh = 64, kL = 64, kKeys = 2 * kL;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qk_full, v_full[2];
  __shared__ __align__(8) uint64_t s_full, p_ready, o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float row_part[2][128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* qbuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // [128][64] bf16
  uint8_t* kbuf = qbuf + kStageA;                                       // [128][64]
  uint8_t* pbuf = qbuf;                                                 // P: 2 chunks of [128][64], over Q | K
  uint8_t* vbuf = qbuf + 2 * kStageA;                                   // 2 x [64 keys][64 ch]

  const int head = blockIdx.y, pair = blockIdx.z;
  const int ch0 = head * 3 * kDh;
  const int row0 = pair * kKeys;  // first token row of the pair in the [n * 64][3C] view

  if (threadIdx.x == 0) {
    mbar_init(&qk_full, 1);
    mbar_init(&v_full[0], 1);
    mbar_init(&v_full[1], 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_ready, kAttnSmWarps * 32);
    mbar_init(&o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qk);
    tma_prefetch_desc(&p.v);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;
  const uint32_t tmem_o = tmem_slot;  // O reuses the score columns once every softmax thread has read its row
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&qk_full, 2 * kStageA);
      tma_load_2d(qbuf, &p.qk, &qk_full, ch0, row0);
      tma_load_2d(kbuf, &p.qk, &qk_full, ch0 + kDh, row0);
      for (int jc = 0; jc < 2; ++jc) {
        mbar_expect_tx(&v_full[jc], 64 * 128);
        tma_load_2d(vbuf + jc * 64 * 128, &p.v, &v_full[jc], ch0 + 2 * kDh, row0 + jc * 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, kKeys);
      mbar_wait(&qk_full, 0);
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(smem_u32(qbuf)), bdesc = umma_desc_sw128(smem_u32(kbuf));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit(&s_full);
      mbar_wait(&p_ready, 0);
      tc_fence_after();
      for (int jc = 0; jc < 2; ++jc) {
        mbar_wait(&v_full[jc], 0);
        tc_fence_after();
        const uint64_t pdesc = umma_desc_sw128(smem_u32(pbuf + jc * kStageA));
        const uint32_t vaddr = smem_u32(vbuf + jc * 64 * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_o, pdesc + 2 * k, attn_desc_mn_sw128(vaddr + k * 2048), attn_mh_idesc_o(kDh), (jc | k) != 0 ? 1u : 0u);
      }
      umma_commit(&o_full);
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;      // key block = image `half` of the pair
    const int row = q * 32 + lane;         // query row: image row >> 6 of the pair, token row & 63
    const bool own = (row >> 6) == half;   // warp-uniform (a warp's 32 rows lie in one image)
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int c_lo = half * kL;
    mbar_wait(&s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
    if (own) {
#pragma unroll 1
      for (int c = c_lo; c < c_lo + kL; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
    float sum = 0.f;
    const float sl = p.scale_log2e;
    const float mxs = mx * sl;
    uint8_t* prow = pbuf + half * kStageA + row * 128;
#pragma unroll 1
    for (int c = c_lo; c < c_lo + kL; c += 32) {
      float e[32];
      if (own) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          e[j] = exp2f(fmaf(__uint_as_float(v[j]), sl, -mxs));
          sum += e[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) e[j] = 0.f;
      }
      const int u0 = (c & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(e[8 * jj + 0], e[8 * jj + 1]);
        o.y = pack_bf16x2(e[8 * jj + 2], e[8 * jj + 3]);
        o.z = pack_bf16x2(e[8 * jj + 4], e[8 * jj + 5]);
        o.w = pack_bf16x2(e[8 * jj + 6], e[8 * jj + 7]);
        *reinterpret_cast<uint4*>(prow + (((u0 + jj) ^ (row & 7)) << 4)) = o;
      }
    }
    row_part[half][row] = sum;
    if (p.p_out != nullptr && own && 2 * pair + (row >> 6) < p.n) {
      // training forward: this thread holds the whole row (its image's 64 keys): normalised probabilities to global memory
      // before the scores are released
      const float inv_p = 1.0f / sum;
      float* gp = p.p_out + ((static_cast<long long>(2 * pair + (row >> 6)) * p.heads + head) * kL + (row & 63)) * kL;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + kL; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = exp2f(fmaf(__uint_as_float(v[j + 0]), sl, -mxs)) * inv_p;
          o.y = exp2f(fmaf(__uint_as_float(v[j + 1]), sl, -mxs)) * inv_p;
          o.z = exp2f(fmaf(__uint_as_float(v[j + 2]), sl, -mxs)) * inv_p;
          o.w = exp2f(fmaf(__uint_as_float(v[j + 3]), sl, -mxs)) * inv_p;
          *reinterpret_cast<float4*>(gp + (c - c_lo) + j) = o;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    mbar_arrive(&p_ready);
    asm volatile("bar.sync 1, %0;" ::"n"(kAttnSmWarps * 32) : "memory");
    sum = row_part[0][row] + row_part[1][row];  // one of the two is the row's sum, the other 0

    mbar_wait(&o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int img = 2 * pair + (row >> 6);
    if (img < p.n) {
      int bo = img, ho = head;
      if (p.swap) {
        const int flat = img * p.heads + head;  // "(b head)" index reinterpreted as "(head b)" (models/iddpm.py:44-46)
        bo = flat % p.n;
        ho = flat / p.n;
      }
      // both threads of a row write 32 of its 64 output channels
      __nv_bfloat16* orow = p.out + (static_cast<long long>(bo) * kL + (row & 63)) * (p.heads * kDh) + ho * kDh + half * 32;
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_off + half * 32, v);
      tmem_ld_wait();
      uint4* dp = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[8 * jj + 0]) * inv, __uint_as_float(v[8 * jj + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[8 * jj + 2]) * inv, __uint_as_float(v[8 * jj + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[8 * jj + 4]) * inv, __uint_as_float(v[8 * jj + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[8 * jj + 6]) * inv, __uint_as_float(v[8 * jj + 7]) * inv);
        dp[jj] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 128);
  }
}

int attn_tc_mh64_forward(const void* qkv, int n, int heads, float scale, int swap, void* out, float* p_out,
                         cudaStream_t stream) {
  AttnTcMh64Params p;
  memset(&p, 0, sizeof(p));
  const int c3 = heads * 3 * 64;
  uint64_t dims[2] = {(uint64_t)c3, (uint64_t)n * 64};
  uint64_t strides[1] = {(uint64_t)c3 * 2};
  uint32_t boxqk[2] = {64u, 128u}, boxv[2] = {64u, 64u};
  int rc;
  if ((rc = encode_map(&p.qk, qkv, 2, dims, strides, boxqk))) return rc;
  if ((rc = encode_map(&p.v, qkv, 2, dims, strides, boxv))) return rc;
  p.n = n; p.heads = heads; p.swap = swap;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.p_out = p_out;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_mh64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMh64Smem);
    if (e != cudaSuccess) { set_error("attn_tc_mh64: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid(1, heads, (n + 1) / 2);
  cudaError_t e = launch_pdl(attn_tc_mh64_kernel, grid, dim3(kAttnThreads), kMh64Smem, stream, p);
  return check_launch_err(e, "attn_tc_mh64_kernel");
}

constexpr int kAttnMhSmem = kPBytes + kSeq * 128 + 1024;  // P (over Q | K) + V + alignment slack: two CTAs per SM

bool attn_tc_mh_supported(int act_dtype, int heads, int L, int dh, int row_stride, int head_stride, long long batch_stride,
                          int v_transposed, const void* q, const void* k, const void* v, const void* out) {
  if (act_dtype != DMME_BF16 || v_transposed || heads < 1) return false;
  if (!((L == kSeq && (dh == 64 || dh == 32)) || (L == 64 && dh == 64))) return false;
  if (head_stride != 3 * dh || row_stride != heads * 3 * dh || batch_stride != static_cast<long long>(L) * row_stride) return false;
  const __nv_bfloat16* qb = static_cast<const __nv_bfloat16*>(q);
  if (static_cast<const __nv_bfloat16*>(k) != qb + dh || static_cast<const __nv_bfloat16*>(v) != qb + 2 * dh) return false;
  return (reinterpret_cast<uintptr_t>(q) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
}

template <int DH>
static int attn_tc_mh_launch(const void* qkv, int n, int heads, float scale, int swap, void* out, float* p_out,
                             cudaStream_t stream) {
  AttnTcMhParams p;
  memset(&p, 0, sizeof(p));
  const int c3 = heads * 3 * DH;
  uint64_t dims[3] = {(uint64_t)c3, (uint64_t)kSeq, (uint64_t)n};
  uint64_t strides[2] = {(uint64_t)c3 * 2, (uint64_t)kSeq * c3 * 2};
  uint32_t boxq[3] = {64u, 128u, 1u}, boxk[3] = {64u, 256u, 1u}, boxv[3] = {64u, 64u, 1u};
  int rc;
  if ((rc = encode_map(&p.q, qkv, 3, dims, strides, boxq))) return rc;
  if ((rc = encode_map(&p.k, qkv, 3, dims, strides, boxk))) return rc;
  if ((rc = encode_map(&p.v, qkv, 3, dims, strides, boxv))) return rc;
  p.n = n; p.heads = heads; p.swap = swap;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.p_out = p_out;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_mh_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnMhSmem);
    if (e != cudaSuccess) { set_error("attn_tc_mh: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid(kSeq / 128, heads, n);
  cudaError_t e = launch_pdl(attn_tc_mh_kernel<DH>, grid, dim3(kAttnThreads), kAttnMhSmem, stream, p);
  return check_launch_err(e, "attn_tc_mh_kernel");
}

int attn_tc_mh_forward(const void* qkv, int n, int heads, int L, int dh, float scale, int swap, void* out, float* p_out,
                       cudaStream_t stream) {
  if (L == 64) return attn_tc_mh64_forward(qkv, n, heads, scale, swap, out, p_out, stream);
  return dh == 64 ? attn_tc_mh_launch<64>(qkv, n, heads, scale, swap, out, p_out, stream)
                  : attn_tc_mh_launch<32>(qkv, n, heads, scale, swap, out, p_out, stream);
}

bool attn_tc_supported(int act_dtype, int heads, int L, int dh, int row_stride, long long batch_stride,
                       int v_transposed, long long v_batch_stride, int swap) {
  return act_dtype == DMME_BF16 && heads == 1 && L == kSeq && dh % 64 == 0 && dh >= 64 && dh <= 256 &&
         row_stride == dh && batch_stride == static_cast<long long>(L) * dh && v_transposed &&
         v_batch_stride == static_cast<long long>(L) * dh && !swap;
}

int attn_tc_forward(const void* q, const void* k, const void* vt, int n, int d, float scale, void* out,
                    cudaStream_t stream) {
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)kSeq, (uint64_t)n};
    uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)kSeq * d * 2};
    uint32_t boxq[3] = {64u, 128u, 1u}, boxk[3] = {64u, 256u, 1u};
    if ((rc = encode_map(&p.q, q, 3, dims, strides, boxq))) return rc;
    if ((rc = encode_map(&p.k, k, 3, dims, strides, boxk))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)kSeq, (uint64_t)d, (uint64_t)n};
    uint64_t strides[2] = {(uint64_t)kSeq * 2, (uint64_t)kSeq * d * 2};
    uint32_t box[3] = {64u, (uint32_t)d, 1u};
    if ((rc = encode_map(&p.vt, vt, 3, dims, strides, box))) return rc;
  }
  p.n = n; p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) { set_error("attn_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid(kSeq / 128, n);
  cudaError_t e = launch_pdl(attn_tc_kernel, grid, dim3(kAttnThreads), kAttnSmem, stream, p);
  return check_launch_err(e, "attn_tc_kernel");
}

}  // namespace dmme
