// Backward of the multi-head attention core (MultiHeadAttention.forward_attention, models/iddpm.py:36-59) on tcgen05,
// fused: one CTA per (image, head) keeps Q, K, V, dO in shared memory, recomputes the softmax matrix from Q and K, and
// produces dQ, dK, dV without a single L x L matrix ever reaching global memory.  (The strided-product path it replaces
// wrote and re-read three fp32 [n heads][L][L] matrices per site -- 134 MB each at 256 tokens and batch 128 -- and needed
// the forward pass to save P.)
//
//   stats   S = Q_qb K^T (128 x 256, TMEM)          -> row max m, row sum l            (per 128-query block qb)
//           delta = rowsum(dO o O)                                                        (from the saved forward output)
//   main    for key block kb, query block qb (128 x 128 tiles):
//             S_t = Q_qb K_kb^T, dP_t = dO_qb V_kb^T                                      (TMEM columns [0,128), [128,256))
//             P_t = exp2(S_t scale log2e - m) / l,  dS_t = P_t o (dP_t - delta) scale     -> bf16 tiles in shared memory
//             dV_kb += P_t^T dO_qb,  dK_kb += dS_t^T Q_qb,  dQ_qb += dS_t K_kb            (TMEM columns [256,512))
//
// Operand layouts: every tile is [rows][64 elements] with SWIZZLE_128B.  Such a tile is K-major when its rows are the M / N
// index (Q, K, V, dO in S_t and dP_t; dS_t in dQ) and MN-major when its rows are the reduction index (P_t^T, dS_t^T, and dO,
// Q, K as the B operands of dV, dK, dQ) -- the same bytes, only the descriptor's major bits differ (as in conv_wgrad_tc.cu).
// q | k | v of head h are channel ranges of the packed [n][L][3C] tensor (models/iddpm.py:38-39); dO and O are read at the
// reference's "(b head) -> (head b)" position (models/iddpm.py:44-46).
//
// L = 64 (the 8x8 sites): one CTA takes TWO images of one head (128 contiguous token rows); the 128 x 128 tile then holds
// two independent 64 x 64 problems on its diagonal and the off-diagonal blocks are masked (P = dS = 0), as in
// attn_tc_mh64_kernel.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..9 = softmax / dS / epilogue (two warps per TMEM lane quarter,
// each half of a tile's columns).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct AttnBwdParams {
  CUtensorMap qkv;   // packed [n * L rows][3C], box [64 ch][RB rows]
  CUtensorMap dout;  // [n * L rows][C], box [64 ch][RB rows]
  int n, heads, L;
  float scale, scale_log2e;
  int swap;
  const __nv_bfloat16* o;     // forward output [n][L][C]
  const __nv_bfloat16* dout_p;
  __nv_bfloat16* dqkv;        // packed like qkv
};

constexpr int kBwdThreads = 320;
constexpr int kBwdSmWarps = 8;
constexpr int kBwdTile = 128 * 128;  // [128 rows][64 bf16]

__host__ __device__ constexpr uint32_t bwd_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}
// MN-major SWIZZLE_128B operand: rows = reduction index (128 B = 64 elements of the M / N index), 8-row groups 1024 B apart,
// 64-element blocks of the M / N index `lbo` bytes apart
__device__ __forceinline__ uint64_t bwd_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// L256: 256 tokens per image, two query blocks and two key blocks of 128 per CTA; otherwise 64 tokens and two images per CTA.
// DH = 32 (the 128-channel sites, models/iddpm.py: four heads of 32 channels): the same 64-channel TMA boxes, each starting at
// its own tensor -- [q | k], [k | v], [v | next head], [dO_h | dO_h+1] -- so every operand's first 32 elements are the head's.
// The products that reduce over the channels (S, dP) take two K = 16 steps instead of four; the ones that produce channels
// (dV, dK, dQ) keep N = 64 and their upper 32 accumulator columns -- products with the neighbouring tensor -- are not stored.
template <bool L256, int DH>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnBwdParams p) {
  constexpr int kDh = DH;
  constexpr int kN = 64;  // channels per operand tile / accumulator block
  constexpr int NB = L256 ? 2 : 1;          // 128-row blocks of queries / keys held by the CTA
  constexpr int kRows = NB * 128;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t in_full, st_full, p_ready, acc_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float row_part[2][128];
  __shared__ float row_m[kRows], row_il[kRows], row_delta[kRows];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* qbuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // [kRows][64]
  uint8_t* kbuf = qbuf + NB * kBwdTile;
  uint8_t* vbuf = kbuf + NB * kBwdTile;
  uint8_t* dobuf = vbuf + NB * kBwdTile;
  uint8_t* pbuf = dobuf + NB * kBwdTile;    // P_t: two 64-key blocks of [128 q][64]
  uint8_t* dsbuf = pbuf + 2 * kBwdTile;     // dS_t, same layout

  const int head = blockIdx.x;
  const int unit = blockIdx.y;              // L256: image; else: pair of images
  const int C = p.heads * kDh;
  const int ch0 = head * 3 * kDh;
  const int row0 = unit * kRows;            // first token row in the [n * L][.] views

  if (threadIdx.x == 0) {
    mbar_init(&in_full, 1);
    mbar_init(&st_full, 1);
    mbar_init(&p_ready, kBwdSmWarps * 32);
    mbar_init(&acc_done, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qkv);
    tma_prefetch_desc(&p.dout);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tm_s = tmem, tm_dp = tmem + 128, tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384;
  pdl_trigger();
  pdl_wait();

  // where this (image, head)'s rows of O / dO live: the forward wrote them at the regrouped position
  auto out_pos = [&](int img, int& bo, int& ho) {
    bo = img; ho = head;
    if (p.swap) {
      const int flat = img * p.heads + head;
      bo = flat % p.n;
      ho = flat / p.n;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&in_full, 4 * NB * kBwdTile);
      for (int b = 0; b < NB; ++b) {
        tma_load_2d(qbuf + b * kBwdTile, &p.qkv, &in_full, ch0, row0 + b * 128);
        tma_load_2d(kbuf + b * kBwdTile, &p.qkv, &in_full, ch0 + kDh, row0 + b * 128);
        tma_load_2d(vbuf + b * kBwdTile, &p.qkv, &in_full, ch0 + 2 * kDh, row0 + b * 128);
      }
      if (L256) {
        int bo, ho;
        out_pos(unit, bo, ho);
        for (int b = 0; b < NB; ++b) tma_load_2d(dobuf + b * kBwdTile, &p.dout, &in_full, ho * kDh, bo * p.L + b * 128);
      } else {
        // two images, each 64 rows at its own regrouped position (an image past the batch: rows out of bounds -> zeros)
        for (int i = 0; i < 2; ++i) {
          int bo, ho;
          out_pos(2 * unit + i, bo, ho);
          const int r = 2 * unit + i < p.n ? bo * p.L : p.n * p.L;
          tma_load_2d(dobuf + i * 64 * 128, &p.dout, &in_full, ho * kDh, r);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int n_st = 0, n_pr = 0;
      mbar_wait(&in_full, 0);
      tc_fence_after();
      const uint32_t qa = smem_u32(qbuf), ka = smem_u32(kbuf), va = smem_u32(vbuf), doa = smem_u32(dobuf);
      const uint32_t pa = smem_u32(pbuf), dsa = smem_u32(dsbuf);
      // ---- row statistics: S = Q_qb K^T over all keys ----
      for (int qb = 0; qb < NB; ++qb) {
        if (qb > 0) { mbar_wait(&p_ready, (n_pr++) & 1); tc_fence_after(); }
        const uint64_t ad = umma_desc_sw128(qa + qb * kBwdTile), bd = umma_desc_sw128(ka);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) umma_bf16(tm_s, ad + 2 * k, bd + 2 * k, bwd_idesc(128, kRows, 0, 0), k != 0 ? 1u : 0u);
        umma_commit(&st_full);
        ++n_st;
      }
      mbar_wait(&p_ready, (n_pr++) & 1);
      tc_fence_after();
      // ---- main loop ----
      for (int kb = 0; kb < NB; ++kb) {
        for (int qb = 0; qb < NB; ++qb) {
          const uint64_t qd = umma_desc_sw128(qa + qb * kBwdTile), kd = umma_desc_sw128(ka + kb * kBwdTile);
          const uint64_t dod = umma_desc_sw128(doa + qb * kBwdTile), vd = umma_desc_sw128(va + kb * kBwdTile);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k) umma_bf16(tm_s, qd + 2 * k, kd + 2 * k, bwd_idesc(128, 128, 0, 0), k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k) umma_bf16(tm_dp, dod + 2 * k, vd + 2 * k, bwd_idesc(128, 128, 0, 0), k != 0 ? 1u : 0u);
          umma_commit(&st_full);
          ++n_st;
          mbar_wait(&p_ready, (n_pr++) & 1);
          tc_fence_after();
          // dV_kb += P_t^T dO_qb ; dK_kb += dS_t^T Q_qb   (M = 128 keys, N = 64, reduction over the tile's 128 query rows)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tm_dv, bwd_desc_mn(pa + k * 2048, kBwdTile), bwd_desc_mn(doa + qb * kBwdTile + k * 2048, kBwdTile),
                      bwd_idesc(128, kN, 1, 1), (qb | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tm_dk, bwd_desc_mn(dsa + k * 2048, kBwdTile), bwd_desc_mn(qa + qb * kBwdTile + k * 2048, kBwdTile),
                      bwd_idesc(128, kN, 1, 1), (qb | k) != 0 ? 1u : 0u);
          // dQ_qb += dS_t K_kb   (M = 128 queries, N = 64, reduction over the tile's 128 keys: two 64-key blocks of dS_t)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tm_dq + qb * kN, umma_desc_sw128(dsa + (k >> 2) * kBwdTile) + 2 * (k & 3),
                      bwd_desc_mn(ka + kb * kBwdTile + k * 2048, kBwdTile), bwd_idesc(128, kN, 0, 1), (kb | k) != 0 ? 1u : 0u);
          umma_commit(&acc_done);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;                // row inside a 128-row block
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int tid = threadIdx.x - 64;             // 0..255
    const float sl = p.scale_log2e;
    int n_st = 0, n_acc = 0;
    // ---- delta = rowsum(dO o O): one row per thread, straight from global memory (128 B of each) ----
    {
      const int r = tid;                          // row of the CTA's kRows (L256) / of the pair's 128 (else)
      float d = 0.f;
      int img, tok;
      if (L256) { img = unit; tok = r; } else { img = 2 * unit + (r >> 6); tok = r & 63; }
      if (r < kRows && img < p.n) {
        int bo, ho;
        out_pos(img, bo, ho);
        const long long off = (static_cast<long long>(bo) * p.L + tok) * C + ho * kDh;
        const uint4* po = reinterpret_cast<const uint4*>(p.o + off);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dout_p + off);
#pragma unroll
        for (int j = 0; j < kDh / 8; ++j) {
          const uint4 a = __ldg(po + j), b = __ldg(pd + j);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a0, a1, b0, b1;
            unpack_bf16x2(aw[e], a0, a1);
            unpack_bf16x2(bw[e], b0, b1);
            d = fmaf(a0, b0, fmaf(a1, b1, d));
          }
        }
      }
      if (r < kRows) row_delta[r] = d;
    }
    // ---- row statistics ----
    for (int qb = 0; qb < NB; ++qb) {
      mbar_wait(&st_full, (n_st++) & 1);
      tc_fence_after();
      // this thread's columns: half of the keys (L256: 128 of 256; else: the 64 keys of image `half`, and only rows of
      // that image have any)
      constexpr int kCols = L256 ? 128 : 64;
      const int c_lo = half * kCols;
      const bool own = L256 || (row >> 6) == half;
      float mx = -INFINITY;
      if (own) {
#pragma unroll 1
        for (int c = c_lo; c < c_lo + kCols; c += 32) {
          uint32_t v[32];
          tmem_ld32(tm_s + lane_off + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      if (L256) {
        row_part[half][row] = mx;
        asm volatile("bar.sync 1, %0;" ::"n"(kBwdSmWarps * 32) : "memory");
        mx = fmaxf(mx, row_part[half ^ 1][row]);
        asm volatile("bar.sync 1, %0;" ::"n"(kBwdSmWarps * 32) : "memory");
      }
      const float mxs = mx * sl;
      float sum = 0.f;
      if (own) {
#pragma unroll 1
        for (int c = c_lo; c < c_lo + kCols; c += 32) {
          uint32_t v[32];
          tmem_ld32(tm_s + lane_off + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum += exp2f(fmaf(__uint_as_float(v[j]), sl, -mxs));
        }
      }
      if (L256) {
        row_part[half][row] = sum;
        asm volatile("bar.sync 1, %0;" ::"n"(kBwdSmWarps * 32) : "memory");
        sum += row_part[half ^ 1][row];
        if (half == 0) { row_m[qb * 128 + row] = mxs; row_il[qb * 128 + row] = 1.0f / sum; }
      } else if (own) {
        row_m[row] = mxs;
        row_il[row] = 1.0f / sum;
      }
      tc_fence_before();
      mbar_arrive(&p_ready);
      // row_part is reused by the next block's exchange; row_m / row_il / row_delta become visible to every thread
      asm volatile("bar.sync 1, %0;" ::"n"(kBwdSmWarps * 32) : "memory");
    }
    // ---- main loop ----
    for (int kb = 0; kb < NB; ++kb) {
      for (int qb = 0; qb < NB; ++qb) {
        mbar_wait(&st_full, (n_st++) & 1);
        tc_fence_after();
        // the previous tile's dV / dK / dQ MMAs still read P_t and dS_t in shared memory (across a key-block boundary the
        // flush below has already waited for them)
        if (qb > 0) mbar_wait(&acc_done, (n_acc++) & 1);
        const int r = qb * 128 + row;
        const float mxs = row_m[r], il = row_il[r], dl = row_delta[r];
        const bool own = L256 || (row >> 6) == half;  // else: keys of the other image: P = dS = 0
        uint8_t* prow = pbuf + half * kBwdTile + row * 128;   // this thread's 64 keys = block `half`, row `row`
        uint8_t* dsrow = dsbuf + half * kBwdTile + row * 128;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          float pv[32], dv[32];
          if (own) {
            uint32_t s[32], d[32];
            tmem_ld32(tm_s + lane_off + half * 64 + c, s);
            tmem_ld32(tm_dp + lane_off + half * 64 + c, d);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pv[j] = exp2f(fmaf(__uint_as_float(s[j]), sl, -mxs)) * il;
              dv[j] = pv[j] * (__uint_as_float(d[j]) - dl) * p.scale;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { pv[j] = 0.f; dv[j] = 0.f; }
          }
          const int u0 = c >> 3;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint4 o, e;
            o.x = pack_bf16x2(pv[8 * jj + 0], pv[8 * jj + 1]); o.y = pack_bf16x2(pv[8 * jj + 2], pv[8 * jj + 3]);
            o.z = pack_bf16x2(pv[8 * jj + 4], pv[8 * jj + 5]); o.w = pack_bf16x2(pv[8 * jj + 6], pv[8 * jj + 7]);
            e.x = pack_bf16x2(dv[8 * jj + 0], dv[8 * jj + 1]); e.y = pack_bf16x2(dv[8 * jj + 2], dv[8 * jj + 3]);
            e.z = pack_bf16x2(dv[8 * jj + 4], dv[8 * jj + 5]); e.w = pack_bf16x2(dv[8 * jj + 6], dv[8 * jj + 7]);
            const uint32_t sw = static_cast<uint32_t>((u0 + jj) ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + sw) = o;
            *reinterpret_cast<uint4*>(dsrow + sw) = e;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&p_ready);
      }
      // ---- dV_kb, dK_kb complete: TMEM lane = key row, 32 of the 64 channels per thread ----
      mbar_wait(&acc_done, (n_acc++) & 1);
      tc_fence_after();
      {
        int img, tok;
        if (L256) { img = unit; tok = kb * 128 + row; } else { img = 2 * unit + (row >> 6); tok = row & 63; }
        uint32_t a[32], b[32];
        tmem_ld32(tm_dv + lane_off + half * 32, a);
        tmem_ld32(tm_dk + lane_off + half * 32, b);
        tmem_ld_wait();
        if (img < p.n && half * 32 < kDh) {  // (DH = 32: columns 32..63 belong to the neighbouring tensor)
          __nv_bfloat16* base = p.dqkv + (static_cast<long long>(img) * p.L + tok) * (3 * C) + ch0 + half * 32;
          uint4* dk = reinterpret_cast<uint4*>(base + kDh);
          uint4* dv = reinterpret_cast<uint4*>(base + 2 * kDh);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(a[8 * jj + 0]), __uint_as_float(a[8 * jj + 1]));
            o.y = pack_bf16x2(__uint_as_float(a[8 * jj + 2]), __uint_as_float(a[8 * jj + 3]));
            o.z = pack_bf16x2(__uint_as_float(a[8 * jj + 4]), __uint_as_float(a[8 * jj + 5]));
            o.w = pack_bf16x2(__uint_as_float(a[8 * jj + 6]), __uint_as_float(a[8 * jj + 7]));
            dv[jj] = o;
            o.x = pack_bf16x2(__uint_as_float(b[8 * jj + 0]), __uint_as_float(b[8 * jj + 1]));
            o.y = pack_bf16x2(__uint_as_float(b[8 * jj + 2]), __uint_as_float(b[8 * jj + 3]));
            o.z = pack_bf16x2(__uint_as_float(b[8 * jj + 4]), __uint_as_float(b[8 * jj + 5]));
            o.w = pack_bf16x2(__uint_as_float(b[8 * jj + 6]), __uint_as_float(b[8 * jj + 7]));
            dk[jj] = o;
          }
        }
      }
      // the next key block's first dV / dK MMA (accumulate = 0) is issued only after every thread's p_ready arrival of that
      // block, i.e. after these reads
    }
    // ---- dQ ----
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) {
      int img, tok;
      if (L256) { img = unit; tok = qb * 128 + row; } else { img = 2 * unit + (row >> 6); tok = row & 63; }
      uint32_t a[32];
      tmem_ld32(tm_dq + qb * kN + lane_off + half * 32, a);
      tmem_ld_wait();
      if (img < p.n && half * 32 < kDh) {
        uint4* dq = reinterpret_cast<uint4*>(p.dqkv + (static_cast<long long>(img) * p.L + tok) * (3 * C) + ch0 + half * 32);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(a[8 * jj + 0]), __uint_as_float(a[8 * jj + 1]));
          o.y = pack_bf16x2(__uint_as_float(a[8 * jj + 2]), __uint_as_float(a[8 * jj + 3]));
          o.z = pack_bf16x2(__uint_as_float(a[8 * jj + 4]), __uint_as_float(a[8 * jj + 5]));
          o.w = pack_bf16x2(__uint_as_float(a[8 * jj + 6]), __uint_as_float(a[8 * jj + 7]));
          dq[jj] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

bool attn_bwd_tc_supported(int act_dtype, int heads, int L, int dh) {
  return act_dtype == DMME_BF16 && heads >= 1 && (dh == 64 || dh == 32) && (L == 256 || L == 64);
}

template <bool L256, int DH>
static int attn_bwd_tc_launch(const AttnBwdParams& p, int units, cudaStream_t stream) {
  constexpr int NB = L256 ? 2 : 1;
  constexpr int smem = (4 * NB + 4) * kBwdTile + 1024;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<L256, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaError_t e = launch_pdl(attn_bwd_tc_kernel<L256, DH>, dim3(p.heads, units), dim3(kBwdThreads), smem, stream, p);
  return check_launch_err(e, "attn_bwd_tc_kernel");
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_attention_bwd_fused_supported(int heads, int L, int dh, int act_dtype) {
  return attn_bwd_tc_supported(act_dtype, heads, L, dh) ? 1 : 0;
}

extern "C" int dmme_attention_bwd_fused(const void* qkv, const void* out, const void* dout, void* dqkv, int n, int heads,
                                        int L, int dh, float scale, int head_batch_swap, int act_dtype, void* stream) {
  DMME_REQUIRE(qkv && out && dout && dqkv, DMME_E_BADARG, "attention_bwd_fused: null pointer");
  DMME_REQUIRE(n > 0 && attn_bwd_tc_supported(act_dtype, heads, L, dh), DMME_E_SHAPE,
               "attention_bwd_fused: bf16, 64- or 32-channel heads, 256 or 64 tokens (got L = %d, dh = %d)", L, dh);
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  const int C = heads * dh;
  const int rb = L == 256 ? 128 : 128;
  {
    uint64_t dims[2] = {(uint64_t)(3 * C), (uint64_t)n * L};
    uint64_t strides[1] = {(uint64_t)(3 * C) * 2};
    uint32_t box[2] = {64u, (uint32_t)rb};
    int rc = encode_map(&p.qkv, qkv, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)n * L};
    uint64_t strides[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64u, L == 256 ? 128u : 64u};
    int rc = encode_map(&p.dout, dout, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.n = n; p.heads = heads; p.L = L;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.swap = head_batch_swap;
  p.o = static_cast<const __nv_bfloat16*>(out);
  p.dout_p = static_cast<const __nv_bfloat16*>(dout);
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dh == 32) return L == 256 ? attn_bwd_tc_launch<true, 32>(p, n, st) : attn_bwd_tc_launch<false, 32>(p, (n + 1) / 2, st);
  return L == 256 ? attn_bwd_tc_launch<true, 64>(p, n, st) : attn_bwd_tc_launch<false, 64>(p, (n + 1) / 2, st);
}
