// The whole single-head attention block of the 16x16 level (Attention.forward, models/ddpm.py:54-75) as ONE kernel:
//
//   h   = GroupNorm(x)                 (a, b) per (image, channel) from dmme_groupnorm_coeff, applied to the tile in shared memory
//   q,k,v = qkv_proj(h)                three 256-deep GEMMs, weights streamed by TMA, results kept in shared memory as bf16
//   S   = q k^T, P = softmax(scale S)  fp32 in TMEM, P rounded to bf16 into shared memory
//   O   = P v                          fp32 in TMEM
//   out = x + proj(O)                  computed transposed (lane = output channel) so the stores are 64 contiguous bytes per
//                                      token and the GroupNorm statistics of `out` accumulate in the thread
//
// Nothing between x and out touches global memory (the four-launch path wrote and re-read the normalised tensor, Q, K,
// V^T and O: 10 x 33.5 MB at batch 256).  L = 256 tokens, C = 256 channels.
//
// One CLUSTER OF TWO CTAs per image, cta_group::2 MMAs (M = 256, N = 256, K = 16): CTA r owns tokens [128 r, 128 r + 128).
// Every product has a natural split in which each CTA stages only what it produced itself:
//   Q, K  [tok][ch]  = H W^T     A = H (own tokens)           B = W rows [128 r, +128)      D lanes = own tokens
//   V^T   [ch][tok]  = Wv H^T    A = Wv rows [128 r, +128)    B = H (own tokens)            D lanes = channels [128 r, +128)
//   S     [q][key]   = Q K^T     A = Q (own queries)          B = K (own keys)              D lanes = own queries
//   O     [q][ch]    = P V       A = P (own queries)          B = V^T (own channel rows)    D lanes = own queries
//   out^T [ch][tok]  = Wp O^T    A = Wp rows [128 r, +128)    B = O (own tokens)            D lanes = channels [128 r, +128)
// Shared memory per CTA: three 64 KB regions R1..R3, each four [128 rows][64] bf16 SWIZZLE_128B blocks, + a 2 x 16 KB ring:
//   R1: x -> H (in place)            -> V^T                      -> next image's x (from "O complete")
//   R2: Wq slab -> Q -> P            -> Wproj slab (from "O complete")   -> next image's Wq (from "out^T complete")
//   R3: Wk slab -> K -> O (bf16)     -> next image's Wk (from "out^T complete")
//   ring: the four k-blocks of the Wv slab
// TMEM per CTA: T0 = columns [0, 256): Q, V^T, O;  T1 = [256, 512): K, S, out^T.
// Warps: 0 = TMA producer, 1 = MMA issuer (leader CTA only), 2..9 = 256 workers (the four accumulator drains, softmax,
// output epilogue), 10..13 = GroupNorm transform of the NEXT image's tile.  Worker phases of the two CTAs are joined on the
// leader's mbarriers (remote arrives); P V and the projection start on the 64-wide k-blocks the workers have finished.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct AttnBlockParams {
  CUtensorMap x;      // [n][256][256] bf16 as (c, token, n), box [64][128][1]
  CUtensorMap wqkv;   // [3C][C] bf16 as (k, row), box [64][C/2]: B-operand halves of Wq / Wk
  CUtensorMap wqkv_a; // the same tensor, box [64][128]: A-operand rows of Wv
  CUtensorMap wproj;  // [C][C] bf16, box [64][128]
  int n;
  float scale_log2e;
  const float* gn_ab;         // [n][C][2], or null: coefficients from stats_in / gamma / beta in the kernel
  const long long* stats_in;  // micro-group sums of x written by its producer ([n][C/4][2])
  const float* gamma; const float* beta;
  int cpg; float eps;
  const float* bias_qkv;      // [768]
  const float* bias_proj;     // [256]
  const __nv_bfloat16* xres;  // the raw x again (residual), [n][256][256]
  __nv_bfloat16* out;         // [n][256][256]
  long long* stats;           // optional [n][64][2]
  long long* trace;           // debugging: clock64 timeline of cluster 0's leader (dmme_debug_set_attn_block_trace), or null
};

static long long* g_ab_trace = nullptr;
// role 0: worker warp 2 lane 0 (14 events per image), role 1: MMA thread (10 events per image), role 2: inside the Q drain
// and the output epilogue (16 per image)
__device__ __forceinline__ void ab_trace(long long* trace, int role, int idx) {
  if (trace != nullptr && blockIdx.x == 0 && idx < 256) trace[role * 256 + idx] = clock64();
}

constexpr int kAbC = 256, kAbL = 256;
constexpr int kAbBlk = 128 * 128;          // one [128 rows][64 bf16] block
constexpr int kAbRegion = 4 * kAbBlk;      // 64 KB
constexpr int kAbWorkers = 256;
constexpr int kAbXform = 64;               // GroupNorm transform warps (run one image ahead of the workers)
constexpr int kAbThreads = 64 + kAbWorkers + kAbXform;
constexpr int kAbSmem = 3 * kAbRegion + 2 * kAbBlk + 1024;  // the 256-channel kernels

enum {
  AB_X_FULL = 0, AB_H_READY, AB_WQ_FULL, AB_WK_FULL, AB_WP_FULL, AB_WV_FULL0, AB_WV_FULL1, AB_WV_EMPTY0, AB_WV_EMPTY1,
  AB_Q_DONE, AB_K_DONE, AB_V_DONE, AB_S_DONE, AB_O_DONE, AB_D_DONE,
  AB_QS_READY, AB_KS_READY, AB_EPI_DONE,  // joined by the 2 x 256 workers
  AB_PV_READY0,                           // + key block: P block and V^T block written (2 x 128 workers of that half each)
  AB_OS_READY0 = AB_PV_READY0 + 4,        // + channel block of O
  AB_NBARS = AB_OS_READY0 + 4
};

// L2 prefetch of one tensor-map box (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1)
               : "memory");
}

// wait on a barrier that threads of the peer CTA arrive on (release.cluster): acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0x3ff) == 0 && (clock64() - t0) > DMME_MBAR_TIMEOUT_CYCLES) {
      printf("dmme: cluster mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// (a, b) of y = a x + b for channels c0 .. c0 + 7 of image img: from the coefficient tensor of dmme_groupnorm_coeff, or --
// saving that launch -- from the producer's fixed-point micro-group sums with gn_coefficients' arithmetic (groupnorm.cu: one
// source, no scale / shift; same bits).  cpg is a multiple of 4.
__device__ __forceinline__ void ab_coeff8(const float* gn_ab, const long long* stats_in, const float* gamma, const float* beta,
                                          int cpg, float eps, int hw, int C, int img, int c0, float (&a)[8], float (&b)[8]) {
  if (gn_ab != nullptr) {
    const float4* abp = reinterpret_cast<const float4*>(gn_ab + (static_cast<long long>(img) * C + c0) * 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t4 = __ldg(abp + j);
      a[2 * j] = t4.x; b[2 * j] = t4.y; a[2 * j + 1] = t4.z; b[2 * j + 1] = t4.w;
    }
    return;
  }
  const float inv_cnt = 1.0f / (static_cast<float>(hw) * cpg);
  const double unfix = 1.0 / static_cast<double>(1 << DMME_STATS_FRAC_BITS);
  float mean = 0.f, rs = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    if (e == 0 || c % cpg == 0) {
      const int g0 = (c / cpg) * cpg;
      long long s1 = 0, s2 = 0;
      for (int cc = g0; cc < g0 + cpg; cc += 4) {
        const long long* st = stats_in + (static_cast<long long>(img) * (C >> 2) + (cc >> 2)) * 2;
        s1 += st[0];
        s2 += st[1];
      }
      mean = static_cast<float>(static_cast<double>(s1) * unfix) * inv_cnt;
      const float ex2 = static_cast<float>(static_cast<double>(s2) * unfix) * inv_cnt;
      const float var = fmaxf(ex2 - mean * mean, 0.f);
      rs = rsqrtf(var + eps);
    }
    const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
    a[e] = rs * ga;
    b[e] = be - mean * rs * ga;
  }
}

// 32 fp32 values -> 64 bytes of a [rows][64 bf16] SWIZZLE_128B block row: 16-byte units u0 .. u0 + 3, XOR-ed with (row & 7)
__device__ __forceinline__ void ab_store_units(uint8_t* block_row, int u0, int swz, const float (&f)[32]) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    uint4 o;
    o.x = pack_bf16x2(f[8 * jj + 0], f[8 * jj + 1]);
    o.y = pack_bf16x2(f[8 * jj + 2], f[8 * jj + 3]);
    o.z = pack_bf16x2(f[8 * jj + 4], f[8 * jj + 5]);
    o.w = pack_bf16x2(f[8 * jj + 6], f[8 * jj + 7]);
    *reinterpret_cast<uint4*>(block_row + (((u0 + jj) ^ swz) << 4)) = o;
  }
}
// columns [c, c + 32) of row `row` of a region of [128 rows][64] blocks
__device__ __forceinline__ void ab_store_chunk(uint8_t* prow, int c, int row, const float (&f)[32]) {
  ab_store_units(prow + (c >> 6) * kAbBlk, (c & 63) >> 3, row & 7, f);
}

// TMEM [this warp's 32 lanes][32 NCH columns from c_lo] fp32 -> bf16 rows of a region of [rows][64] SWIZZLE_128B blocks
// `blk_stride` bytes apart.  MODE 1: + lv; 2: * lv; 3: as is.  bar_a / bar_b (shared::cluster addresses, 0 = none): arrived
// on after the first / second 64-column block has been written (a consumer MMA may start on that k-block while the other is
// still being drained).
template <int MODE, int NCH>
__device__ __forceinline__ void ab_drain(uint32_t tmem, int c_lo, int row, uint8_t* region, int blk_stride, float lv,
                                         uint32_t bar_a, uint32_t bar_b) {
  uint8_t* prow = region + row * 128;
#pragma unroll 1
  for (int i = 0; i < NCH; ++i) {
    const int c = c_lo + 32 * i;
    uint32_t v[32];
    tmem_ld32(tmem + c, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      f[j] = MODE == 1 ? __uint_as_float(v[j]) + lv : (MODE == 2 ? __uint_as_float(v[j]) * lv : __uint_as_float(v[j]));
    ab_store_units(prow + (c >> 6) * blk_stride, (c & 63) >> 3, row & 7, f);
    if ((i & 1) && bar_a != 0u) {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_cluster(i == 1 ? bar_a : bar_b);
    }
  }
}

// the same with a per-COLUMN bias (Q) read from shared memory (broadcast 16-byte loads).  The first version fetched the bias
// with global loads inside the loop: the 228 KB shared-memory carve-out leaves next to no L1, every chunk paid an L2 round
// trip and the drain took 3 k clocks instead of 1 k.
template <int NCH>
__device__ __forceinline__ void ab_drain_colbias(uint32_t tmem, int c_lo, int row, uint8_t* region, const float* sbias,
                                                 long long* trace, int tbase) {
  uint8_t* prow = region + row * 128;
#pragma unroll 1
  for (int i = 0; i < NCH; ++i) {
    const int c = c_lo + 32 * i;
    uint32_t v[32];
    tmem_ld32(tmem + c, v);
    float f[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + 4 * j);
      f[4 * j + 0] = b4.x; f[4 * j + 1] = b4.y; f[4 * j + 2] = b4.z; f[4 * j + 3] = b4.w;
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
    ab_store_chunk(prow, c, row, f);
    if (tbase >= 0) ab_trace(trace, 2, tbase + i);
  }
}

// C = 256 or 128 channels (both: 256 tokens).  With C = 128 the two products whose M side is the CHANNEL axis (V^T and out^T)
// would have M = 128, i.e. 64 rows per CTA of a cta_group::2 pair; they are issued with M = 256 instead, both CTAs staging
// ALL 128 weight rows, so that each CTA ends up with the complete [128 channels][256 tokens] result in its own TMEM (the
// products are small: 2 x the MMA work of a K = 128 product) and keeps the part it needs: channel rows [64 r, +64) of V^T (its
// N-half of P V), tokens [128 r, +128) of out^T.
template <int C>
__global__ void __launch_bounds__(kAbThreads, 1) attn_block_kernel(const __grid_constant__ AttnBlockParams p) {
  constexpr int KB = C / 64;               // 64-wide k-blocks of a C-deep product = blocks of an [128][C] region
  constexpr int kRegA = 256 * C;           // bytes of an [128 rows][C] bf16 region (= of V^T: [C/2 rows][256 keys])
  constexpr int kHalfBlk = (C / 2) * 128;  // bytes of one k-block of a B-operand half: [C/2 rows][64]
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[AB_NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float row_part[2][128];  // softmax row exchange; between images: the C Q biases

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* R1 = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* R2 = R1 + kRegA;        // 64 KB: P is [128][256 keys] whatever C is
  uint8_t* R3 = R2 + kAbRegion;
  uint8_t* ring = R3 + kRegA;

  const uint32_t rank = cluster_ctarank();
  const int cid = static_cast<int>(blockIdx.x >> 1), ncl = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < AB_NBARS; ++i) {
      const int count = i == AB_H_READY ? 2 * kAbXform : (i >= AB_PV_READY0 ? kAbWorkers : (i >= AB_QS_READY ? 2 * kAbWorkers : 1));
      mbar_init(&bars[i], count);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.x);
    tma_prefetch_desc(&p.wqkv);
    tma_prefetch_desc(&p.wqkv_a);
    tma_prefetch_desc(&p.wproj);
  }
  if (warp == 1) tmem_alloc_2sm(&tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers exist before either signals the other's
  tc_fence_after();
  const uint32_t T0 = tmem_slot, T1 = tmem_slot + 256;
  pdl_trigger();  // after the TMEM allocation (see common.cuh)

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      pdl_wait();
      // B-operand half of Wq / Wk: rows [row0, row0 + C/2) x C k; both CTAs' halves complete on the leader's barrier
      auto load_half_slab = [&](uint8_t* dst, int bar, int row0) {
        if (rank == 0) mbar_expect_tx(&bars[bar], 2 * KB * kHalfBlk);
        const uint32_t lead = mapa_u32(&bars[bar], 0);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d_2sm(dst + kb * kHalfBlk, &p.wqkv, lead, kb * 64, row0);
      };
      auto load_x = [&](int img) {
        mbar_expect_tx(&bars[AB_X_FULL], kRegA);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(R1 + kb * kAbBlk, &p.x, &bars[AB_X_FULL], kb * 64, static_cast<int>(rank) * 128, img);
      };
      const int rh = static_cast<int>(rank) * (C / 2);        // this CTA's N-half of Wq / Wk
      const int ra = C == 256 ? static_cast<int>(rank) * 128 : 0;  // its 128 A-operand rows of Wv / Wp (C = 128: all of them)
      int it = 0, wv_it = 0;
      if (cid < p.n) {
        load_x(cid);
        load_half_slab(R2, AB_WQ_FULL, rh);
        load_half_slab(R3, AB_WK_FULL, C + rh);
      }
      for (int img = cid; img < p.n; img += ncl, ++it) {
        const uint32_t ph = it & 1;
        for (int kb = 0; kb < KB; ++kb, ++wv_it) {
          const int s = wv_it & 1;
          mbar_wait(&bars[AB_WV_EMPTY0 + s], ((wv_it >> 1) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&bars[AB_WV_FULL0 + s], 2 * kAbBlk);
          tma_load_2d_2sm(ring + s * kAbBlk, &p.wqkv_a, mapa_u32(&bars[AB_WV_FULL0 + s], 0), kb * 64, 2 * C + ra);
        }
        const int nxt = img + ncl;
        mbar_wait(&bars[AB_O_DONE], ph);  // P (R2) and V^T (R1) have been read
        {
          if (rank == 0) mbar_expect_tx(&bars[AB_WP_FULL], 2 * KB * kAbBlk);
          const uint32_t lead = mapa_u32(&bars[AB_WP_FULL], 0);
          for (int kb = 0; kb < KB; ++kb) tma_load_2d_2sm(R2 + kb * kAbBlk, &p.wproj, lead, kb * 64, ra);
        }
        if (nxt < p.n) load_x(nxt);
        if (nxt < p.n) {
          mbar_wait(&bars[AB_D_DONE], ph);  // Wproj (R2) and O (R3) have been read
          load_half_slab(R2, AB_WQ_FULL, rh);
          load_half_slab(R3, AB_WK_FULL, C + rh);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc_c = umma_idesc_bf16(256, C);    // N = channels: Q, K, O
      constexpr uint32_t idesc_l = umma_idesc_bf16(256, 256);  // N = tokens / keys: V^T, S, out^T
      // one 64-wide k-block: four K = 16 steps
      auto mma_kb = [&](uint32_t tm, const uint8_t* a_blk, const uint8_t* b_blk, uint32_t idesc, bool first) {
        const uint64_t ad = umma_desc_sw128(smem_u32(a_blk)), bd = umma_desc_sw128(smem_u32(b_blk));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_2sm(tm, ad + 2 * k, bd + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
      };
      int it = 0, wv_it = 0;
      for (int img = cid; img < p.n; img += ncl, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait_cluster(&bars[AB_H_READY], ph);
        mbar_wait(&bars[AB_WQ_FULL], ph);
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 0);
        // Q = H Wq^T  (T0: the previous image's O was drained before its projection was issued)
        for (int kb = 0; kb < KB; ++kb) mma_kb(T0, R1 + kb * kAbBlk, R2 + kb * kHalfBlk, idesc_c, kb == 0);
        umma_commit_2sm(&bars[AB_Q_DONE], 3);
        mbar_wait(&bars[AB_WK_FULL], ph);
        if (it > 0) mbar_wait_cluster(&bars[AB_EPI_DONE], ph ^ 1);  // T1: the previous image's output has been read
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 1);
        for (int kb = 0; kb < KB; ++kb) mma_kb(T1, R1 + kb * kAbBlk, R3 + kb * kHalfBlk, idesc_c, kb == 0);  // K = H Wk^T
        umma_commit_2sm(&bars[AB_K_DONE], 3);
        mbar_wait_cluster(&bars[AB_QS_READY], ph);  // T0 drained, Q in R2
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 2);
        for (int kb = 0; kb < KB; ++kb, ++wv_it) {  // V^T = Wv H^T
          const int s = wv_it & 1;
          mbar_wait(&bars[AB_WV_FULL0 + s], (wv_it >> 1) & 1);
          tc_fence_after();
          ab_trace(p.trace, 1, it * 10 + 3 + kb);
          mma_kb(T0, ring + s * kAbBlk, R1 + kb * kAbBlk, idesc_l, kb == 0);
          umma_commit_2sm(&bars[AB_WV_EMPTY0 + s], 3);
        }
        umma_commit_2sm(&bars[AB_V_DONE], 3);
        mbar_wait_cluster(&bars[AB_KS_READY], ph);  // T1 drained, K in R3
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 7);
        for (int kb = 0; kb < KB; ++kb) mma_kb(T1, R2 + kb * kAbBlk, R3 + kb * kAbBlk, idesc_l, kb == 0);  // S = Q K^T
        umma_commit_2sm(&bars[AB_S_DONE], 3);
        // O = P V, key blocks in the order the two worker halves deliver them (0, 2, then 1, 3).  Blocks 0 and 2 together
        // also say that every worker has drained V^T out of T0.
        mbar_wait_cluster(&bars[AB_PV_READY0 + 0], ph);
        mbar_wait_cluster(&bars[AB_PV_READY0 + 2], ph);
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 8);
        mma_kb(T0, R2 + 0 * kAbBlk, R1 + 0 * kHalfBlk, idesc_c, true);
        mma_kb(T0, R2 + 2 * kAbBlk, R1 + 2 * kHalfBlk, idesc_c, false);
        mbar_wait_cluster(&bars[AB_PV_READY0 + 1], ph);
        mbar_wait_cluster(&bars[AB_PV_READY0 + 3], ph);
        tc_fence_after();
        mma_kb(T0, R2 + 1 * kAbBlk, R1 + 1 * kHalfBlk, idesc_c, false);
        mma_kb(T0, R2 + 3 * kAbBlk, R1 + 3 * kHalfBlk, idesc_c, false);
        umma_commit_2sm(&bars[AB_O_DONE], 3);
        // out^T = Wp O^T, channel blocks of O as they are drained (T1: S was read before the last P blocks were signalled).
        // C = 256: half h of the workers delivers blocks 2 h, 2 h + 1; C = 128: block h
        mbar_wait(&bars[AB_WP_FULL], ph);
        if constexpr (C == 256) {
          mbar_wait_cluster(&bars[AB_OS_READY0 + 0], ph);
          mbar_wait_cluster(&bars[AB_OS_READY0 + 2], ph);
          tc_fence_after();
          ab_trace(p.trace, 1, it * 10 + 9);
          mma_kb(T1, R2 + 0 * kAbBlk, R3 + 0 * kAbBlk, idesc_l, true);
          mma_kb(T1, R2 + 2 * kAbBlk, R3 + 2 * kAbBlk, idesc_l, false);
          mbar_wait_cluster(&bars[AB_OS_READY0 + 1], ph);
          mbar_wait_cluster(&bars[AB_OS_READY0 + 3], ph);
          tc_fence_after();
          mma_kb(T1, R2 + 1 * kAbBlk, R3 + 1 * kAbBlk, idesc_l, false);
          mma_kb(T1, R2 + 3 * kAbBlk, R3 + 3 * kAbBlk, idesc_l, false);
        } else {
          mbar_wait_cluster(&bars[AB_OS_READY0 + 0], ph);
          mbar_wait_cluster(&bars[AB_OS_READY0 + 2], ph);
          tc_fence_after();
          ab_trace(p.trace, 1, it * 10 + 9);
          mma_kb(T1, R2 + 0 * kAbBlk, R3 + 0 * kAbBlk, idesc_l, true);
          mma_kb(T1, R2 + 1 * kAbBlk, R3 + 1 * kAbBlk, idesc_l, false);
        }
        umma_commit_2sm(&bars[AB_D_DONE], 3);
      }
    }
  } else if (warp >= 2 + kAbWorkers / 32) {
    // =========================== GroupNorm transform warps ===========================
    // x tile -> H in place, one image ahead of the workers (the tile of image i + 1 lands once P V of image i is complete and
    // H is not needed before the workers have stored image i: two warps are enough and leave the shared-memory port to the
    // drains): thread = one 16-byte unit column (8 channels, their (a, b) in registers) x every (64 / units)-th token row
    constexpr int kUnits = C / 8, kStep = kAbXform / kUnits;
    const int t2 = static_cast<int>(threadIdx.x) - (64 + kAbWorkers);
    const int cu = t2 % kUnits, kb = cu >> 3, u = cu & 7, r0 = t2 / kUnits;
    const uint32_t bar_h = mapa_u32(&bars[AB_H_READY], 0);
    pdl_wait();
    int it = 0;
    for (int img = cid; img < p.n; img += ncl, ++it) {
      float a[8], b[8];
      ab_coeff8(p.gn_ab, p.stats_in, p.gamma, p.beta, p.cpg, p.eps, kAbL, C, img, kb * 64 + u * 8, a, b);
      mbar_wait(&bars[AB_X_FULL], it & 1);
      if (t2 == 0) ab_trace(p.trace, 0, it * 14 + 0);
      uint8_t* blk = R1 + kb * kAbBlk;
#pragma unroll 2
      for (int jb = 0; jb < 128 / kStep; jb += 4) {
        uint4 v[4];
        uint8_t* addr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = r0 + kStep * (jb + j);
          addr[j] = blk + r * 128 + ((u ^ (r & 7)) << 4);
          v[j] = *reinterpret_cast<const uint4*>(addr[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float lo, hi;
            unpack_bf16x2(w[e], lo, hi);
            w[e] = pack_bf16x2(fmaf(a[2 * e], lo, b[2 * e]), fmaf(a[2 * e + 1], hi, b[2 * e + 1]));
          }
          *reinterpret_cast<uint4*>(addr[j]) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async();
      mbar_arrive_cluster(bar_h);
      if (t2 == 0) ab_trace(p.trace, 0, it * 14 + 1);
    }
  } else {
    // =========================== workers ===========================
    const int wq = warp & 3;             // TMEM lane quarter
    const int half = (warp - 2) >> 2;    // which half of the accumulator columns
    const int row = wq * 32 + lane;      // TMEM lane = tile row
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t bar_qs = mapa_u32(&bars[AB_QS_READY], 0), bar_ks = mapa_u32(&bars[AB_KS_READY], 0),
                   bar_epi = mapa_u32(&bars[AB_EPI_DONE], 0),
                   bar_pv = mapa_u32(&bars[AB_PV_READY0 + 2 * half], 0),  // this half's two key blocks: + 0, + 8 bytes
                   bar_os = mapa_u32(&bars[AB_OS_READY0 + 2 * half], 0);
    // the channel of this thread where lanes are channels (V^T, out^T): C = 256: this CTA's 128; C = 128: all of them
    const int och = C == 256 ? static_cast<int>(rank) * 128 + row : row;
    const int chp = och & ~1, odd = lane & 1;  // the channel pair it stores after the epilogue's lane exchange
    // V^T rows this CTA keeps: C = 256: all of its lanes; C = 128: channels [64 r, +64), i.e. lane quarters 2 r, 2 r + 1
    const bool v_active = C == 256 || (wq >> 1) == static_cast<int>(rank);
    const int v_row = C == 256 ? row : row - 64 * static_cast<int>(rank);
    // out^T columns (tokens) this thread stores: C = 256: [128 half, +128) of all 256; C = 128: [128 r + 64 half, +64)
    constexpr int kOutChunks = C == 256 ? 4 : 2;
    const int t_lo = C == 256 ? half * 128 : static_cast<int>(rank) * 128 + half * 64;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    pdl_wait();
    const float bias_v = __ldg(p.bias_qkv + 2 * C + och);
    const float bias_lo = __ldg(p.bias_proj + chp), bias_hi = __ldg(p.bias_proj + chp + 1);
    float* sbias = &row_part[0][0];
    const int wt = static_cast<int>(threadIdx.x) - 64;
    const float bias_q = wt < C ? __ldg(p.bias_qkv + wt) : 0.f;  // staged in row_part for every image's Q drain
    int it = 0;
    for (int img = cid; img < p.n; img += ncl, ++it) {
      const uint32_t ph = it & 1;
      if (wt < C) sbias[wt] = bias_q;
      asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");
      // ---- Q: T0 -> R2 ----
      mbar_wait(&bars[AB_Q_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 0);
      ab_drain_colbias<C / 64>(T0 + lane_off, half * (C / 2), row, R2, sbias, p.trace, (warp == 2 && lane == 0) ? it * 16 + 1 : -1);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_cluster(bar_qs);
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 3);
      // ---- K: T1 -> R3.  The K bias is left out: it adds (q_i + b_q) . b_k to every score of row i, which the softmax over
      // the keys cancels exactly (the reference adds it and rounds K + b_k; this is the same function of the inputs) ----
      mbar_wait(&bars[AB_K_DONE], ph);
      tc_fence_after();
      ab_drain<3, C / 64>(T1 + lane_off, half * (C / 2), row, R3, kAbBlk, 0.f, 0u, 0u);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_cluster(bar_ks);
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 5);
      // ---- V^T: T0 -> R1 (lanes = channels, columns = tokens = keys; blocks of [C/2 channel rows][64 keys]) ----
      mbar_wait(&bars[AB_V_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 6);
      if (v_active) ab_drain<1, 4>(T0 + lane_off, half * 128, v_row, R1, kHalfBlk, bias_v, 0u, 0u);
      // ---- softmax of this query row (two threads per row, half of the keys each): T1 -> P in R2 ----
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 7);
      mbar_wait(&bars[AB_S_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 8);
      const int c_lo = half * 128;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + 128; c += 32) {
        uint32_t v[32];
        tmem_ld32(T1 + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      row_part[half][row] = mx;
      asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");
      mx = fmaxf(mx, row_part[half ^ 1][row]);
      asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");
      float sum = 0.f;
      {
        const float sl = p.scale_log2e;
        const float mxs = mx * sl;
        uint8_t* prow = R2 + row * 128;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int c = c_lo + 32 * i;
          uint32_t v[32];
          tmem_ld32(T1 + lane_off + c, v);
          tmem_ld_wait();
          float e[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = exp2f(fmaf(__uint_as_float(v[j]), sl, -mxs));
            sum += e[j];
          }
          ab_store_chunk(prow, c, row, e);
          if (i & 1) {
            // a 64-key block of P (and, since the V^T drain above, of V^T) is complete: its P V k-block may be issued
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive_cluster(bar_pv + (i == 1 ? 0u : 8u));
          }
        }
      }
      row_part[half][row] = sum;
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 9);
      asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");
      sum = row_part[0][row] + row_part[1][row];
      asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");  // both halves have read before the next image's maxima land
      // ---- O / rowsum: T0 -> R3, signalled per 64-channel block ----
      mbar_wait(&bars[AB_O_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 10);
      ab_drain<2, C / 64>(T0 + lane_off, half * (C / 2), row, R3, kAbBlk, 1.0f / sum, bar_os, bar_os + 8u);
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 11);
      // ---- out^T: T1 (lane = channel, column = token) + bias + x -> global, statistics of the stored values.  Neighbouring
      // lanes exchange every second value so that a thread stores a channel PAIR of one token (4-byte accesses, half as many
      // of them) ----
      {
        const long long ibase = static_cast<long long>(img) * (kAbL * C) + chp;
        const uint32_t* __restrict__ xr = reinterpret_cast<const uint32_t*>(p.xres + ibase) + odd * (C / 2);
        uint32_t* __restrict__ op = reinterpret_cast<uint32_t*>(p.out + ibase) + odd * (C / 2);
        // all of the residual is requested before the projection is awaited: one chunk of look-ahead left an L2 round trip
        // per chunk exposed
        uint32_t av[16 * kOutChunks];
        if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 5);
#pragma unroll
        for (int i = 0; i < 16 * kOutChunks; ++i) av[i] = __ldg(xr + (t_lo + 2 * i) * (C / 2));
        if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 6);
        mbar_wait(&bars[AB_D_DONE], ph);
        tc_fence_after();
        if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 12);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int ci = 0; ci < kOutChunks; ++ci) {
          const int tok0 = t_lo + ci * 32;
          uint32_t v[32];
          tmem_ld32(T1 + lane_off + tok0, v);
          tmem_ld_wait();
          float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // even lanes take token tok0 + 2 i, odd lanes token tok0 + 2 i + 1, both channels of the pair
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? v[2 * i] : v[2 * i + 1], 1);
            const float dlo = __uint_as_float(odd ? recv : v[2 * i]), dhi = __uint_as_float(odd ? v[2 * i + 1] : recv);
            float xl, xh;
            unpack_bf16x2(av[ci * 16 + i], xl, xh);
            const uint32_t uo = pack_bf16x2(xl + (dlo + bias_lo), xh + (dhi + bias_hi));
            op[(tok0 + 2 * i) * (C / 2)] = uo;
            float lo, hi;
            unpack_bf16x2(uo, lo, hi);
            p1[i & 3] += lo + hi;
            p2[i & 3] = fmaf(lo, lo, p2[i & 3]);
            p2[(i + 2) & 3] = fmaf(hi, hi, p2[(i + 2) & 3]);
          }
          s1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
          s2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
          if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 8 + ci);
        }
        tc_fence_before();
        mbar_arrive_cluster(bar_epi);  // T1 may be overwritten by the next image's K
        if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 13);
        if (p.stats) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
          if ((lane & 3) == 0) {
            unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                     (static_cast<long long>(img) * (C >> 2) + (och >> 2)) * 2;
            atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(s1 * kFix)));
            atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(s2 * kFix)));
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may leave while the other can still signal its barriers / read its tiles
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_slot, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same block at L = 16 tokens (the 4x4 middle ResBlock of the default DDPM UNet: 256 channels): EIGHT images per CTA as
// one 128-row tile (row = image * 16 + token), plain cta_group::1 MMAs with M = 128:
//   Q, K [row][ch]     = H W^T          two N = 128 halves each, weights (B operand) streamed through the ring
//   V^T  [ch][row]     = Wv H^T         two M = 128 halves (A operand = weights)
//   S    [row][row']   = Q K^T          128 x 128; only the eight 16 x 16 diagonal blocks are scores of one image
//   P                  = softmax of the diagonal block of each row, zeros elsewhere (bf16 [128][128])
//   O    [row][ch]     = P V            N = 256
//   out^T[ch][row]     = Wp O^T         two M = 128 halves, epilogue as above (lane = channel)
// The four-launch path took 123 us at batch 256 for this site (GroupNorm, 1x1 qkv, CUDA-core attention, 1x1 projection:
// each launch latency-bound on 4096 positions).  Shared memory: R1 = x -> H -> V^T, R2 = Q -> P, R3 = K -> O, 2 x 16 KB ring
// for the 32 [128 rows][64] weight blocks.  TMEM: Q | K -> V^T (2 x 128 columns) | S (128) -> O (256) | out^T (2 x 128).
// ---------------------------------------------------------------------------------------------------------------------
struct AttnBlock16Params {
  CUtensorMap x;      // [n * 16][256] bf16 as (c, row), box [64][128]; rows past the batch are zero-filled
  CUtensorMap wqkv;   // [768][256] bf16 as (k, row), box [64][128]
  CUtensorMap wproj;  // [256][256]
  int n;
  float scale_log2e;
  const float* gn_ab;
  const long long* stats_in;
  const float* gamma; const float* beta;
  int cpg; float eps;
  const float* bias_qkv;
  const float* bias_proj;
  const __nv_bfloat16* xres;
  __nv_bfloat16* out;
  long long* stats;
  long long* trace;
};

constexpr int kA16Threads = 64 + kAbWorkers;
// Weight staging: 32 blocks of [128 rows][64] (16 KB) stream through up to 14 SLOTS: the two ring stages, and the blocks of
// the data regions while those are not yet (R2, R3: until Q and K are drained) or no longer (R1, R2: once P V is complete)
// in use.  Slot of block j (kernel-wide schedule, the same in the producer and the MMA thread):
//   Wq  j =  0.. 7 -> R2.0-3, R3.0-3        all requested at once, with the x tile
//   Wk  j =  8.. 9 -> ring 0, 1             j = 10..15 -> the slot of block j - 10 once that block has been consumed
//   Wv  j = 16..23 -> ring 0 / 1 alternating (R2 / R3 hold Q / K by then): the one weight matrix that pays the ring's depth
//   Wp  j = 24..31 -> R1.0-3, R2.0-3        requested when P V is complete, lands while O is drained
// With two ring stages only, the 32 blocks were 16 exposed round trips to HBM (the step's activations evict the weights from
// L2 between steps): 44 us per launch, of which ~25 us waiting.
constexpr int kA16Slots = 14;  // 0, 1: ring; 2..5: R2 blocks; 6..9: R3 blocks; 10..13: R1 blocks
__host__ __device__ constexpr int a16_slot(int j) {
  return j < 8 ? 2 + j : (j < 10 ? j - 8 : (j < 16 ? 2 + (j - 10) : (j < 24 ? (j & 1) : (j < 28 ? 10 + (j - 24) : 2 + (j - 28)))));
}
enum {
  B16_X_FULL = 0, B16_SLOT_FULL0, B16_SLOT_EMPTY0 = B16_SLOT_FULL0 + kA16Slots, B16_Q_DONE = B16_SLOT_EMPTY0 + kA16Slots,
  B16_K_DONE, B16_V_DONE, B16_S_DONE, B16_O_DONE, B16_D_DONE,
  B16_H_READY, B16_QK_DRAINED, B16_PV_READY, B16_OS_READY,  // joined by the 256 workers
  B16_NBARS
};

__global__ void __launch_bounds__(kA16Threads, 1) attn_block16_kernel(const __grid_constant__ AttnBlock16Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[B16_NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float row_part[256];  // the Q biases, then the softmax row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* R1 = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* R2 = R1 + kAbRegion;
  uint8_t* R3 = R2 + kAbRegion;
  uint8_t* ring = R3 + kAbRegion;
  const int img0 = static_cast<int>(blockIdx.x) * 8;
  const int row0 = img0 * 16;
  const int nrows = p.n * 16;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B16_NBARS; ++i) mbar_init(&bars[i], i >= B16_H_READY ? kAbWorkers : 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.x);
    tma_prefetch_desc(&p.wqkv);
    tma_prefetch_desc(&p.wproj);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t T = tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      mbar_expect_tx(&bars[B16_X_FULL], kAbRegion);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(R1 + kb * kAbBlk, &p.x, &bars[B16_X_FULL], kb * 64, row0);
      // 32 weight blocks in the order the MMA thread consumes them: Wq, Wk, Wv, Wp, each as two 128-row halves x 4 k-blocks
      unsigned long long uses = 0;  // 3 bits per slot: how many blocks it has held so far (at most 5: the ring stages)
      for (int j = 0; j < 32; ++j) {
        const int g = j >> 3, h = (j >> 2) & 1, kb = j & 3, sl = a16_slot(j);
        const uint32_t u = static_cast<uint32_t>(uses >> (3 * sl)) & 7u;
        if (j == 24) mbar_wait(&bars[B16_O_DONE], 0);  // V^T (R1) and P (R2) have been read
        if (u > 0) mbar_wait(&bars[B16_SLOT_EMPTY0 + sl], (u - 1) & 1);
        uses += 1ull << (3 * sl);
        uint8_t* dst = sl < 2 ? ring + sl * kAbBlk : (sl < 6 ? R2 + (sl - 2) * kAbBlk : (sl < 10 ? R3 + (sl - 6) * kAbBlk : R1 + (sl - 10) * kAbBlk));
        mbar_expect_tx(&bars[B16_SLOT_FULL0 + sl], kAbBlk);
        tma_load_2d(dst, g < 3 ? &p.wqkv : &p.wproj, &bars[B16_SLOT_FULL0 + sl], kb * 64, (g < 3 ? g * kAbC : 0) + h * 128);
      }
    } else if (lane == 1) {
      // the weights do not depend on the previous kernel: ask for Wv and Wp (the blocks that are loaded late) in L2 now
      for (int j = 16; j < 32; ++j)
        tma_prefetch_l2_2d((j >> 3) < 3 ? &p.wqkv : &p.wproj, (j & 3) * 64, ((j >> 3) < 3 ? (j >> 3) * kAbC : 0) + ((j >> 2) & 1) * 128);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc128 = umma_idesc_bf16(128, 128), idesc256 = umma_idesc_bf16(128, 256);
      int j = 0;
      // one streamed weight block against one resident 64-wide k-block: weights are the B operand (Q, K) or the A operand
      unsigned long long uses = 0;
      auto ring_block = [&](uint32_t tm, const uint8_t* resident, bool weights_are_a, int kb) {
        const int sl = a16_slot(j);
        const uint32_t u = static_cast<uint32_t>(uses >> (3 * sl)) & 7u;
        uses += 1ull << (3 * sl);
        const uint8_t* src = sl < 2 ? ring + sl * kAbBlk : (sl < 6 ? R2 + (sl - 2) * kAbBlk : (sl < 10 ? R3 + (sl - 6) * kAbBlk : R1 + (sl - 10) * kAbBlk));
        mbar_wait(&bars[B16_SLOT_FULL0 + sl], u & 1);
        tc_fence_after();
        const uint64_t wd = umma_desc_sw128(smem_u32(src)), rd = umma_desc_sw128(smem_u32(resident + kb * kAbBlk));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm, (weights_are_a ? wd : rd) + 2 * k, (weights_are_a ? rd : wd) + 2 * k, idesc128, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&bars[B16_SLOT_EMPTY0 + sl]);
        ++j;
      };
      mbar_wait(&bars[B16_H_READY], 0);
      tc_fence_after();
      ab_trace(p.trace, 1, 0);
      for (int g = 0; g < 2; ++g) {  // Q -> columns [0, 256), K -> [256, 512)
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < 4; ++kb) ring_block(T + g * 256 + h * 128, R1, false, kb);
        umma_commit(&bars[g == 0 ? B16_Q_DONE : B16_K_DONE]);
        ab_trace(p.trace, 1, 1 + g);
      }
      mbar_wait(&bars[B16_QK_DRAINED], 0);
      tc_fence_after();
      ab_trace(p.trace, 1, 3);
      for (int a = 0; a < 2; ++a)  // V^T halves -> columns [0, 128), [128, 256)
        for (int kb = 0; kb < 4; ++kb) ring_block(T + a * 128, R1, true, kb);
      umma_commit(&bars[B16_V_DONE]);
      ab_trace(p.trace, 1, 4);
      for (int kb = 0; kb < 4; ++kb) {  // S -> columns [256, 384)
        const uint64_t ad = umma_desc_sw128(smem_u32(R2 + kb * kAbBlk)), bd = umma_desc_sw128(smem_u32(R3 + kb * kAbBlk));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(T + 256, ad + 2 * k, bd + 2 * k, idesc128, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(&bars[B16_S_DONE]);
      mbar_wait(&bars[B16_PV_READY], 0);
      tc_fence_after();
      ab_trace(p.trace, 1, 5);
      for (int kb = 0; kb < 2; ++kb) {  // O -> columns [0, 256): P blocks of 64 keys x V^T blocks [256 ch][64 keys]
        const uint64_t ad = umma_desc_sw128(smem_u32(R2 + kb * kAbBlk)), bd = umma_desc_sw128(smem_u32(R1 + kb * 2 * kAbBlk));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(T, ad + 2 * k, bd + 2 * k, idesc256, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(&bars[B16_O_DONE]);
      mbar_wait(&bars[B16_OS_READY], 0);
      tc_fence_after();
      ab_trace(p.trace, 1, 6);
      for (int a = 0; a < 2; ++a)  // out^T halves -> columns [256, 384), [384, 512)
        for (int kb = 0; kb < 4; ++kb) ring_block(T + 256 + a * 128, R3, true, kb);
      umma_commit(&bars[B16_D_DONE]);
      ab_trace(p.trace, 1, 7);
    }
  } else {
    const int wq = warp & 3, half = (warp - 2) >> 2, row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int wt = static_cast<int>(threadIdx.x) - 64;
    const int odd = lane & 1;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    pdl_wait();
    if (wt == 0) ab_trace(p.trace, 0, 0);
    row_part[wt] = __ldg(p.bias_qkv + wt);
    // ---- GroupNorm of the tile in place: thread = (image of the group, 16-byte unit column), the image's 16 token rows ----
    {
      const int gi = wt >> 5, cu = wt & 31, kb = cu >> 3, u = cu & 7;
      float a[8], b[8];
      const bool valid = img0 + gi < p.n;
      if (valid) ab_coeff8(p.gn_ab, p.stats_in, p.gamma, p.beta, p.cpg, p.eps, 16, kAbC, img0 + gi, kb * 64 + u * 8, a, b);
      mbar_wait(&bars[B16_X_FULL], 0);
      if (wt == 0) ab_trace(p.trace, 0, 1);
      if (valid) {
        uint8_t* blk = R1 + kb * kAbBlk;
#pragma unroll
        for (int tb = 0; tb < 16; tb += 4) {
          uint4 v[4];
          uint8_t* addr[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int r = gi * 16 + tb + t;
            addr[t] = blk + r * 128 + ((u ^ (r & 7)) << 4);
            v[t] = *reinterpret_cast<const uint4*>(addr[t]);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float lo, hi;
              unpack_bf16x2(w[e], lo, hi);
              w[e] = pack_bf16x2(fmaf(a[2 * e], lo, b[2 * e]), fmaf(a[2 * e + 1], hi, b[2 * e + 1]));
            }
            *reinterpret_cast<uint4*>(addr[t]) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&bars[B16_H_READY]);
      if (wt == 0) ab_trace(p.trace, 0, 2);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");  // the Q biases are staged
    // ---- Q: columns [0, 256) -> R2 (+ bias), K: [256, 512) -> R3 (no bias: the softmax over an image's keys cancels it).
    // Both after the K product: R2 / R3 stage Wq / Wk blocks until then ----
    mbar_wait(&bars[B16_K_DONE], 0);
    tc_fence_after();
    if (wt == 0) ab_trace(p.trace, 0, 3);
    ab_drain_colbias<4>(T + lane_off, half * 128, row, R2, row_part, nullptr, -1);
    ab_drain<3, 4>(T + 256 + lane_off, half * 128, row, R3, kAbBlk, 0.f, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&bars[B16_QK_DRAINED]);
    if (wt == 0) ab_trace(p.trace, 0, 4);
    // ---- V^T: two accumulators [128 channels][128 rows] -> R1 as two k-blocks [256 channel rows][64 keys] ----
    mbar_wait(&bars[B16_V_DONE], 0);
    tc_fence_after();
    if (wt == 0) ab_trace(p.trace, 0, 5);
#pragma unroll 1
    for (int a = 0; a < 2; ++a) {
      const int ch = a * 128 + row;
      const float bias_v = __ldg(p.bias_qkv + 2 * kAbC + ch);
      uint8_t* prow = R1 + half * 2 * kAbBlk + ch * 128;  // this half's 64 keys = k-block `half`
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        uint32_t v[32];
        tmem_ld32(T + a * 128 + lane_off + half * 64 + 32 * i, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) f[jx] = __uint_as_float(v[jx]) + bias_v;
        ab_store_units(prow, 4 * i, ch & 7, f);
      }
    }
    // ---- softmax: a row's scores are the 16 columns of its own image; P = zeros elsewhere.  The thread of the half that
    // holds the image's key block computes them, the other one writes zeros ----
    if (wt == 0) ab_trace(p.trace, 0, 6);
    mbar_wait(&bars[B16_S_DONE], 0);
    tc_fence_after();
    if (wt == 0) ab_trace(p.trace, 0, 7);
    {
      uint8_t* prow = R2 + half * kAbBlk + row * 128;
      float z[32];
#pragma unroll
      for (int jx = 0; jx < 32; ++jx) z[jx] = 0.f;
      if ((wq >> 1) == half) {
        // the warp's 32 rows are two images: columns [32 wq, 32 wq + 32) hold both diagonal blocks
        uint32_t v[32];
        tmem_ld32(T + 256 + lane_off + 32 * wq, v);
        tmem_ld_wait();
        float sc[16];
#pragma unroll
        for (int jx = 0; jx < 16; ++jx) sc[jx] = __uint_as_float(lane < 16 ? v[jx] : v[16 + jx]);
        float mx = sc[0];
#pragma unroll
        for (int jx = 1; jx < 16; ++jx) mx = fmaxf(mx, sc[jx]);
        const float sl = p.scale_log2e, mxs = mx * sl;
        float sum = 0.f;
        float e[32];
#pragma unroll
        for (int jx = 0; jx < 16; ++jx) {
          const float ev = exp2f(fmaf(sc[jx], sl, -mxs));
          sum += ev;
          e[jx] = lane < 16 ? ev : 0.f;       // keys 32 (wq & 1) + [0, 16) of the block: the first image of the warp
          e[16 + jx] = lane < 16 ? 0.f : ev;  // ... + [16, 32): the second
        }
        row_part[row] = sum;
        ab_store_units(prow, 4 * (wq & 1), row & 7, e);
        ab_store_units(prow, 4 * ((wq & 1) ^ 1), row & 7, z);
      } else {
        ab_store_units(prow, 0, row & 7, z);
        ab_store_units(prow, 4, row & 7, z);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&bars[B16_PV_READY]);
    if (wt == 0) ab_trace(p.trace, 0, 8);
    asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");
    const float inv = 1.0f / row_part[row];
    // ---- O / rowsum: columns [0, 256) -> R3 ----
    mbar_wait(&bars[B16_O_DONE], 0);
    tc_fence_after();
    if (wt == 0) ab_trace(p.trace, 0, 9);
    ab_drain<2, 4>(T + lane_off, half * 128, row, R3, kAbBlk, inv, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&bars[B16_OS_READY]);
    if (wt == 0) ab_trace(p.trace, 0, 10);
    // ---- out^T: two accumulators [128 channels][128 rows]; this thread: rows [64 half, +64) of both ----
    {
      uint32_t av[64];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int chp = (a * 128 + row) & ~1;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int r = row0 + half * 64 + 2 * i + odd;
          av[a * 32 + i] = r < nrows ? __ldg(reinterpret_cast<const uint32_t*>(p.xres + static_cast<long long>(r) * kAbC + chp)) : 0u;
        }
      }
      mbar_wait(&bars[B16_D_DONE], 0);
      tc_fence_after();
      if (wt == 0) ab_trace(p.trace, 0, 11);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int och = a * 128 + row, chp = och & ~1;
        const float bias_lo = __ldg(p.bias_proj + chp), bias_hi = __ldg(p.bias_proj + chp + 1);
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int tok0 = half * 64 + ci * 32;  // two images: rows [tok0, +16) and [tok0 + 16, +16)
          uint32_t v[32];
          tmem_ld32(T + 256 + a * 128 + lane_off + tok0, v);
          tmem_ld_wait();
          float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? v[2 * i] : v[2 * i + 1], 1);
            const float dlo = __uint_as_float(odd ? recv : v[2 * i]), dhi = __uint_as_float(odd ? v[2 * i + 1] : recv);
            float xl, xh;
            unpack_bf16x2(av[a * 32 + ci * 16 + i], xl, xh);
            const uint32_t uo = pack_bf16x2(xl + (dlo + bias_lo), xh + (dhi + bias_hi));
            const int r = row0 + tok0 + 2 * i + odd;
            if (r < nrows) *reinterpret_cast<uint32_t*>(p.out + static_cast<long long>(r) * kAbC + chp) = uo;
            float lo, hi;
            unpack_bf16x2(uo, lo, hi);
            s1[i >> 3] += lo + hi;
            s2[i >> 3] = fmaf(lo, lo, fmaf(hi, hi, s2[i >> 3]));
          }
          if (p.stats) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float t1 = s1[g], t2 = s2[g];
              t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
              t1 += __shfl_xor_sync(0xffffffffu, t1, 2); t2 += __shfl_xor_sync(0xffffffffu, t2, 2);
              const int img = img0 + ((tok0 + 16 * g) >> 4);
              if ((lane & 3) == 0 && img < p.n) {
                unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                         (static_cast<long long>(img) * (kAbC >> 2) + (och >> 2)) * 2;
                atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(t1 * kFix)));
                atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(t2 * kFix)));
              }
            }
          }
        }
      }
    }
  }

  if (threadIdx.x == 64) ab_trace(p.trace, 0, 12);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

bool attn_block_supported(int act_dtype, int heads, int L, int c) {
  return act_dtype == DMME_BF16 && heads == 1 && ((L == kAbL && (c == 256 || c == 128)) || (L == 16 && c == kAbC));
}

struct AbNorm { const float* gn_ab; const long long* stats_in; const float* gamma; const float* beta; int cpg; float eps; };

static int attn_block16_forward(const void* x, const AbNorm& nm, const void* wqkv, const float* bias_qkv, const void* wproj,
                                const float* bias_proj, int n, float scale, void* out, long long* stats, cudaStream_t stream) {
  AttnBlock16Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  uint64_t dims[2] = {(uint64_t)kAbC, (uint64_t)n * 16};
  uint64_t strides[1] = {(uint64_t)kAbC * 2};
  uint32_t box[2] = {64u, 128u};
  if ((rc = encode_map(&p.x, x, 2, dims, strides, box))) return rc;
  dims[1] = 3 * kAbC;
  if ((rc = encode_map(&p.wqkv, wqkv, 2, dims, strides, box))) return rc;
  dims[1] = kAbC;
  if ((rc = encode_map(&p.wproj, wproj, 2, dims, strides, box))) return rc;
  p.n = n;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.gn_ab = nm.gn_ab; p.stats_in = nm.stats_in; p.gamma = nm.gamma; p.beta = nm.beta; p.cpg = nm.cpg; p.eps = nm.eps;
  p.bias_qkv = bias_qkv;
  p.bias_proj = bias_proj;
  p.xres = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.stats = stats;
  p.trace = g_ab_trace;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_block16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmem);
    if (e != cudaSuccess) { set_error("attn_block16: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaError_t e = launch_pdl(attn_block16_kernel, dim3((n + 7) / 8), dim3(kA16Threads), kAbSmem, stream, p);
  return check_launch_err(e, "attn_block16_kernel");
}

template <int C>
static int attn_block_launch(const AttnBlockParams& p, int n, cudaStream_t stream) {
  constexpr int smem = 2 * 256 * C + kAbRegion + 2 * kAbBlk + 1024;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_block_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("attn_block: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const int pairs = n < device_sm_count() / 2 ? n : device_sm_count() / 2;
  cudaError_t e = launch_pdl_pair(attn_block_kernel<C>, dim3(2 * pairs), dim3(kAbThreads), smem, stream, p);
  return check_launch_err(e, "attn_block_kernel");
}

int attn_block_forward(const void* x, const AbNorm& nm, const void* wqkv, const float* bias_qkv, const void* wproj,
                       const float* bias_proj, int n, int L, int c, float scale, void* out, long long* stats, cudaStream_t stream) {
  if (L == 16) return attn_block16_forward(x, nm, wqkv, bias_qkv, wproj, bias_proj, n, scale, out, stats, stream);
  AttnBlockParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)c, (uint64_t)kAbL, (uint64_t)n};
    uint64_t strides[2] = {(uint64_t)c * 2, (uint64_t)kAbL * c * 2};
    uint32_t box[3] = {64u, 128u, 1u};
    if ((rc = encode_map(&p.x, x, 3, dims, strides, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)c, (uint64_t)3 * c};
    uint64_t strides[1] = {(uint64_t)c * 2};
    uint32_t box_half[2] = {64u, (uint32_t)c / 2}, box_a[2] = {64u, 128u};
    if ((rc = encode_map(&p.wqkv, wqkv, 2, dims, strides, box_half))) return rc;
    if ((rc = encode_map(&p.wqkv_a, wqkv, 2, dims, strides, box_a))) return rc;
    dims[1] = c;
    if ((rc = encode_map(&p.wproj, wproj, 2, dims, strides, box_a))) return rc;
  }
  p.n = n;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.gn_ab = nm.gn_ab; p.stats_in = nm.stats_in; p.gamma = nm.gamma; p.beta = nm.beta; p.cpg = nm.cpg; p.eps = nm.eps;
  p.bias_qkv = bias_qkv;
  p.bias_proj = bias_proj;
  p.xres = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.stats = stats;
  p.trace = g_ab_trace;
  return c == 256 ? attn_block_launch<256>(p, n, stream) : attn_block_launch<128>(p, n, stream);
}

}  // namespace dmme

extern "C" void dmme_debug_set_attn_block_trace(long long* buf) { dmme::g_ab_trace = buf; }

extern "C" int dmme_attention_block_supported(int heads, int L, int c, int act_dtype) {
  return dmme::attn_block_supported(act_dtype, heads, L, c) ? 1 : 0;
}

extern "C" int dmme_attention_block_fwd(const void* x, const float* gn_ab, const long long* stats_in, const float* gamma,
                                        const float* beta, int groups, float eps, const void* wqkv, const float* bias_qkv,
                                        const void* wproj, const float* bias_proj, int n, int heads, int L, int c, float scale,
                                        void* out, long long* stats, int act_dtype, void* stream) {
  DMME_REQUIRE(x && wqkv && bias_qkv && wproj && bias_proj && out && n > 0, DMME_E_BADARG,
               "dmme_attention_block_fwd: null pointer or empty batch");
  DMME_REQUIRE(gn_ab != nullptr || stats_in != nullptr, DMME_E_BADARG,
               "dmme_attention_block_fwd: the block's GroupNorm needs gn_ab or the producer's statistics (stats_in)");
  DMME_REQUIRE(dmme::attn_block_supported(act_dtype, heads, L, c), DMME_E_SHAPE,
               "dmme_attention_block_fwd: only bf16, one head, 256 tokens x 256 / 128 channels or 16 tokens x 256 channels (got heads=%d L=%d c=%d)", heads, L, c);
  DMME_REQUIRE(gn_ab != nullptr || (groups > 0 && c % groups == 0 && (c / groups) % 4 == 0), DMME_E_SHAPE,
               "dmme_attention_block_fwd: in-kernel GroupNorm coefficients need channels per group to be a multiple of 4");
  DMME_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wqkv) | reinterpret_cast<uintptr_t>(wproj) |
                 reinterpret_cast<uintptr_t>(gn_ab) | reinterpret_cast<uintptr_t>(bias_qkv)) & 15u) == 0,
               DMME_E_BADARG, "dmme_attention_block_fwd: x / weights / gn_ab / bias_qkv must be 16-byte aligned");
  dmme::AbNorm nm{gn_ab, stats_in, gamma, beta, groups > 0 ? c / groups : 0, eps};
  return dmme::attn_block_forward(x, nm, wqkv, bias_qkv, wproj, bias_proj, n, L, c, scale, out, stats,
                                  static_cast<cudaStream_t>(stream));
}
