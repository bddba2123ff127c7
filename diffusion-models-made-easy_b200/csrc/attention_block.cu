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
  CUtensorMap wqkv;   // [768][256] bf16 as (k, row), box [64][128]
  CUtensorMap wproj;  // [256][256] bf16
  int n;
  float scale_log2e;
  const float* gn_ab;         // [n][256][2]
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
constexpr int kAbSmem = 3 * kAbRegion + 2 * kAbBlk + 1024;

enum {
  AB_X_FULL = 0, AB_H_READY, AB_WQ_FULL, AB_WK_FULL, AB_WP_FULL, AB_WV_FULL0, AB_WV_FULL1, AB_WV_EMPTY0, AB_WV_EMPTY1,
  AB_Q_DONE, AB_K_DONE, AB_V_DONE, AB_S_DONE, AB_O_DONE, AB_D_DONE,
  AB_QS_READY, AB_KS_READY, AB_EPI_DONE,  // joined by the 2 x 256 workers
  AB_PV_READY0,                           // + key block: P block and V^T block written (2 x 128 workers of that half each)
  AB_OS_READY0 = AB_PV_READY0 + 4,        // + channel block of O
  AB_NBARS = AB_OS_READY0 + 4
};

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

__device__ __forceinline__ void ab_store_chunk(uint8_t* prow, int c, int row, const float (&f)[32]) {
  uint8_t* pc = prow + (c >> 6) * kAbBlk;
  const int u0 = (c & 63) >> 3;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    uint4 o;
    o.x = pack_bf16x2(f[8 * jj + 0], f[8 * jj + 1]);
    o.y = pack_bf16x2(f[8 * jj + 2], f[8 * jj + 3]);
    o.z = pack_bf16x2(f[8 * jj + 4], f[8 * jj + 5]);
    o.w = pack_bf16x2(f[8 * jj + 6], f[8 * jj + 7]);
    *reinterpret_cast<uint4*>(pc + (((u0 + jj) ^ (row & 7)) << 4)) = o;
  }
}

// TMEM [this warp's 32 lanes][128 columns from c_lo] fp32 -> bf16 rows of a region (four [128][64] SWIZZLE_128B blocks).
// MODE 1: + lv; 2: * lv; 3: as is.  bar_a / bar_b (shared::cluster addresses, 0 = none): arrived on after the first / second 64-column
// block has been written (a consumer MMA may start on that k-block while the other is still being drained).
template <int MODE>
__device__ __forceinline__ void ab_drain(uint32_t tmem, int c_lo, int row, uint8_t* region, float lv, uint32_t bar_a,
                                         uint32_t bar_b) {
  uint8_t* prow = region + row * 128;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int c = c_lo + 32 * i;
    uint32_t v[32];
    tmem_ld32(tmem + c, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      f[j] = MODE == 1 ? __uint_as_float(v[j]) + lv : (MODE == 2 ? __uint_as_float(v[j]) * lv : __uint_as_float(v[j]));
    ab_store_chunk(prow, c, row, f);
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
__device__ __forceinline__ void ab_drain_colbias(uint32_t tmem, int c_lo, int row, uint8_t* region, const float* sbias,
                                                 long long* trace, int tbase) {
  uint8_t* prow = region + row * 128;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
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

__global__ void __launch_bounds__(kAbThreads, 1) attn_block_kernel(const __grid_constant__ AttnBlockParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[AB_NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float row_part[2][128];  // softmax row exchange; between images: the 256 Q biases

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* R1 = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* R2 = R1 + kAbRegion;
  uint8_t* R3 = R2 + kAbRegion;
  uint8_t* ring = R3 + kAbRegion;

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
      auto load_slab = [&](uint8_t* dst, const CUtensorMap* m, int bar, int row0) {
        // weight rows [row0, row0 + 128) x 256 k: this CTA's half of an operand; both halves complete on the leader's barrier
        if (rank == 0) mbar_expect_tx(&bars[bar], 2 * kAbRegion);
        const uint32_t lead = mapa_u32(&bars[bar], 0);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_2sm(dst + kb * kAbBlk, m, lead, kb * 64, row0);
      };
      auto load_x = [&](int img) {
        mbar_expect_tx(&bars[AB_X_FULL], kAbRegion);
        for (int kb = 0; kb < 4; ++kb)
          tma_load_3d(R1 + kb * kAbBlk, &p.x, &bars[AB_X_FULL], kb * 64, static_cast<int>(rank) * 128, img);
      };
      const int r128 = static_cast<int>(rank) * 128;
      int it = 0, wv_it = 0;
      if (cid < p.n) {
        load_x(cid);
        load_slab(R2, &p.wqkv, AB_WQ_FULL, r128);
        load_slab(R3, &p.wqkv, AB_WK_FULL, kAbC + r128);
      }
      for (int img = cid; img < p.n; img += ncl, ++it) {
        const uint32_t ph = it & 1;
        for (int kb = 0; kb < 4; ++kb, ++wv_it) {
          const int s = wv_it & 1;
          mbar_wait(&bars[AB_WV_EMPTY0 + s], ((wv_it >> 1) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&bars[AB_WV_FULL0 + s], 2 * kAbBlk);
          tma_load_2d_2sm(ring + s * kAbBlk, &p.wqkv, mapa_u32(&bars[AB_WV_FULL0 + s], 0), kb * 64, 2 * kAbC + r128);
        }
        const int nxt = img + ncl;
        mbar_wait(&bars[AB_O_DONE], ph);  // P (R2) and V^T (R1) have been read
        load_slab(R2, &p.wproj, AB_WP_FULL, r128);
        if (nxt < p.n) load_x(nxt);
        if (nxt < p.n) {
          mbar_wait(&bars[AB_D_DONE], ph);  // Wproj (R2) and O (R3) have been read
          load_slab(R2, &p.wqkv, AB_WQ_FULL, r128);
          load_slab(R3, &p.wqkv, AB_WK_FULL, kAbC + r128);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, 256);
      auto gemm = [&](uint32_t tm, const uint8_t* A, const uint8_t* B) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t ad = umma_desc_sw128(smem_u32(A + kb * kAbBlk)), bd = umma_desc_sw128(smem_u32(B + kb * kAbBlk));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_2sm(tm, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
      };
      // one k-block of a product whose operand blocks are delivered in the order 0, 2, 1, 3 (see the workers)
      auto gemm_kb = [&](uint32_t tm, const uint8_t* A, const uint8_t* B, int kb, bool first) {
        const uint64_t ad = umma_desc_sw128(smem_u32(A + kb * kAbBlk)), bd = umma_desc_sw128(smem_u32(B + kb * kAbBlk));
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
        gemm(T0, R1, R2);  // Q = H Wq^T  (T0: the previous image's O was drained before its projection was issued)
        umma_commit_2sm(&bars[AB_Q_DONE], 3);
        mbar_wait(&bars[AB_WK_FULL], ph);
        if (it > 0) mbar_wait_cluster(&bars[AB_EPI_DONE], ph ^ 1);  // T1: the previous image's output has been read
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 1);
        gemm(T1, R1, R3);  // K = H Wk^T
        umma_commit_2sm(&bars[AB_K_DONE], 3);
        mbar_wait_cluster(&bars[AB_QS_READY], ph);  // T0 drained, Q in R2
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 2);
        for (int kb = 0; kb < 4; ++kb, ++wv_it) {  // V^T = Wv H^T
          const int s = wv_it & 1;
          mbar_wait(&bars[AB_WV_FULL0 + s], (wv_it >> 1) & 1);
          tc_fence_after();
          ab_trace(p.trace, 1, it * 10 + 3 + kb);
          const uint64_t ad = umma_desc_sw128(smem_u32(ring + s * kAbBlk)), bd = umma_desc_sw128(smem_u32(R1 + kb * kAbBlk));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_2sm(T0, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&bars[AB_WV_EMPTY0 + s], 3);
        }
        umma_commit_2sm(&bars[AB_V_DONE], 3);
        mbar_wait_cluster(&bars[AB_KS_READY], ph);  // T1 drained, K in R3
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 7);
        gemm(T1, R2, R3);  // S = Q K^T
        umma_commit_2sm(&bars[AB_S_DONE], 3);
        // O = P V, key blocks in the order the two worker halves deliver them.  Blocks 0 and 2 together also say that every
        // worker has drained V^T out of T0.
        mbar_wait_cluster(&bars[AB_PV_READY0 + 0], ph);
        mbar_wait_cluster(&bars[AB_PV_READY0 + 2], ph);
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 8);
        gemm_kb(T0, R2, R1, 0, true);
        gemm_kb(T0, R2, R1, 2, false);
        mbar_wait_cluster(&bars[AB_PV_READY0 + 1], ph);
        mbar_wait_cluster(&bars[AB_PV_READY0 + 3], ph);
        tc_fence_after();
        gemm_kb(T0, R2, R1, 1, false);
        gemm_kb(T0, R2, R1, 3, false);
        umma_commit_2sm(&bars[AB_O_DONE], 3);
        // out^T = Wp O^T, channel blocks of O as they are drained (T1: S was read before the last P blocks were signalled)
        mbar_wait(&bars[AB_WP_FULL], ph);
        mbar_wait_cluster(&bars[AB_OS_READY0 + 0], ph);
        mbar_wait_cluster(&bars[AB_OS_READY0 + 2], ph);
        tc_fence_after();
        ab_trace(p.trace, 1, it * 10 + 9);
        gemm_kb(T1, R2, R3, 0, true);
        gemm_kb(T1, R2, R3, 2, false);
        mbar_wait_cluster(&bars[AB_OS_READY0 + 1], ph);
        mbar_wait_cluster(&bars[AB_OS_READY0 + 3], ph);
        tc_fence_after();
        gemm_kb(T1, R2, R3, 1, false);
        gemm_kb(T1, R2, R3, 3, false);
        umma_commit_2sm(&bars[AB_D_DONE], 3);
      }
    }
  } else if (warp >= 2 + kAbWorkers / 32) {
    // =========================== GroupNorm transform warps ===========================
    // x tile -> H in place, one image ahead of the workers (the tile of image i + 1 lands once P V of image i is complete and
    // H is not needed before the workers have stored image i: two warps are enough and leave the shared-memory port to the
    // drains): thread = one 16-byte unit column (8 channels, their (a, b) in registers) x every 2nd token row
    const int t2 = static_cast<int>(threadIdx.x) - (64 + kAbWorkers);
    const int cu = t2 & 31, kb = cu >> 3, u = cu & 7, r0 = t2 >> 5;
    const uint32_t bar_h = mapa_u32(&bars[AB_H_READY], 0);
    pdl_wait();
    int it = 0;
    for (int img = cid; img < p.n; img += ncl, ++it) {
      const float4* abp = reinterpret_cast<const float4*>(p.gn_ab + (static_cast<long long>(img) * kAbC + kb * 64 + u * 8) * 2);
      float a[8], b[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t4 = __ldg(abp + j);
        a[2 * j] = t4.x; b[2 * j] = t4.y; a[2 * j + 1] = t4.z; b[2 * j + 1] = t4.w;
      }
      mbar_wait(&bars[AB_X_FULL], it & 1);
      if (t2 == 0) ab_trace(p.trace, 0, it * 14 + 0);
      uint8_t* blk = R1 + kb * kAbBlk;
#pragma unroll 2
      for (int jb = 0; jb < 128 / (kAbXform / 32); jb += 4) {
        uint4 v[4];
        uint8_t* addr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = r0 + (kAbXform / 32) * (jb + j);
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
    const int half = (warp - 2) >> 2;    // which 128 of the 256 accumulator columns
    const int row = wq * 32 + lane;      // TMEM lane = tile row
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int c_lo = half * 128;
    const uint32_t bar_qs = mapa_u32(&bars[AB_QS_READY], 0), bar_ks = mapa_u32(&bars[AB_KS_READY], 0),
                   bar_epi = mapa_u32(&bars[AB_EPI_DONE], 0),
                   bar_pv = mapa_u32(&bars[AB_PV_READY0 + 2 * half], 0),  // this half's two key blocks: + 0, + 8 bytes
                   bar_os = mapa_u32(&bars[AB_OS_READY0 + 2 * half], 0);
    const int och = static_cast<int>(rank) * 128 + row;  // this thread's channel where lanes are channels (V^T, out^T)
    const int chp = och & ~1, odd = lane & 1;            // the channel pair it stores after the epilogue's lane exchange
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    pdl_wait();
    const float bias_v = __ldg(p.bias_qkv + 2 * kAbC + och);
    const float bias_lo = __ldg(p.bias_proj + chp), bias_hi = __ldg(p.bias_proj + chp + 1);
    float* sbias = &row_part[0][0];
    const int wt = static_cast<int>(threadIdx.x) - 64;
    const float bias_q = __ldg(p.bias_qkv + wt);  // staged in row_part for every image's Q drain (the softmax reuses it)
    int it = 0;
    for (int img = cid; img < p.n; img += ncl, ++it) {
      const uint32_t ph = it & 1;
      sbias[wt] = bias_q;
      asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory");
      // ---- Q: T0 -> R2 ----
      mbar_wait(&bars[AB_Q_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 0);
      ab_drain_colbias(T0 + lane_off, c_lo, row, R2, sbias, p.trace, (warp == 2 && lane == 0) ? it * 16 + 1 : -1);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_cluster(bar_qs);
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 3);
      // ---- K: T1 -> R3.  The K bias is left out: it adds (q_i + b_q) . b_k to every score of row i, which the softmax over
      // the keys cancels exactly (the reference adds it and rounds K + b_k; this is the same function of the inputs) ----
      mbar_wait(&bars[AB_K_DONE], ph);
      tc_fence_after();
      ab_drain<3>(T1 + lane_off, c_lo, row, R3, 0.f, 0u, 0u);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_cluster(bar_ks);
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 5);
      // ---- V^T: T0 -> R1 (lanes = channels, columns = tokens = keys) ----
      mbar_wait(&bars[AB_V_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 6);
      ab_drain<1>(T0 + lane_off, c_lo, row, R1, bias_v, 0u, 0u);
      // ---- softmax of this query row (two threads per row, half of the keys each): T1 -> P in R2 ----
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 7);
      mbar_wait(&bars[AB_S_DONE], ph);
      tc_fence_after();
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 8);
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
      ab_drain<2>(T0 + lane_off, c_lo, row, R3, 1.0f / sum, bar_os, bar_os + 8u);
      if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 11);
      // ---- out^T: T1 (lane = channel, column = token) + bias + x -> global, statistics of the stored values.  Neighbouring
      // lanes exchange every second value so that a thread stores a channel PAIR of one token (4-byte accesses, half as many
      // of them: the 2-byte version spent 6 k clocks per tile in the load / store unit) ----
      {
        const long long ibase = static_cast<long long>(img) * (kAbL * kAbC) + chp;
        const uint32_t* __restrict__ xr = reinterpret_cast<const uint32_t*>(p.xres + ibase) + odd * (kAbC / 2);
        uint32_t* __restrict__ op = reinterpret_cast<uint32_t*>(p.out + ibase) + odd * (kAbC / 2);
        // all of the residual is requested before the projection is awaited: one chunk of look-ahead left an L2 round trip
        // per chunk exposed (5 - 6 k clocks per tile)
        uint32_t av[64];
        if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 5);
#pragma unroll
        for (int i = 0; i < 64; ++i) av[i] = __ldg(xr + (c_lo + 2 * i) * (kAbC / 2));
        if (warp == 2 && lane == 0) ab_trace(p.trace, 2, it * 16 + 6);
        mbar_wait(&bars[AB_D_DONE], ph);
        tc_fence_after();
        if (warp == 2 && lane == 0) ab_trace(p.trace, 0, it * 14 + 12);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int tok0 = c_lo + ci * 32;
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
            op[(tok0 + 2 * i) * (kAbC / 2)] = uo;
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
                                     (static_cast<long long>(img) * (kAbC >> 2) + (och >> 2)) * 2;
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

bool attn_block_supported(int act_dtype, int heads, int L, int c) {
  return act_dtype == DMME_BF16 && heads == 1 && L == kAbL && c == kAbC;
}

int attn_block_forward(const void* x, const float* gn_ab, const void* wqkv, const float* bias_qkv, const void* wproj,
                       const float* bias_proj, int n, float scale, void* out, long long* stats, cudaStream_t stream) {
  AttnBlockParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)kAbC, (uint64_t)kAbL, (uint64_t)n};
    uint64_t strides[2] = {(uint64_t)kAbC * 2, (uint64_t)kAbL * kAbC * 2};
    uint32_t box[3] = {64u, 128u, 1u};
    if ((rc = encode_map(&p.x, x, 3, dims, strides, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)kAbC, (uint64_t)3 * kAbC};
    uint64_t strides[1] = {(uint64_t)kAbC * 2};
    uint32_t box[2] = {64u, 128u};
    if ((rc = encode_map(&p.wqkv, wqkv, 2, dims, strides, box))) return rc;
    dims[1] = kAbC;
    if ((rc = encode_map(&p.wproj, wproj, 2, dims, strides, box))) return rc;
  }
  p.n = n;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.gn_ab = gn_ab;
  p.bias_qkv = bias_qkv;
  p.bias_proj = bias_proj;
  p.xres = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.stats = stats;
  p.trace = g_ab_trace;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmem);
    if (e != cudaSuccess) { set_error("attn_block: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const int pairs = n < device_sm_count() / 2 ? n : device_sm_count() / 2;
  cudaError_t e = launch_pdl_pair(attn_block_kernel, dim3(2 * pairs), dim3(kAbThreads), kAbSmem, stream, p);
  return check_launch_err(e, "attn_block_kernel");
}

}  // namespace dmme

extern "C" void dmme_debug_set_attn_block_trace(long long* buf) { dmme::g_ab_trace = buf; }

extern "C" int dmme_attention_block_supported(int heads, int L, int c, int act_dtype) {
  return dmme::attn_block_supported(act_dtype, heads, L, c) ? 1 : 0;
}

extern "C" int dmme_attention_block_fwd(const void* x, const float* gn_ab, const void* wqkv, const float* bias_qkv,
                                        const void* wproj, const float* bias_proj, int n, int heads, int L, int c, float scale,
                                        void* out, long long* stats, int act_dtype, void* stream) {
  DMME_REQUIRE(x && gn_ab && wqkv && bias_qkv && wproj && bias_proj && out && n > 0, DMME_E_BADARG,
               "dmme_attention_block_fwd: null pointer or empty batch");
  DMME_REQUIRE(dmme::attn_block_supported(act_dtype, heads, L, c), DMME_E_SHAPE,
               "dmme_attention_block_fwd: only bf16, one head, 256 tokens x 256 channels (got heads=%d L=%d c=%d)", heads, L, c);
  DMME_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wqkv) | reinterpret_cast<uintptr_t>(wproj) |
                 reinterpret_cast<uintptr_t>(gn_ab) | reinterpret_cast<uintptr_t>(bias_qkv)) & 15u) == 0,
               DMME_E_BADARG, "dmme_attention_block_fwd: x / weights / gn_ab / bias_qkv must be 16-byte aligned");
  return dmme::attn_block_forward(x, gn_ab, wqkv, bias_qkv, wproj, bias_proj, n, scale, out, stats,
                                  static_cast<cudaStream_t>(stream));
}
