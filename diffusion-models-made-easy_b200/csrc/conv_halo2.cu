// "Wide" variant of conv_halo.cu for the 32x32 and 16x16 levels: ONE weight tile feeds TWO position tiles (all 512 TMEM
// columns).  Kept as a selectable kernel (DMME_CONV_HALO2) and as the record of a measured negative result; AUTO does
// not use it.
//
// Hypothesis: conv_halo_kernel streams a [128 cout][64] weight tile (16 KB) from L2 for every 4 MMAs, ~41-43 B/clk/SM,
// and would be bound by the L2 -> SM feed; using each weight tile for 8 MMAs halves that.  Measurement (B200, batch 256,
// tools/prof_conv.py): 128->128@32 79.8 us vs 67.6 us for conv_halo, 256->256@16 82.0 vs 71.7 us -- slower.  The ncu
// capture of conv_halo (profiles/r1_conv_halo_ncu_full.txt) shows why: L2 -> SM traffic is 465 MB per launch at only
// 19.5% of the LTS peak, so the feed was never the limit.  The limit is the SM's shared-memory port: an SS-mode
// tcgen05.mma with M = 128, N = 240, K = 16 reads (128 + 240) x 32 B = 11.8 KB per 120 clocks = 98 B/clk of the 128 B/clk
// port, TMA writes add ~25 B/clk, and the row-shifted (not 1024-byte aligned) operand starts cost extra.  Two position
// tiles per weight tile do not change bytes per MMA, and the single (undoubled) accumulator exposes the epilogue drain.
// The way past ~65% tensor-pipe activity is cta_group::2 (each CTA then reads only half of the N-side operand), which
// needs M = 256 output channels -- a next-round item for the 256-channel levels.
//
// Everything else (padded-row position space, tap = descriptor shift, transposed accumulators, epilogue) is
// conv_halo.cu's scheme.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct ConvHalo2Params {
  CUtensorMap a[4];  // src0, src1, res0, res1: box = one padded row [W+2 px][64 ch]
  CUtensorMap b;     // weights [cout][K] bf16, box [128][64]
  int chunks0, chunks1, rchunks0, rchunks1;
  int n, h, wp;
  int rt;            // padded rows per tile
  int n_mma;         // MMA N: rt * wp rounded up to a multiple of 16
  int total_rows;    // n * (h + 2)
  int imgs_per_tile; // > 0: tiles hold whole padded images (small resolutions), one TMA box per image
  int m_tiles, n_tiles;   // m_tiles counts PAIRS of position tiles
  const float* bias;
  const float* temb;
  int temb_rows, temb_ld;
  const __nv_bfloat16* addend;
  __nv_bfloat16* out;
  long long* stats;
};

constexpr int kHalo2BN = 128;            // output channels per unit (MMA M)
constexpr int kHalo2Cols = 256;          // TMEM columns per accumulator stage
constexpr int kHalo2ASlot = 40 * 1024;   // one tile: >= ((rt + 2) * (W+2) + 1) * 128 bytes; a stage holds two
constexpr int kHalo2BSlot = kHalo2BN * 128;
constexpr int kHalo2AStages = 2;
constexpr int kHalo2BStages = 4;
constexpr int kHalo2Smem = kHalo2AStages * 2 * kHalo2ASlot + kHalo2BStages * kHalo2BSlot + 1024;
constexpr int kHalo2EpiWarps = 8;
constexpr int kHalo2Threads = (kHalo2EpiWarps + 3) * 32;
constexpr int kW2ProdA = kHalo2EpiWarps, kW2ProdB = kHalo2EpiWarps + 1, kW2Mma = kHalo2EpiWarps + 2;

template <int W, int COUT>
__global__ void __launch_bounds__(kHalo2Threads, 1) conv_halo2_kernel(const __grid_constant__ ConvHalo2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kHalo2AStages], a_empty[kHalo2AStages];
  __shared__ __align__(8) uint64_t b_full[kHalo2BStages], b_empty[kHalo2BStages];
  __shared__ __align__(8) uint64_t acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;

  constexpr int WP = W + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* abuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* bbuf = abuf + kHalo2AStages * 2 * kHalo2ASlot;

  const int cchunks = p.chunks0 + p.chunks1;
  const int nck = cchunks + p.rchunks0 + p.rchunks1;
  const int units = p.m_tiles * p.n_tiles;
  constexpr int kRowBytes = WP * 128;
  const int nr = p.rt + 2;  // halo rows: one above and one below the tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHalo2AStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kHalo2BStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, kHalo2EpiWarps * 32);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == kW2ProdA && lane == 0) {
    tma_prefetch_desc(&p.a[0]);
    if (p.chunks1) tma_prefetch_desc(&p.a[1]);
    if (p.rchunks0) tma_prefetch_desc(&p.a[2]);
    if (p.rchunks1) tma_prefetch_desc(&p.a[3]);
    tma_prefetch_desc(&p.b);
  }
  if (warp == kW2Mma) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == kW2ProdA) {
    // =========================== halo-tile producer ===========================
    if (lane == 0) {
      int a_it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int mp = u / p.n_tiles;
        for (int ck = 0; ck < nck; ++ck, ++a_it) {
          int which, cc;
          if (ck < cchunks) {
            which = ck < p.chunks0 ? 0 : 1;
            cc = (which ? ck - p.chunks0 : ck) * 64;
          } else {
            const int rk = ck - cchunks;
            which = rk < p.rchunks0 ? 2 : 3;
            cc = (which == 3 ? rk - p.rchunks0 : rk) * 64;
          }
          const int as = a_it % kHalo2AStages;
          mbar_wait(&a_empty[as], ((a_it / kHalo2AStages) & 1) ^ 1);
          mbar_expect_tx(&a_full[as], 2 * nr * kRowBytes);
          for (int h = 0; h < 2; ++h) {
            const int pr0 = (2 * mp + h) * p.rt - 1;  // first halo row of this tile (padded-row index, may be -1)
            // slot layout: 128 bytes of slack (tap (-1,-1) of position 0 reaches one row back), then the halo rows
            uint8_t* dst = abuf + (as * 2 + h) * kHalo2ASlot + 128;
            for (int i = 0; i < nr; ++i) {
              const int pr = pr0 + i;
              int ni, yy;
              if (pr < 0) { ni = -1; yy = 0; }             // before the first image: whole row out of bounds -> zeros
              else { ni = pr / (p.h + 2); yy = pr - ni * (p.h + 2) - 1; }  // yy = -1 or h: padding row; ni >= n -> zeros
              tma_load_5d(dst + i * kRowBytes, &p.a[which], &a_full[as], cc, -1, 0, yy, ni);
            }
          }
        }
      }
    }
  } else if (warp == kW2ProdB) {
    // =========================== weight-tile producer ===========================
    if (lane == 0) {
      int b_it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int col0 = (u % p.n_tiles) * kHalo2BN;
        for (int ck = 0; ck < nck; ++ck) {
          const bool is_conv = ck < cchunks;
          const int ntaps = is_conv ? 9 : 1;
          const int kb0 = is_conv ? ck : 9 * cchunks + (ck - cchunks);
          const int kbs = is_conv ? cchunks : 0;
          for (int tap = 0; tap < ntaps; ++tap, ++b_it) {
            const int bs = b_it % kHalo2BStages;
            mbar_wait(&b_empty[bs], ((b_it / kHalo2BStages) & 1) ^ 1);
            mbar_expect_tx(&b_full[bs], kHalo2BSlot);
            tma_load_2d(bbuf + bs * kHalo2BSlot, &p.b, &b_full[bs], (kb0 + tap * kbs) * 64, col0);
          }
        }
      }
    }
  } else if (warp == kW2Mma) {
    // =========================== MMA issuer ===========================
    // the whole warp walks the loop (converged waits); one elected lane issues the MMAs and their commits
    const uint32_t idesc = umma_idesc_bf16(kHalo2BN, p.n_mma);  // M = 128 output channels, N = positions of the tile
    int a_it = 0, b_it = 0, u_it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++u_it) {
      mbar_wait(&acc_empty, (u_it & 1) ^ 1);  // the epilogue has drained the previous unit
      tc_fence_after();
      for (int ck = 0; ck < nck; ++ck, ++a_it) {
        const bool is_conv = ck < cchunks;
        const int ntaps = is_conv ? 9 : 1;
        const int as = a_it % kHalo2AStages;
        mbar_wait(&a_full[as], (a_it / kHalo2AStages) & 1);
        tc_fence_after();
        // position 0 of a tile = first pixel slot of the tile's first row = halo row 1 of its slot
        const uint32_t x0_addr = smem_u32(abuf + as * 2 * kHalo2ASlot) + 128u + kRowBytes;
        for (int tap = 0; tap < ntaps; ++tap, ++b_it) {
          const int bs = b_it % kHalo2BStages;
          mbar_wait(&b_full[bs], (b_it / kHalo2BStages) & 1);
          tc_fence_after();
          const int d = is_conv ? (tap / 3 - 1) * WP + (tap % 3 - 1) : 0;
          if (elect_one()) {
            const uint64_t wdesc = umma_desc_sw128(smem_u32(bbuf + bs * kHalo2BSlot));  // [128 cout][64]: M side, used twice
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t xdesc = umma_desc_sw128(x0_addr + static_cast<uint32_t>(h * kHalo2ASlot + d * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + h * kHalo2Cols, wdesc + 2 * k, xdesc + 2 * k, idesc, (ck | tap | k) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[bs]);
            if (tap == ntaps - 1) {
              umma_commit(&a_empty[as]);
              if (ck == nck - 1) umma_commit(&acc_full);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    // thread = one output channel (TMEM lane); warp (q, half) owns channels [32q, 32q+32) and every second row of the
    // unit's 2 * rt rows (tile 0 in TMEM columns [0, 256), tile 1 in [256, 512))
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    const bool temb_per_image = p.temb && p.temb_rows != 1;
    int u_it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++u_it) {
      const int mt = u / p.n_tiles, nt = u - mt * p.n_tiles;
      const int ch = nt * kHalo2BN + q * 32 + lane;
      const float bias_c = p.bias ? __ldg(p.bias + ch) : 0.f;
      float s1 = 0.f, s2 = 0.f, bt = bias_c;
      int cur_n = -1;

      auto flush_stats = [&]() {  // warp-uniform call: per-image sums of this lane's channel -> micro-group atomics
        if (p.stats && cur_n >= 0) {
          float a1 = s1, a2 = s2;
          a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
          if ((lane & 3) == 0) {
            unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                     (static_cast<long long>(cur_n) * (COUT >> 2) + (ch >> 2)) * 2;
            atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(a1 * kFix)));
            atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(a2 * kFix)));
          }
        }
        s1 = 0.f; s2 = 0.f;
      };

      mbar_wait(&acc_full, u_it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int rr2 = half; rr2 < 2 * p.rt; rr2 += 2) {  // the unit's two tiles are adjacent: rows mt*2*rt + [0, 2 rt)
        const int h = rr2 >= p.rt ? 1 : 0, rr = rr2 - h * p.rt;
        const int pr = mt * 2 * p.rt + rr2;
        if (pr >= p.total_rows) break;
        const int n = pr / (p.h + 2);
        const int yy = pr - n * (p.h + 2) - 1;
        if (yy < 0 || yy >= p.h) continue;  // padding row: junk accumulator columns
        if (n != cur_n) {
          flush_stats();
          cur_n = n;
          bt = bias_c;
          if (p.temb) bt += __ldg(p.temb + static_cast<long long>(temb_per_image ? n : 0) * p.temb_ld + ch);
        }
        const long long o = (static_cast<long long>(n) * p.h + yy) * W * COUT + ch;
        __nv_bfloat16* op = p.out + o;
        const uint32_t taddr = tmem_base + lane_off + static_cast<uint32_t>(h * kHalo2Cols + rr * WP + 1);
        // the addend row is fetched before the accumulator row is awaited, and every load is issued before the first
        // store: out and addend may alias as far as the compiler knows, and a load placed after a store waits for it
        // (measured: 550 us instead of 71 us per launch)
        float av[W];
        if (p.addend) {
          const __nv_bfloat16* __restrict__ ap = p.addend + o;
#pragma unroll
          for (int i = 0; i < W; ++i) av[i] = __bfloat162float(__ldg(ap + i * COUT));
        } else {
#pragma unroll
          for (int i = 0; i < W; ++i) av[i] = 0.f;
        }
        uint32_t v[W];
        if constexpr (W == 32) tmem_ld32(taddr, v);
        else if constexpr (W == 16) tmem_ld16(taddr, v);
        else if constexpr (W == 8) tmem_ld8(taddr, v);
        else tmem_ld4(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < W; ++i) {
          const float val = __uint_as_float(v[i]) + bt + av[i];
          const __nv_bfloat16 r = __float2bfloat16_rn(val);
          op[i * COUT] = r;
          const float rf = __bfloat162float(r);
          s1 += rf;
          s2 = fmaf(rf, rf, s2);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty);
      flush_stats();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kW2Mma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int g_sm_count2 = 0;

bool conv_halo2_supported(const dmme_conv_desc& d) {
  if (d.act_dtype != DMME_BF16 || d.in_layout != DMME_IN_NHWC || d.out_layout != DMME_OUT_NHWC) return false;
  if (d.upsample || d.ksize != 3 || d.stride != 1) return false;
  if (d.c0 <= 0 || d.c0 % 64 || d.c1 % 64 || d.rc0 % 64 || d.rc1 % 64) return false;
  if (d.cout != 128 && d.cout != 256) return false;
  if ((d.w_in != 16 && d.w_in != 32) || d.h_in != d.w_in) return false;
  if (static_cast<long long>(d.n) * (d.h_in + 2) > (1 << 24)) return false;
  return true;
}

template <int W, int COUT>
static int launch_halo2(const ConvHalo2Params& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo2_kernel<W, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHalo2Smem);
    if (e != cudaSuccess) {
      set_error("conv_halo2: cudaFuncSetAttribute(%d bytes): %s", kHalo2Smem, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int units = p.m_tiles * p.n_tiles;
  const int grid = units < g_sm_count2 ? units : g_sm_count2;
  conv_halo2_kernel<W, COUT><<<grid, kHalo2Threads, kHalo2Smem, stream>>>(p);
  return check_launch("conv_halo2_kernel");
}

int conv_halo2_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_halo2_supported(d), DMME_E_SHAPE, "conv_halo2: unsupported shape/layout");
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_halo2: null src0/weight/out");
  DMME_REQUIRE(d.c1 == 0 || d.src1, DMME_E_BADARG, "conv_halo2: c1 > 0 but src1 is null");
  DMME_REQUIRE(d.rc0 == 0 || d.res0, DMME_E_BADARG, "conv_halo2: rc0 > 0 but res0 is null");
  DMME_REQUIRE(d.rc1 == 0 || d.res1, DMME_E_BADARG, "conv_halo2: rc1 > 0 but res1 is null");
  if (g_sm_count2 == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count2, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count2 <= 0) g_sm_count2 = 148;
  }
  ConvHalo2Params p;
  memset(&p, 0, sizeof(p));
  p.chunks0 = d.c0 / 64; p.chunks1 = d.c1 / 64; p.rchunks0 = d.rc0 / 64; p.rchunks1 = d.rc1 / 64;
  p.n = d.n; p.h = d.h_in; p.wp = d.w_in + 2;
  p.total_rows = d.n * (d.h_in + 2);
  p.n_tiles = d.cout / kHalo2BN;
  p.imgs_per_tile = 0;
  // rows per tile: at most what fits 256 accumulator columns (7 rows of 34, 14 of 18); fewer when that evens out the
  // static schedule (cost ~ waves x (MMA columns of the unit + the non-overlapped drain, ~ 110 columns' worth))
  {
    const int rt_max = kHalo2Cols / p.wp;
    long long best_cost = -1;
    for (int rt = rt_max; rt >= (rt_max + 1) / 2; --rt) {
      const int n_mma = ((rt * p.wp + 15) / 16) * 16;
      const long long pairs = (p.total_rows + 2 * rt - 1) / (2 * rt);
      const long long units = pairs * p.n_tiles;
      const long long cost = ((units + g_sm_count2 - 1) / g_sm_count2) * (2 * n_mma + 110);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; p.rt = rt; p.n_mma = n_mma; }
    }
  }
  p.m_tiles = (p.total_rows + 2 * p.rt - 1) / (2 * p.rt);
  p.bias = d.bias; p.temb = d.temb; p.temb_rows = d.temb_rows; p.temb_ld = d.temb_ld;
  p.addend = static_cast<const __nv_bfloat16*>(d.addend);
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.stats = d.stats;
  // the MMA reads n_mma + (W+3) position rows past the first tile position; keep that inside the slot
  DMME_REQUIRE((1 + (p.rt + 2) * p.wp) * 128 <= kHalo2ASlot && (1 + 2 * p.wp + 1 + p.n_mma) * 128 <= kHalo2ASlot,
               DMME_E_SHAPE, "conv_halo2: halo tile does not fit its shared-memory slot");

  auto act_map = [&](CUtensorMap* m, const void* ptr, int c) -> int {
    uint64_t dims[5] = {(uint64_t)c, (uint64_t)d.w_in, 1, (uint64_t)d.h_in, (uint64_t)d.n};
    uint64_t strides[4] = {(uint64_t)c * 2, (uint64_t)d.w_in * c * 2, (uint64_t)d.w_in * c * 2,
                           (uint64_t)d.h_in * d.w_in * c * 2};
    uint32_t box[5] = {64u, (uint32_t)p.wp, 1u, 1u, 1u};
    return encode_map(m, ptr, 5, dims, strides, box);
  };
  int rc;
  if ((rc = act_map(&p.a[0], d.src0, d.c0))) return rc;
  if (d.c1 && (rc = act_map(&p.a[1], d.src1, d.c1))) return rc;
  if (d.rc0 && (rc = act_map(&p.a[2], d.res0, d.rc0))) return rc;
  if (d.rc1 && (rc = act_map(&p.a[3], d.res1, d.rc1))) return rc;
  {
    const uint64_t ktot = 9ull * (d.c0 + d.c1) + d.rc0 + d.rc1;
    uint64_t dims[2] = {ktot, (uint64_t)d.cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64u, (uint32_t)kHalo2BN};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
  }
  if (d.w_in == 32) return d.cout == 128 ? launch_halo2<32, 128>(p, stream) : launch_halo2<32, 256>(p, stream);
  return d.cout == 128 ? launch_halo2<16, 128>(p, stream) : launch_halo2<16, 256>(p, stream);
}

}  // namespace dmme
