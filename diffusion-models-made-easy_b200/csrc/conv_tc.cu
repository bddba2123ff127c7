// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[M = 128 output pixels][N = BN output channels] (fp32, TMEM)
//       = sum over k-blocks  A[128][64] (bf16, TMA-gathered activation tile)  x  B[BN][64]^T (bf16 weights)
//
// * Activations are NHWC bf16.  A k-block is (filter tap, 64-channel chunk): the A tile of a tap is the
//   output tile shifted by the tap offset, fetched by ONE tiled TMA load whose out-of-bounds rows and
//   columns are zero-filled by the hardware -- that is the conv padding, no im2col buffer exists.
// * The skip-connection concat (models/ddpm.py:310) is two tensor maps: chunks [0, c0/64) come from
//   src0, the rest from src1.  A fused 1x1 residual conv (models/ddpm.py:109,131) is extra k-blocks
//   reading res0|res1 at the centre tap, accumulated into the same TMEM tile.
// * stride 2 (models/ddpm.py:147) uses a parity view of the input: dim0 = (px, c), dim2 = py.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
//   (TMEM -> registers -> +bias +temb +residual -> bf16 -> global).
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct ConvTcParams {
  CUtensorMap a[4];  // src0, src1, res0, res1 (cta_group::2 kernel: boxes of half a pixel tile)
  CUtensorMap b;     // weights [cout][K] bf16
  int chunks0, chunks1, rchunks0, rchunks1;
  int c0, c1;
  int taps, stride;
  int n, ho, wo;
  int bw, bh, bni;
  int tiles_x, tiles_y;
  int m_tiles, n_tiles;  // work units = m_tiles x n_tiles, walked by a persistent grid (n fastest: neighbours share A)
  int cout;
  const float* bias;
  const float* temb;
  int temb_rows, temb_ld;
  const __nv_bfloat16* addend;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  __nv_bfloat16* out3;
  int out_mode;
  long long* stats;  // [n][cout/4][2] fixed-point micro-group sums (GroupNorm statistics of the stored output)
};

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KB
constexpr int kConvEpiWarps = 8;                       // two per TMEM lane quarter, alternating 32-column chunks
constexpr int kConvThreads = (2 + kConvEpiWarps) * 32;

template <int BN>
__host__ __device__ constexpr int stage_bytes() { return kABytes + BN * kBlockK * 2; }

// WS ("weight stationary", 1x1 convs): the whole [BN][K] weight slab of this CTA's n tile is loaded ONCE and stays in
// shared memory; the grid is a multiple of n_tiles so the round-robin walk keeps every CTA on one n tile, and the ring
// carries activation tiles only.  Without it a K = 256 conv re-reads 128 KB of weights per 64 KB activation tile and
// is bound by the L2 -> SM feed (measured 73 us for the 256 -> 768 qkv projection, 353 TFLOP/s).
template <int BN, int STAGES, bool WS>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t slab_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;      // one accumulator stage
  constexpr uint32_t kTmemCols = 2 * kAccCols;           // double-buffered: epilogue of tile i overlaps the MMAs of tile i+1
  constexpr int kStage = WS ? kABytes : stage_bytes<BN>();
  constexpr int kBTile = BN * kBlockK * 2;  // one k-block of weights

  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 1024 B)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* ring = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  // ---- persistent schedule: tile t -> (m tile, n tile) ----
  const int total_tiles = p.m_tiles * p.n_tiles;
#define DMME_TILE_COORDS(t)                                   \
  const int mt = (t) / p.n_tiles;                             \
  const int col0 = ((t) - mt * p.n_tiles) * BN;               \
  const int tx = mt % p.tiles_x;                              \
  const int ty = (mt / p.tiles_x) % p.tiles_y;                \
  const int ng = mt / (p.tiles_x * p.tiles_y);                \
  const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = ng * p.bni;

  const int cchunks = p.chunks0 + p.chunks1;
  const int conv_kb = p.taps * cchunks;
  const int nkb = conv_kb + p.rchunks0 + p.rchunks1;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int st = 0; st < 2; ++st) { mbar_init(&acc_full[st], 1); mbar_init(&acc_empty[st], kConvEpiWarps * 32); }
    mbar_init(&slab_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a[0]);
    if (p.chunks1) tma_prefetch_desc(&p.a[1]);
    if (p.rchunks0) tma_prefetch_desc(&p.a[2]);
    if (p.rchunks1) tma_prefetch_desc(&p.a[3]);
    tma_prefetch_desc(&p.b);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  uint8_t* slab = ring + STAGES * kStage;  // WS only
  pdl_trigger();  // after the TMEM allocation: a dependent kernel's CTA may now take this SM's free resources

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int it = 0;
      pdl_wait();  // (the packed weights may come from a pack kernel launched just before this one, too)
      if (WS) {
        mbar_expect_tx(&slab_full, static_cast<uint32_t>(nkb) * kBTile);
        const int wcol0 = (blockIdx.x % p.n_tiles) * BN;
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(slab + kb * kBTile, &p.b, &slab_full, kb * kBlockK, wcol0);
      }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      DMME_TILE_COORDS(t)
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], kStage);
        uint8_t* sa = ring + s * kStage;
        int which, cc, cx = x0, cy = y0, cp = 0;
        if (kb < conv_kb) {
          const int tap = kb / cchunks;
          int ch = kb - tap * cchunks;
          which = ch < p.chunks0 ? 0 : 1;
          if (which) ch -= p.chunks0;
          cc = ch * kBlockK;
          if (p.taps == 9) {
            const int r = tap / 3, q = tap - r * 3;
            if (p.stride == 1) {
              cx += q - 1;
              cy += r - 1;
            } else {
              // input pixel 2*o + r - 1: r=0 -> (o-1, parity 1), r=1 -> (o, 0), r=2 -> (o, 1)
              const int csrc = which ? p.c1 : p.c0;
              cc += (q != 1) ? csrc : 0;
              cx += (q == 0) ? -1 : 0;
              cp = (r != 1) ? 1 : 0;
              cy += (r == 0) ? -1 : 0;
            }
          }
        } else {
          int ch = kb - conv_kb;
          which = ch < p.rchunks0 ? 2 : 3;
          if (which == 3) ch -= p.rchunks0;
          cc = ch * kBlockK;
        }
        tma_load_5d(sa, &p.a[which], &full_bar[s], cc, cx, cp, cy, n0);
        if (!WS) tma_load_2d(sa + kABytes, &p.b, &full_bar[s], kb * kBlockK, col0);
      }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      int it = 0, t_it = 0;
      if (WS) mbar_wait(&slab_full, 0);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++t_it) {
      const int stage = t_it & 1;
      mbar_wait(&acc_empty[stage], ((t_it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t dtm = tmem_base + stage * kAccCols;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + s * kStage);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(WS ? smem_u32(slab + kb * kBTile) : sa + kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // +32 bytes (16 bf16) along K inside the 128-byte swizzle row: start-address field += 2
          umma_bf16(dtm, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
      }
      umma_commit(&acc_full[stage]);  // accumulator complete
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which of the quarter's two warps: even / odd 32-column chunks
    int t_it = 0;
    pdl_wait();
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++t_it) {
    DMME_TILE_COORDS(t)
    const int stage = t_it & 1;
    const int row = q * 32 + lane;
    const int wx = row % p.bw;
    const int hy = (row / p.bw) % p.bh;
    const int ni = row / (p.bw * p.bh);
    const int n = n0 + ni, y = y0 + hy, x = x0 + wx;
    const bool valid = (n < p.n) && (y < p.ho) && (x < p.wo);
    const long long pix = (static_cast<long long>(n) * p.ho + y) * p.wo + x;
    const float* trow = p.temb ? p.temb + static_cast<long long>(p.temb_rows == 1 ? 0 : n) * p.temb_ld : nullptr;

    // the addend does not depend on the accumulator: its first chunk is requested before the wait, every further
    // chunk while the previous one is converted and stored (the load latency used to serialise with each chunk)
    const uint4* arow = (p.addend && valid) ? reinterpret_cast<const uint4*>(p.addend + pix * p.cout + col0) : nullptr;
    uint4 apre[4];
    if (arow) {
#pragma unroll
      for (int j = 0; j < 4; ++j) apre[j] = __ldg(arow + half * 4 + j);
    }

    mbar_wait(&acc_full[stage], (t_it >> 1) & 1);
    tc_fence_after();

    int which = 0, ccol0 = col0, cmod = p.cout;
    if (p.out_mode == DMME_OUT_QKV) {
      cmod = p.cout / 3;
      which = col0 / cmod;
      ccol0 = col0 - which * cmod;
    }
    const int L = p.ho * p.wo;

    // lanes of one warp that belong to the same image (GroupNorm statistics are per image)
    const int seg = L < 32 ? L : 32;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);

#pragma unroll 1
    for (int c = half * 32; c < BN; c += 64) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(stage * kAccCols + c), v);
      uint4 acur[4];
      if (arow) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acur[j] = apre[j];
        if (c + 64 < BN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) apre[j] = __ldg(arow + (c + 64) / 8 + j);
        }
      }
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const int col = col0 + c;
      if (valid) {
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
          }
        }
        if (trow) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + col + j));
            f[j] += t4.x; f[j + 1] += t4.y; f[j + 2] += t4.z; f[j + 3] += t4.w;
          }
        }
        if (arow) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 a4 = acur[j];
            float lo, hi;
            unpack_bf16x2(a4.x, lo, hi); f[8 * j + 0] += lo; f[8 * j + 1] += hi;
            unpack_bf16x2(a4.y, lo, hi); f[8 * j + 2] += lo; f[8 * j + 3] += hi;
            unpack_bf16x2(a4.z, lo, hi); f[8 * j + 4] += lo; f[8 * j + 5] += hi;
            unpack_bf16x2(a4.w, lo, hi); f[8 * j + 6] += lo; f[8 * j + 7] += hi;
          }
        }
        if (p.out_mode == DMME_OUT_QKV && which == 2) {
          // V^T: [n][C][L], one pixel per lane -> 64-byte coalesced runs per channel
          __nv_bfloat16* vt = p.out3 + (static_cast<long long>(n) * cmod + ccol0 + c) * L + (y * p.wo + x);
#pragma unroll
          for (int j = 0; j < 32; ++j) vt[static_cast<long long>(j) * L] = __float2bfloat16_rn(f[j]);
        } else {
          __nv_bfloat16* dst;
          if (p.out_mode == DMME_OUT_QKV) {
            dst = (which == 0 ? p.out : p.out2) + pix * cmod + ccol0 + c;
          } else {
            dst = p.out + pix * p.cout + col;
          }
          uint4* dp = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
            o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
            o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            dp[j] = o;
            // keep the rounded values: the statistics below describe the tensor as stored
            unpack_bf16x2(o.x, f[8 * j + 0], f[8 * j + 1]);
            unpack_bf16x2(o.y, f[8 * j + 2], f[8 * j + 3]);
            unpack_bf16x2(o.z, f[8 * j + 4], f[8 * j + 5]);
            unpack_bf16x2(o.w, f[8 * j + 6], f[8 * j + 7]);
          }
        }
      }
      if (p.stats) {  // warp-uniform
        float s1[8], s2[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float a = f[4 * g], b = f[4 * g + 1], cc = f[4 * g + 2], dd = f[4 * g + 3];
          s1[g] = valid ? (a + b) + (cc + dd) : 0.f;
          s2[g] = valid ? (a * a + b * b) + (cc * cc + dd * dd) : 0.f;
        }
        for (int off = seg >> 1; off > 0; off >>= 1) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            s1[g] += __shfl_xor_sync(0xffffffffu, s1[g], off);
            s2[g] += __shfl_xor_sync(0xffffffffu, s2[g], off);
          }
        }
        if ((lane & (seg - 1)) == 0 && n < p.n) {
          unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                   (static_cast<long long>(n) * (p.cout >> 2) + (col >> 2)) * 2;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            atomicAdd(st + 2 * g, static_cast<unsigned long long>(__float2ll_rn(s1[g] * kFix)));
            atomicAdd(st + 2 * g + 1, static_cast<unsigned long long>(__float2ll_rn(s2[g] * kFix)));
          }
        }
      }
    }
    tc_fence_before();
    mbar_arrive(&acc_empty[stage]);
    }
  }
#undef DMME_TILE_COORDS

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ================================================================================================================
// Transposed variant:  D^T[M = 128 output channels][N = NP output pixels] = W[128][K] x X^T
//
// Same TMA-gathered operand tiles and tap schedule as conv_tc_kernel, with the operand roles swapped: the weight tile
// is the MMA's M side, the pixel tile (NP = 128 or 256 pixels: a contiguous run of the flattened [n][y][x] space) the
// N side, so the accumulator has TMEM lane = output channel, column = pixel.  The epilogue is then the halo kernel's:
// thread = channel, a warp stores 32 channels x 2 B = 64 contiguous bytes per pixel (conv_tc_kernel's thread = pixel
// layout writes sixteen bytes into each of 32 different lines per instruction: 2.1 M partial-sector L2 writes per
// 33 MB output, measured 31 us for a 256 -> 256 1x1 conv whose MMAs take 6 us), GroupNorm statistics accumulate in
// the thread over the pixels of an image (no shuffle butterflies), the addend is read with the same coalescing.
// WS: 1x1 convs keep their [128][K] weight slab resident (see conv_tc_kernel).
// ================================================================================================================
constexpr int kTctMaxStages = 8;
constexpr int kTctProducers = 1;  // more producer lanes do not help: the k-block rate is the SS-mode MMA rate

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(lo)) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);
}

struct ConvTctExtra {
  int np;           // pixels per tile (MMA N): 128 or 256
  int stages;       // ring depth
  int stage_bytes;  // np * 128 (+ 16 KB weight tile unless WS)
  long long total_pix;
  int log2_l;       // log2(ho * wo)
  int subpix;       // 1: nearest x2 + 3x3 conv as four 2x2 phase convs on the low-resolution tensor (see below)
  int log2_w;       // log2(wo) (subpix output addressing)
  long long* trace; // debugging: per-role clock64 timestamps of CTA 0 (see dmme_debug_set_conv_trace), or null
  // split-K (SPLIT kernels): a work unit is (pixel tile, channel tile, K slice s of `split`); it accumulates k-blocks
  // [s * nkb / split, (s + 1) * nkb / split) and stores its fp32 tile to partial[s][pixel][cout] (no epilogue terms: the
  // finishing pass conv_splitk.cu adds them after summing the slices)
  int split;
  float* partial;
  long long split_stride;  // elements between two K slices = total_pix * cout
  // NORM kernels (8x8 maps: an image = two 32-pixel chunks of one warp, a GroupNorm group = cpg lanes of it): the
  // GroupNorm(+SiLU) of the stored tensor for up to two consumers, written by the epilogue (dmme_conv_desc.out_norm)
  struct Norm {
    __nv_bfloat16* out;
    const float* gamma; const float* beta;
    const float* scale; const float* shift;
    int ss_rows, ss_ld, cpg, silu;
    float eps;
  } no[2];
};

// trace slots of CTA 0: [role][event index]; role 0 = producer (issue time per k-block), 1 = MMA (operands landed per
// k-block), 2 = epilogue warp 2 (accumulator ready / tile stored per tile); 512 events per role
__device__ __forceinline__ void trace_ev(long long* trace, int role, int idx) {
  if (trace && blockIdx.x == 0 && idx < 512) trace[role * 512 + idx] = clock64();
}

// Sub-pixel mode (x.subpix): UpSample = nearest x2 then conv3x3 (models/ddpm.py:150-173).  On the x2 grid, output
// (2i + a, 2j + b) only ever sees the 2x2 low-resolution neighbourhood {i - 1 + a, i + a} x {j - 1 + b, j + b}, with the
// 3x3 taps that coincide summed: rows {0 | 1+2} for a = 0, {0+1 | 2} for a = 1, likewise for columns.  So the layer is four
// 2x2 convolutions ("phases" (a, b)) of the LOW-resolution tensor, each writing one parity class of the output: 16
// instead of 36 tap-MACs per low-resolution pixel (2.25x fewer FLOPs) and no materialised x2 tensor.  The packed weight
// is [4 phases][cout][4 taps x cin] (summed in fp32, rounded to bf16 once); a work unit is (phase, pixel tile, channel
// tile); TMA zero fill of rows / columns -1 and H, W is exactly the zero padding of the x2 grid.
// CMOD: channel stride of the output rows when known at compile time (128 / 256: immediate store offsets), 0 = runtime
// PAIR: cta_group::2 -- the two CTAs of a cluster compute 256 output channels x NP pixels with one MMA stream (see
//       ptx_sm100.cuh): CTA r of the pair owns channels [128 r, 128 r + 128) (its TMEM lanes, its weight tiles) and
//       stages pixels [r NP/2, (r + 1) NP/2) of the pixel tile (the tensor maps then have half-tile boxes), so its L2 -> SM feed
//       per k-block is 16 + 16 KB instead of 32 + 16 KB for the same 128 x NP outputs.  The 3x3 convs of the 16x16 /
//       8x8 levels are bound by that feed (~45 B/clk/SM measured), not by the tensor pipe.
// Epilogue warps: 8 (two per TMEM lane quarter).  16 for the weight-stationary 1x1 convs (whose accumulator drain is the
// kernel) was tried and is SLOWER: 576 threads leave 112 registers each, the epilogue spills (1.3 KB) and the 256 -> 256
// projection went from 27.6 to 35.7 us, the qkv projection from 64 to 102 us.
template <bool WS>
__host__ __device__ constexpr int tct_epi_warps() { return 8; }

// CLUSTER (with SPLIT, 4x4 maps): the `split` K slices of a tile are the CTAs of ONE thread-block cluster (grid = units, one
//       unit per CTA, cluster rank = K slice).  Instead of storing fp32 partial tiles for a finishing launch, the slices are
//       reduce-scattered through distributed shared memory: CTA r owns pixels [r NP/split, (r + 1) NP/split) of the tile
//       (whole 16-pixel images), every other CTA writes its partial values of those pixels into r's (dead) TMA ring, and r
//       finishes them -- slice sum in the finishing pass's order, bias / temb / addend, bf16 store, statistics, the
//       consumers' GroupNorm(+SiLU) -- exactly as conv_splitk.cu does (same bits).  One launch instead of two and no partial
//       tiles in global memory.
//       CLPX: pixels per image of the maps it handles (16: 4x4, 64: 8x8 -- there only raw output and slice order are
//       bit-equal to the two-launch path, the statistics are summed by one thread per (image, channel) instead of 16 warps).
template <bool WS, int CMOD, bool PAIR, bool SPLIT = false, bool NORM = false, bool CLUSTER = false, int CLPX = 16>
__global__ void __launch_bounds__((2 + tct_epi_warps<WS>()) * 32, 1) conv_tct_kernel(const __grid_constant__ ConvTcParams p,
                                                                                  const ConvTctExtra x) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTctMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kTctMaxStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t slab_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kAccCols = 256;
  constexpr int kWTile = 128 * kBlockK * 2;  // [128 cout][64] bf16
  const int NP = x.np, STAGES = x.stages, kStage = x.stage_bytes;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* ring = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slab = ring + STAGES * kStage;  // WS only

  const int nphase = x.subpix ? 4 : 1;
  const int nsplit = SPLIT ? x.split : 1;
  const int total_tiles = p.m_tiles * p.n_tiles * nphase * nsplit;  // PAIR: n_tiles counts 256-channel tiles
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int cta0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);  // first work unit
  const int cta_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int kTileCh = PAIR ? 256 : 128;
  // unit index: channel tile fastest, then phase (the four phases of a pixel tile read the same pixels), then pixel tile
  // SPLIT: the K slice is the fastest unit index (the slices of one tile run concurrently on neighbouring CTAs)
#define DMME_TCT_COORDS(t)                                    \
  const int tu_ = SPLIT ? (t) / nsplit : (t);                 \
  const int ks_ = SPLIT ? (t) - tu_ * nsplit : 0;             \
  const int kb_lo = SPLIT ? (ks_ * nkb) / nsplit : 0;         \
  const int kb_hi = SPLIT ? ((ks_ + 1) * nkb) / nsplit : nkb; \
  const int nt_ = tu_ % p.n_tiles;                            \
  const int sub = (tu_ / p.n_tiles) % nphase; /* sub-pixel phase */ \
  const int mt = tu_ / (p.n_tiles * nphase);                  \
  const int col0 = nt_ * kTileCh + static_cast<int>(rank) * 128; \
  const int tx = mt % p.tiles_x;                              \
  const int ty = (mt / p.tiles_x) % p.tiles_y;                \
  const int ng = mt / (p.tiles_x * p.tiles_y);                \
  const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = ng * p.bni;

  const int cchunks = p.chunks0 + p.chunks1;
  const int conv_kb = p.taps * cchunks;
  const int nkb = conv_kb + p.rchunks0 + p.rchunks1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    // PAIR: the leader's acc_empty collects the epilogue threads of both CTAs
    for (int st = 0; st < 2; ++st) { mbar_init(&acc_full[st], 1); mbar_init(&acc_empty[st], (PAIR ? 2 : 1) * tct_epi_warps<WS>() * 32); }
    mbar_init(&slab_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  constexpr int kMapBase = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a[kMapBase + 0]);
    if (p.chunks1) tma_prefetch_desc(&p.a[kMapBase + 1]);
    if (p.rchunks0) tma_prefetch_desc(&p.a[kMapBase + 2]);
    if (p.rchunks1) tma_prefetch_desc(&p.a[kMapBase + 3]);
    tma_prefetch_desc(&p.b);
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2sm(&tmem_slot, 2 * kAccCols);
    else tmem_alloc(&tmem_slot, 2 * kAccCols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before either signals the other's
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();  // after the TMEM allocation (see common.cuh)

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // kTctProducers lanes take the k-blocks round-robin (1: a single lane keeps up -- k-blocks land every ~700 clocks
    // whether one or four lanes issue them and whether they carry 32 or 48 KB: the pace is the MMA's)
    if (lane < kTctProducers) {
      int it = 0;
      pdl_wait();
      trace_ev(x.trace, 0, 0);
      if (WS && lane == 0) {
        mbar_expect_tx(&slab_full, static_cast<uint32_t>(nkb) * kWTile);
        const int wcol0 = (blockIdx.x % p.n_tiles) * 128;
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(slab + kb * kWTile, &p.b, &slab_full, kb * kBlockK, wcol0);
      }
      for (int t = cta0; t < total_tiles; t += cta_step) {
        DMME_TCT_COORDS(t)
        // PAIR: this CTA stages the second half of the pixel tile when it is the odd one: the tile is split along its
        // slowest dimension (images when it spans several, rows otherwise)
        const int half_n = PAIR ? (p.bni >= 2 ? static_cast<int>(rank) * (p.bni >> 1) : 0) : 0;
        const int half_y = PAIR ? (p.bni >= 2 ? 0 : static_cast<int>(rank) * (p.bh >> 1)) : 0;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          if (it % kTctProducers != lane) continue;
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          // PAIR: only the leader arms its barrier, for the bytes of both CTAs
          if (!PAIR) mbar_expect_tx(&full_bar[s], kStage);
          else if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * kStage);
          uint8_t* sx = ring + s * kStage;
          int which, cc, cx = x0, cy = y0 + half_y, cp = 0;
          if (kb < conv_kb) {
            const int tap = kb / cchunks;
            int ch = kb - tap * cchunks;
            which = ch < p.chunks0 ? 0 : 1;
            if (which) ch -= p.chunks0;
            cc = ch * kBlockK;
            if (p.taps == 9) {
              const int r = tap / 3, q = tap - r * 3;
              if (p.stride == 1) {
                cx += q - 1;
                cy += r - 1;
              } else {
                const int csrc = which ? p.c1 : p.c0;
                cc += (q != 1) ? csrc : 0;
                cx += (q == 0) ? -1 : 0;
                cp = (r != 1) ? 1 : 0;
                cy += (r == 0) ? -1 : 0;
              }
            } else if (p.taps == 4) {  // sub-pixel phase (a, b) = (sub >> 1, sub & 1), tap (u, v) = (tap >> 1, tap & 1)
              cy += (sub >> 1) - 1 + (tap >> 1);
              cx += (sub & 1) - 1 + (tap & 1);
            }
          } else {
            int ch = kb - conv_kb;
            which = ch < p.rchunks0 ? 2 : 3;
            if (which == 3) ch -= p.rchunks0;
            cc = ch * kBlockK;
          }
          const int wrow = sub * p.cout + col0;  // the phase's block of the packed weight matrix
          if (lane == 0) trace_ev(x.trace, 0, it + 1);
          if (PAIR) {
            const uint32_t lead_bar = mapa_u32(&full_bar[s], 0);
            tma_load_5d_2sm(sx, &p.a[kMapBase + which], lead_bar, cc, cx, cp, cy, n0 + half_n);
            tma_load_2d_2sm(sx + (NP >> 1) * 128, &p.b, lead_bar, kb * kBlockK, wrow);
          } else {
            tma_load_5d(sx, &p.a[which], &full_bar[s], cc, cx, cp, cy, n0);
            if (!WS) tma_load_2d(sx + NP * 128, &p.b, &full_bar[s], kb * kBlockK, wrow);
          }
        }
      }
    }
    if constexpr (CLUSTER) { cluster_sync_all(); cluster_sync_all(); }  // the epilogue warps' exchange barriers (below)
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, NP);  // M = output channels (of the pair), N = pixels
      int it = 0, t_it = 0;
      if (WS) mbar_wait(&slab_full, 0);
      for (int t = cta0; t < total_tiles; t += cta_step, ++t_it) {
        const int stage = t_it & 1;
        const int ksm = SPLIT ? t % nsplit : 0;
        const int kb_lo = SPLIT ? (ksm * nkb) / nsplit : 0, kb_hi = SPLIT ? ((ksm + 1) * nkb) / nsplit : nkb;
        mbar_wait(&acc_empty[stage], ((t_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dtm = tmem_base + stage * kAccCols;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          trace_ev(x.trace, 1, it);
          const uint32_t sx = smem_u32(ring + s * kStage);
          const uint64_t xdesc = umma_desc_sw128(sx);
          const uint64_t wdesc = umma_desc_sw128(WS ? smem_u32(slab + kb * kWTile) : sx + (PAIR ? NP >> 1 : NP) * 128);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (PAIR) umma_bf16_2sm(dtm, wdesc + 2 * k, xdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(dtm, wdesc + 2 * k, xdesc + 2 * k, idesc, (kb != kb_lo || k != 0) ? 1u : 0u);
          }
          if (PAIR) umma_commit_2sm(&empty_bar[s], 3);  // the slot is free in both CTAs
          else umma_commit(&empty_bar[s]);
        }
        if (PAIR) umma_commit_2sm(&acc_full[stage], 3);
        else umma_commit(&acc_full[stage]);
      }
    }
    if constexpr (CLUSTER) { cluster_sync_all(); cluster_sync_all(); }
  } else {
    // =========================== epilogue: thread = output channel ===========================
    // Eight warps share one accumulator, so the epilogue is bound by instruction issue: the common case (a chunk of
    // 32 valid pixels of one image, channel stride known at compile time) costs 6 instructions per output -- add,
    // convert, store with an immediate offset, widen, two statistics updates -- against 24 for the fully general form
    // (measured 34 us -> see profiles/ for a 256 -> 256 1x1 conv at batch 256).
    const int q = warp & 3;            // TMEM lane quarter = 32-channel block of the tile
    const int half = (warp - 2) >> 2;  // which part of the tile's pixels (2 or 4 parts)
    const int ppw = NP / (tct_epi_warps<WS>() / 4);  // pixels per warp and tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    const int L = p.ho * p.wo;
    const int total_pix = static_cast<int>(x.total_pix);
    const bool temb_per_image = p.temb && p.temb_rows != 1;
    const int cmod = CMOD ? CMOD : (p.out_mode == DMME_OUT_QKV ? p.cout / 3 : p.cout);
    const int nchunks = ppw >> 5;  // 32-pixel chunks per warp and tile
    int t_it = 0;
    pdl_wait();
    for (int t = cta0; t < total_tiles; t += cta_step, ++t_it) {
      const int tu = SPLIT ? t / nsplit : t;
      const int ks = SPLIT ? t - tu * nsplit : 0;  // K slice (SPLIT)
      const int mt = tu / (p.n_tiles * nphase);
      const int sub = (tu / p.n_tiles) % nphase;  // sub-pixel phase
      // first channel of this warp's block (warp-uniform)
      const int cb = (tu % p.n_tiles) * kTileCh + static_cast<int>(rank) * 128 + q * 32;
      const int ch = cb + lane;                            // this thread's output channel
      const int stage = t_it & 1;
      float bias_c = p.bias ? __ldg(p.bias + ch) : 0.f;
      if (p.temb && !temb_per_image) bias_c += __ldg(p.temb + ch);  // broadcast row (sampling): no per-image reload
      // q / k / v selector, 0 unless QKV.  Derived from the warp's channel block, NOT from the lane's channel: a
      // lane-dependent value here makes every branch on it divergent as far as ptxas can tell, and the stores of the
      // fast path then re-materialise their uniform address descriptor with two R2UR each
      const int which = cb / cmod;
      const int chm = cb - which * cmod + lane;  // channel inside q / k / v
      __nv_bfloat16* obase = which == 0 ? p.out : (which == 1 ? p.out2 : p.out3);
      const int pix_begin = mt * NP + half * ppw;
      float s1 = 0.f, s2 = 0.f, bt = bias_c;
      int cur_n = -1;
      // NORM: the first chunk of the image in flight (packed stored values) and the consumers' affine terms of this channel
      uint32_t kept[16];
      float ng[2] = {1.f, 1.f}, nb[2] = {0.f, 0.f};
      if constexpr (NORM) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (x.no[k].out == nullptr) continue;
          if (x.no[k].gamma) ng[k] = __ldg(x.no[k].gamma + ch);
          if (x.no[k].beta) nb[k] = __ldg(x.no[k].beta + ch);
        }
      }

      auto flush_stats = [&]() {  // warp-uniform call: per-image sums of this lane's channel -> micro-group atomics
        if (p.stats && cur_n >= 0) {
          float a1 = s1, a2 = s2;
          a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
          if ((lane & 3) == 0) {
            unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                     (static_cast<long long>(cur_n) * (p.cout >> 2) + (ch >> 2)) * 2;
            atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(a1 * kFix)));
            atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(a2 * kFix)));
          }
        }
        s1 = 0.f; s2 = 0.f;
      };
      auto enter_image = [&](int n) {  // warp-uniform
        flush_stats();
        cur_n = n;
        bt = bias_c;
        if (temb_per_image) bt += __ldg(p.temb + static_cast<long long>(n) * p.temb_ld + ch);
      };
      // addend values of one chunk (thread = channel: 64 contiguous bytes per pixel across the warp)
      auto load_addend = [&](float (&av)[32], int pix0) {
        const __nv_bfloat16* __restrict__ ap = p.addend + static_cast<long long>(pix0) * p.cout + ch;
        if (pix0 + 32 <= total_pix) {
#pragma unroll
          for (int i = 0; i < 32; ++i) av[i] = __bfloat162float(__ldg(ap + i * cmod));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) av[i] = (pix0 + i < total_pix) ? __bfloat162float(__ldg(ap + i * cmod)) : 0.f;
        }
      };

      // the addend does not depend on the accumulator: chunk 0 is requested before the accumulator is awaited, chunk
      // c + 1 before chunk c is converted and stored
      float av[32];
      const bool has_add = p.addend != nullptr && pix_begin < total_pix;
      if (has_add) load_addend(av, pix_begin);

      mbar_wait(&acc_full[stage], (t_it >> 1) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) trace_ev(x.trace, 2, 2 * t_it);

      if constexpr (CLUSTER) {
        const int S = nsplit;
        const int crank = static_cast<int>(cluster_ctarank());  // == ks: the K slice this CTA accumulated
        const int own = NP / S;                                  // pixels of the tile this CTA finishes (whole images)
        float* recv = reinterpret_cast<float*>(ring);            // [(S - 1) senders][own pixels][128 channels] fp32
        const int cl = q * 32 + lane;                            // channel inside the 128-channel tile
        const uint32_t tacc = tmem_base + lane_off + static_cast<uint32_t>(stage * kAccCols);
        cluster_sync_all();  // every CTA's accumulator is complete: every TMA ring of the cluster is dead
        // ---- send: this warp's lane quarter x its half of the columns, image by image (16 columns) to the image's owner ----
#pragma unroll 1
        for (int ci = 0; ci < 2 * nchunks; ++ci) {
          const int col0 = half * ppw + ci * 16;
          const int r = col0 / own;  // own is a multiple of 16: an image has one owner
          if (r == crank) continue;
          uint32_t v[16];
          tmem_ld16(tacc + static_cast<uint32_t>(col0), v);
          tmem_ld_wait();
          const int slot = crank < r ? crank : crank - 1;
          const uint32_t dst = mapa_u32(recv, static_cast<uint32_t>(r)) +
                               static_cast<uint32_t>(((slot * own + (col0 - r * own)) * 128 + cl) * 4);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(dst + static_cast<uint32_t>(i * 512)), "r"(v[i]) : "memory");
        }
        cluster_sync_all();  // release / acquire: the other slices' values of this CTA's pixels have landed
        // ---- finish: thread = channel, half h takes images [h, h + 1) * own / 32 of this CTA's pixels; the arithmetic and
        // its order are splitk_finish_small_kernel's (conv_splitk.cu) ----
        // pixels per thread: whole images; a CTA whose pixels do not split into two sets of whole images leaves them to the
        // first half of the warps
        const bool both = (own >> 1) % CLPX == 0;
        const int ppt = both ? own >> 1 : (half == 0 ? own : 0);
        const float kFixc = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
#pragma unroll 1
        for (int ib = 0; ib < ppt; ib += CLPX) {
          const int lcol0 = (both ? half * ppt : 0) + ib;   // first pixel of the image inside this CTA's range
          const int pix = mt * NP + crank * own + lcol0;    // global pixel index
          if (pix >= total_pix) break;
          const int n = pix / CLPX;
          float add = p.bias ? __ldg(p.bias + ch) : 0.f;
          if (p.temb) add += __ldg(p.temb + static_cast<long long>(p.temb_rows == 1 ? 0 : n) * p.temb_ld + ch);
          float rf[CLPX];
          float t1s = 0.f, t2s = 0.f;
          __nv_bfloat16* orow = p.out + static_cast<long long>(pix) * p.cout + ch;
#pragma unroll
          for (int sb = 0; sb < CLPX / 16; ++sb) {
            const int lcol = lcol0 + 16 * sb;
            float vsum[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) vsum[i] = add;
#pragma unroll 1
            for (int j0 = 0; j0 < S; j0 += 4) {  // slices in fours, as the finishing pass sums them
              float a4[4][16];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j0 + jj;
                if (j >= S) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) a4[jj][i] = 0.f;
                } else if (j == crank) {
                  uint32_t t16[16];
                  tmem_ld16(tacc + static_cast<uint32_t>(crank * own + lcol), t16);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) a4[jj][i] = __uint_as_float(t16[i]);
                } else {
                  const float* rp = recv + ((j < crank ? j : j - 1) * own + lcol) * 128 + cl;
#pragma unroll
                  for (int i = 0; i < 16; ++i) a4[jj][i] = rp[i * 128];
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) vsum[i] += (a4[0][i] + a4[1][i]) + (a4[2][i] + a4[3][i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int pi = 16 * sb + i;
              const float ad = p.addend ? __bfloat162float(p.addend[static_cast<long long>(pix + pi) * p.cout + ch]) : 0.f;
              const __nv_bfloat16 rb = __float2bfloat16_rn(vsum[i] + ad);
              orow[static_cast<long long>(pi) * p.cout] = rb;
              rf[pi] = __bfloat162float(rb);
              t1s += rf[pi];
              t2s = fmaf(rf[pi], rf[pi], t2s);
            }
          }
          if (p.stats) {
            float m1 = t1s, m2 = t2s;
            m1 += __shfl_xor_sync(0xffffffffu, m1, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
            m1 += __shfl_xor_sync(0xffffffffu, m1, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
            if ((lane & 3) == 0) {
              unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                       (static_cast<long long>(n) * (p.cout >> 2) + (ch >> 2)) * 2;
              atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(m1 * kFixc)));
              atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(m2 * kFixc)));
            }
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const ConvTctExtra::Norm& qn = x.no[k];
            if (qn.out == nullptr) continue;  // uniform
            const int g0 = (lane / qn.cpg) * qn.cpg;
            float g1 = 0.f, g2 = 0.f;
            for (int j = 0; j < qn.cpg; ++j) {  // ascending channel order, as the finishing pass
              g1 += __shfl_sync(0xffffffffu, t1s, g0 + j);
              g2 += __shfl_sync(0xffffffffu, t2s, g0 + j);
            }
            const float inv_cnt = 1.0f / (static_cast<float>(CLPX) * qn.cpg);
            const float mean = g1 * inv_cnt;
            const float var = fmaxf(g2 * inv_cnt - mean * mean, 0.f);
            const float rs = rsqrtf(var + qn.eps);
            const float ga = qn.gamma ? __ldg(qn.gamma + ch) : 1.f, be = qn.beta ? __ldg(qn.beta + ch) : 0.f;
            float aa = rs * ga, bb = be - mean * rs * ga;
            if (qn.scale) {
              const long long rr = static_cast<long long>(qn.ss_rows == 1 ? 0 : n) * qn.ss_ld;
              const float sc = 1.f + __ldg(qn.scale + rr + ch), sh = __ldg(qn.shift + rr + ch);
              aa *= sc;
              bb = bb * sc + sh;
            }
            __nv_bfloat16* yrow = qn.out + static_cast<long long>(pix) * p.cout + ch;
#pragma unroll
            for (int i = 0; i < CLPX; ++i) {
              float y = fmaf(rf[i], aa, bb);
              if (qn.silu) y = silu_f(y);
              yrow[static_cast<long long>(i) * p.cout] = __float2bfloat16_rn(y);
            }
          }
        }
      } else if constexpr (SPLIT) {
        // fp32 partial tile of this K slice: thread = channel, a warp stores 128 contiguous bytes per pixel
        float* __restrict__ pp = x.partial + ks * x.split_stride + ch;
#pragma unroll 1
        for (int ci = 0; ci < nchunks; ++ci) {
          const int pix0 = pix_begin + ci * 32;
          if (pix0 >= total_pix) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_off + static_cast<uint32_t>(stage * kAccCols + half * ppw + ci * 32), v);
          tmem_ld_wait();
          float* po = pp + static_cast<long long>(pix0) * p.cout;
          if (pix0 + 32 <= total_pix) {
#pragma unroll
            for (int i = 0; i < 32; ++i) po[static_cast<long long>(i) * p.cout] = __uint_as_float(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (pix0 + i < total_pix) po[static_cast<long long>(i) * p.cout] = __uint_as_float(v[i]);
          }
        }
      }
#pragma unroll 1
      for (int ci = 0; ci < (SPLIT ? 0 : nchunks); ++ci) {
        const int pix0 = pix_begin + ci * 32;  // first pixel of this 32-pixel chunk (warp-uniform)
        if (pix0 >= total_pix) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + static_cast<uint32_t>(stage * kAccCols + half * ppw + ci * 32), v);
        float f[32];
        if (has_add) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = av[i];
          if (ci + 1 < nchunks && pix0 + 32 < total_pix) load_addend(av, pix0 + 32);
        }
        tmem_ld_wait();
        const bool full = pix0 + 32 <= total_pix;
        __nv_bfloat16 r[32];
        if (full && L >= 32) {
          // ---- fast path: 32 valid pixels of one image ----
          const int n = pix0 >> x.log2_l;
          if (n != cur_n) enter_image(n);
          if (has_add) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += __uint_as_float(v[i]) + bt;
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + bt;
          }
          if (x.subpix) {
            // low-resolution pixel (n, i, j) -> output pixel (n, 2 i + a, 2 j + b) of the x2 tensor
            const int rem = pix0 - (n << x.log2_l);
            const int i0 = rem >> x.log2_w, j0 = rem & (p.wo - 1);
            const long long obase_px = (static_cast<long long>(n) * (2 * p.ho) + 2 * i0 + (sub >> 1)) * (2 * p.wo) + 2 * j0 + (sub & 1);
            __nv_bfloat16* op = obase + obase_px * cmod + chm;
            const int row_step = 4 * p.wo * cmod;  // two output rows per low-resolution row
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const uint32_t u = pack_bf16x2(f[i], f[i + 1]);
              // the chunk starts at column j0 of row i0 and may wrap over several (power-of-two wide) rows
              const int q0 = j0 + i, q1 = j0 + i + 1;
              op[(q0 >> x.log2_w) * row_step + ((q0 & (p.wo - 1)) - j0) * 2 * cmod] =
                  __ushort_as_bfloat16(static_cast<unsigned short>(u & 0xffffu));
              op[(q1 >> x.log2_w) * row_step + ((q1 & (p.wo - 1)) - j0) * 2 * cmod] =
                  __ushort_as_bfloat16(static_cast<unsigned short>(u >> 16));
              float lo, hi;
              unpack_bf16x2(u, lo, hi);
              s1 += lo + hi;
              s2 = fmaf(lo, lo, s2);
              s2 = fmaf(hi, hi, s2);
            }
            continue;
          }
          if (which != 2) {
            __nv_bfloat16* op = obase + static_cast<long long>(pix0) * cmod + chm;
            // four independent partial sums per statistic: one serial chain of 64 dependent FADD / FFMA per chunk
            // left the two warps of a scheduler nothing to issue
            float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t cur[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              // packed conversion (F2FP, FMA pipe) instead of two F2F (quarter-rate conversion pipe)
              const uint32_t u = pack_bf16x2(f[i], f[i + 1]);
#ifndef DMME_EXP_NOSTORE
              op[i * cmod] = __ushort_as_bfloat16(static_cast<unsigned short>(u & 0xffffu));
              op[(i + 1) * cmod] = __ushort_as_bfloat16(static_cast<unsigned short>(u >> 16));
#else
              if (u == 0x12345678u) op[i * cmod] = __ushort_as_bfloat16(static_cast<unsigned short>(u & 0xffffu));
#endif
              if constexpr (NORM) cur[i >> 1] = u;
              float lo, hi;
              unpack_bf16x2(u, lo, hi);
              p1[(i >> 1) & 3] += lo + hi;
              p2[(i >> 1) & 3] = fmaf(lo, lo, p2[(i >> 1) & 3]);
              p2[((i >> 1) + 2) & 3] = fmaf(hi, hi, p2[((i >> 1) + 2) & 3]);
            }
            s1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
            s2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
            if constexpr (NORM) {
              // 8x8 maps: chunks 2 j, 2 j + 1 of this warp are one image.  After the second, s1 / s2 are the image's sums of
              // this channel's STORED values; a GroupNorm group is cpg neighbouring lanes, so the consumers' norm(+SiLU) of
              // the image is finished here (what a stand-alone GroupNorm launch did: 10 launches of the 8x8 level)
              if ((ci & 1) == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) kept[i] = cur[i];
              } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const ConvTctExtra::Norm& q = x.no[k];
                  if (q.out == nullptr) continue;  // uniform
                  float t1 = s1, t2 = s2;
                  for (int o = 1; o < q.cpg; o <<= 1) {
                    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
                  }
                  const float inv_cnt = 1.0f / (64.0f * q.cpg);
                  const float mean = t1 * inv_cnt;
                  const float var = fmaxf(t2 * inv_cnt - mean * mean, 0.f);
                  const float rs = rsqrtf(var + q.eps);
                  float aa = rs * ng[k], bb = nb[k] - mean * rs * ng[k];
                  if (q.scale) {
                    const long long r = static_cast<long long>(q.ss_rows == 1 ? 0 : cur_n) * q.ss_ld;
                    const float sc = 1.f + __ldg(q.scale + r + ch), sh = __ldg(q.shift + r + ch);
                    aa *= sc;
                    bb = bb * sc + sh;
                  }
                  __nv_bfloat16* yp = q.out + static_cast<long long>(pix0 - 32) * cmod + chm;
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    float lo, hi;
                    unpack_bf16x2(i < 16 ? kept[i & 15] : cur[i & 15], lo, hi);
                    float y0 = fmaf(lo, aa, bb), y1 = fmaf(hi, aa, bb);
                    if (q.silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
                    const uint32_t yu = pack_bf16x2(y0, y1);
                    yp[(2 * i) * cmod] = __ushort_as_bfloat16(static_cast<unsigned short>(yu & 0xffffu));
                    yp[(2 * i + 1) * cmod] = __ushort_as_bfloat16(static_cast<unsigned short>(yu >> 16));
                  }
                }
              }
            }
            continue;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint32_t u = pack_bf16x2(f[i], f[i + 1]);
            r[i] = __ushort_as_bfloat16(static_cast<unsigned short>(u & 0xffffu));
            r[i + 1] = __ushort_as_bfloat16(static_cast<unsigned short>(u >> 16));
            float lo, hi;
            unpack_bf16x2(u, lo, hi);
            s1 += lo + hi;
            s2 = fmaf(lo, lo, s2);
            s2 = fmaf(hi, hi, s2);
          }
        } else {
          // ---- general path: ragged tail and / or several images per chunk (L = 16, 8, 4, ...) ----
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((i & (L - 1)) == 0 || i == 0) {  // warp-uniform
              const int n = (pix0 + i) >> x.log2_l;
              if (pix0 + i < total_pix && n != cur_n) enter_image(n);
            }
            const float val = __uint_as_float(v[i]) + bt + (has_add ? f[i] : 0.f);
            r[i] = __float2bfloat16_rn(val);
            if (pix0 + i < total_pix) {
              const float rf = __bfloat162float(r[i]);
              s1 += rf;
              s2 = fmaf(rf, rf, s2);
            }
          }
          if (which != 2) {
            __nv_bfloat16* op = obase + static_cast<long long>(pix0) * cmod + chm;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (pix0 + i < total_pix) op[i * cmod] = r[i];
            continue;
          }
        }
        // V^T [n][C][L]: this thread's channel row, runs of 8 pixels = 16 bytes (L % 8 == 0)
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int pg = pix0 + 8 * g8;
          if (pg < total_pix) {
            const int n = pg >> x.log2_l;
            const int l = pg - (n << x.log2_l);
            uint4 o;
            o.x = pack2(r[8 * g8 + 0], r[8 * g8 + 1]); o.y = pack2(r[8 * g8 + 2], r[8 * g8 + 3]);
            o.z = pack2(r[8 * g8 + 4], r[8 * g8 + 5]); o.w = pack2(r[8 * g8 + 6], r[8 * g8 + 7]);
            *reinterpret_cast<uint4*>(obase + (static_cast<long long>(n) * cmod + chm) * L + l) = o;
          }
        }
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_cluster(mapa_u32(&acc_empty[stage], 0));  // the leader's MMA thread waits for both CTAs
      else mbar_arrive(&acc_empty[stage]);
      if (warp == 2 && lane == 0) trace_ev(x.trace, 2, 2 * t_it + 1);
      flush_stats();
    }
  }
#undef DMME_TCT_COORDS

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // neither CTA may leave while the other can still signal its barriers / read its tiles
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, 2 * kAccCols);
    else tmem_dealloc(tmem_base, 2 * kAccCols);
  }
}

// activation map: NHWC bf16 [n][h][w][c] seen as (c', w', parity, h', n)
static int make_act_map(CUtensorMap* out, const void* ptr, int n, int h, int w, int c, int stride, int bw, int bh,
                        int bni) {
  uint64_t dims[5], strides[4];
  uint32_t box[5] = {64u, (uint32_t)bw, 1u, (uint32_t)bh, (uint32_t)bni};
  const uint64_t e = 2;
  if (stride == 1) {
    dims[0] = c; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = n;
    strides[0] = (uint64_t)c * e;          // w
    strides[1] = (uint64_t)w * c * e;      // dummy parity dim
    strides[2] = (uint64_t)w * c * e;      // h
    strides[3] = (uint64_t)h * w * c * e;  // n
  } else {
    dims[0] = 2 * (uint64_t)c; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = n;
    strides[0] = 2 * (uint64_t)c * e;      // w' (two pixels)
    strides[1] = (uint64_t)w * c * e;      // row parity
    strides[2] = 2 * (uint64_t)w * c * e;  // h' (two rows)
    strides[3] = (uint64_t)h * w * c * e;  // n
  }
  return encode_map(out, ptr, 5, dims, strides, box);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

bool conv_tc_supported(const dmme_conv_desc& d) {
  if (d.act_dtype != DMME_BF16 || d.in_layout != DMME_IN_NHWC) return false;
  if (d.out_layout != DMME_OUT_NHWC && d.out_layout != DMME_OUT_QKV) return false;
  if (d.upsample && d.upsample != 3) return false;
  if (d.upsample == 3) {
    // sub-pixel phases of nearest x2 + conv3x3: transposed kernel only, plain bias epilogue, whole 32-pixel chunks
    if (d.ksize != 3 || d.stride != 1 || d.out_layout != DMME_OUT_NHWC || d.c1 || d.rc0 || d.rc1 || d.temb || d.addend)
      return false;
    if (d.cout % 128 || d.w_in > 128 || d.h_in * d.w_in < 32) return false;
  }
  if (!(d.ksize == 3 || (d.ksize == 1 && d.stride == 1))) return false;
  if (d.stride != 1 && d.stride != 2) return false;
  if (d.c0 <= 0 || d.c0 % 64 || d.c1 % 64 || d.rc0 % 64 || d.rc1 % 64) return false;
  if (d.cout % 64) return false;
  if (d.stride == 2 && ((d.h_in | d.w_in) & 1)) return false;
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  if (!is_pow2(ho) || !is_pow2(wo)) return false;
  if (d.out_layout == DMME_OUT_QKV) {
    if (d.cout % 3) return false;
    const int c = d.cout / 3;
    if (c % 64) return false;
    if ((ho * wo) % 128 && (128 % (ho * wo))) return false;
  }
  return true;
}

constexpr int kWsStages = 6;
constexpr int kWsSlabMax = 128 * 1024;  // + 6 x 16 KB ring + alignment = 225 KB of the 227 KB a CTA may take

template <int BN, int STAGES, bool WS>
static int launch_conv_tc(const ConvTcParams& p, int m_tiles, cudaStream_t stream) {
  constexpr int smem = WS ? STAGES * kABytes + kWsSlabMax + 1024 : STAGES * stage_bytes<BN>() + 1024;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int sm_count = device_sm_count();
  ConvTcParams q = p;
  q.m_tiles = m_tiles;
  q.n_tiles = p.cout / BN;
  const int total = q.m_tiles * q.n_tiles;
  // persistent: one CTA per SM (it owns the whole TMEM: two accumulator stages), tiles dealt round-robin
  int grid = total < sm_count ? total : sm_count;
  if (WS) grid -= grid % q.n_tiles;  // every CTA stays on one n tile (n is the fastest tile index)
  cudaError_t e = launch_pdl(conv_tc_kernel<BN, STAGES, WS>, dim3(grid), dim3(kConvThreads), smem, stream, q);
  return check_launch_err(e, "conv_tc_kernel");
}

static long long* g_conv_trace = nullptr;
// cta_group::2 variant: 0 never (default: measured 34.8 vs 31.9 us at 8x8 and 98 vs 93 us at 16x16 -- the mainloop is paced
// by the SS-mode MMA, not by the operand feed the pairing halves), 1 where it applies, 2 wherever supported (tests)
static int g_tct_pair_mode = 0;
static int g_tct_mode = 1;  // 1: transposed kernel where it applies, 0: never (A/B measurements), 2: wherever supported
static int g_sm_count_tc = 0;

// pixel-tile width of the transposed kernel for this problem, 0 = use conv_tc_kernel
static int tct_tile_pixels(const dmme_conv_desc& d) {
  if (d.upsample == 3) {  // four phase units per (pixel tile, channel tile)
    const long long tp = static_cast<long long>(d.n) * d.h_in * d.w_in;
    return ceil_div_ll(tp, 256) * (d.cout / 128) * 4 >= 120 ? 256 : 128;
  }
  if (g_tct_mode == 0) return 0;
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  if (d.cout % 128 || wo > 256) return 0;  // a tile is at least one run of 128 / 256 pixels of a row
  if (d.out_layout == DMME_OUT_QKV && ((d.cout / 3) % 128 || (ho * wo) % 8)) return 0;
  const long long total_pix = static_cast<long long>(d.n) * ho * wo;
  const int n_tiles = d.cout / 128;
  if (g_tct_pair_mode == 2 && d.cout % 256 == 0 && d.out_layout == DMME_OUT_NHWC) return 256;  // tests: force the pair kernel
  if (wo <= 256 && ceil_div_ll(total_pix, 256) * n_tiles >= 120) return 256;
  if (ceil_div_ll(total_pix, 128) * n_tiles >= 100 || g_tct_mode == 2) return 128;
  return 0;
}

// ---- split-K plan ----------------------------------------------------------------------------------------------------
// The 4x4 / 8x8 levels (and every level at the small per-GPU batches of the strong-scaling run) have fewer pixel x channel
// tiles than SMs, and one CTA then walks the whole K loop alone at the ~45 B/clk a single SM ingests from L2 (measured:
// 19 us for 36 k-blocks whether 16 or 128 CTAs run).  Splitting K makes every (tile, K slice) a work unit; a finishing
// pass sums the slices and -- because it sees whole images -- also applies the consumers' GroupNorm(+SiLU).
static int g_splitk_mode = 1;
// 1: 4x4 maps reduce the K slices inside a thread-block cluster (one launch) where the plan has 2 / 4 / 8 slices; 0: partial
// tiles + finishing pass; A/B and tests: 2 = plan restricted to the cluster shapes, 3 = that plan WITHOUT the cluster
static int g_splitk_cluster = 1;
struct SplitPlan { int np, split; };
static int g_splitk_force_np = 0, g_splitk_force_split = 0;  // A/B (dmme_debug_force_splitk_4x4)

// cluster split-K: 4x4 and 8x8 maps, 2 / 4 / 8 slices, every CTA finishing whole images (8x8: the cost model's own plans at
// 128 / 64 images per GPU are such shapes: 2.20 -> 2.04 and 1.48 -> 1.45 ms per step).  The plan is NOT bent towards these
// shapes beyond the one-wave rule in splitk_plan: restricting the slice counts to 2 / 4 / 8 measured slower at the small per-GPU batches
// (batch 128: 2.42 vs 2.23 ms per step, batch 32: 1.19 vs 1.13 -- the cost model prefers 9 - 16 slices of 128-pixel tiles
// there), so the in-cluster reduction is taken where the plan already has such a shape (batch 256: 4 slices of 256 pixels)
static bool splitk_cluster_ok(const dmme_conv_desc& d, const SplitPlan& plan) {
  if (!(g_splitk_cluster == 1 || g_splitk_cluster == 2) || plan.split < 2) return false;
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  if (ho * wo != 16 && ho * wo != 64) return false;
  if (plan.split != 2 && plan.split != 4 && plan.split != 8) return false;
  return (plan.np / plan.split) % (ho * wo) == 0;  // every CTA finishes whole images
}

static SplitPlan splitk_plan(const dmme_conv_desc& d) {
  SplitPlan none{0, 1};
  if (g_splitk_mode == 0 || !conv_tc_supported(d)) return none;
  if (d.ksize != 3 || d.upsample || d.out_layout != DMME_OUT_NHWC || d.cout % 128) return none;
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  if (wo > 8 || (ho * wo) > 64) return none;  // the 8x8 level and below (measured: at 16x16 GEMM + finish only ties the unsplit conv + norm)
  const long long total_pix = static_cast<long long>(d.n) * ho * wo;
  if (total_pix * d.cout > (1ll << 26)) return none;
  const int sm = device_sm_count();
  const int nkb = 9 * ((d.c0 + d.c1) / 64) + (d.rc0 + d.rc1) / 64;
  const int n_tiles = d.cout / 128;
  auto unit_clocks = [](int kb, int bytes_per_kb) { return kb * (bytes_per_kb / 48.0) + 2500.0; };
  // what runs without the workspace: the transposed kernel where it has a unit for most SMs, else 128 px x 64 ch tiles
  double base_cost;
  {
    const int np0 = tct_tile_pixels(d);
    if (np0) base_cost = static_cast<double>(ceil_div_ll(ceil_div_ll(total_pix, np0) * n_tiles, sm)) * unit_clocks(nkb, np0 * 128 + 16384);
    else base_cost = static_cast<double>(ceil_div_ll(ceil_div_ll(total_pix, 128) * (d.cout / 64), sm)) * unit_clocks(nkb, 24576);
  }
  if ((ho * wo == 16 || ho * wo == 64) && g_splitk_force_np > 0 && g_splitk_force_split > 1 && 2 * g_splitk_force_split <= nkb)
    return SplitPlan{g_splitk_force_np, g_splitk_force_split};  // A/B and tests: a fixed shape for the 4x4 / 8x8 maps
  if (g_splitk_cluster <= 1 && ho * wo == 16 && nkb >= 8) {  // (mode 0 keeps the plan and only executes it as GEMM + finishing pass)
    // 4x4 maps: four slices as one cluster of four whenever that is about one wave of CTAs (batch 256: 256-pixel tiles,
    // batch 128: 128-pixel tiles -- measured 2.20 vs 2.25 ms per step there against the cost model's nine slices + finishing
    // pass; two or eight slices, and four at other batches, measured slower: tools/sweep_step.py with DMME_SPLITK_FORCE_4X4)
    for (int np = 256; np >= 128; np >>= 1) {
      const long long units = ceil_div_ll(total_pix, np) * n_tiles * 4;
      if (units >= 100 && units <= sm) return SplitPlan{np, 4};
    }
  }
  SplitPlan best = none;
  double best_cost = 0.8 * base_cost + 6000.0;  // + the stand-alone GroupNorm launch the finishing pass replaces
  if (g_splitk_mode == 2) best_cost = 1e30;      // tests: split wherever the kernel supports it
  for (int np = 128; np <= 256; np *= 2) {
    if (np == 256 && wo > 256) continue;
    const long long base_units = ceil_div_ll(total_pix, np) * n_tiles;
    for (int split = 2; split <= 16 && 2 * split <= nkb; ++split) {
      // A/B (dmme_set_conv_splitk_cluster(2)): only the shapes the in-cluster reduction takes
      if (g_splitk_cluster >= 2 && ho * wo == 16 && !((split == 2 || split == 4 || split == 8) && (np / split) % 16 == 0)) continue;
      const long long units = base_units * split;
      if (units > 2 * sm) break;
      const double gemm = static_cast<double>(ceil_div_ll(units, sm)) * unit_clocks(ceil_div(nkb, split), np * 128 + 16384);
      const double finish = 4000.0 + static_cast<double>(split) * total_pix * d.cout * 4.0 / (sm * 64.0);
      if (gemm + finish < best_cost) { best_cost = gemm + finish; best = SplitPlan{np, split}; }
    }
  }
  return best;
}

// The unsplit transposed kernel can finish the consumers' GroupNorm(+SiLU) in its epilogue on 8x8 maps with 128 / 256 output
// channels: a warp's pixel range is whole images (two 32-pixel chunks each), the channel stride is a compile-time constant.
bool conv_tct_epilogue_norm(const dmme_conv_desc& d) {
  if (!conv_tc_supported(d) || d.ksize != 3 || d.upsample || d.out_layout != DMME_OUT_NHWC) return false;
  if (d.cout != 128 && d.cout != 256) return false;
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  if (ho * wo != 64) return false;
  return tct_tile_pixels(d) != 0;
}

long long conv_splitk_workspace(const dmme_conv_desc& d) {
  const SplitPlan plan = splitk_plan(d);
  if (plan.split <= 1) return 0;
  const long long total_pix = static_cast<long long>(d.n) * (d.h_in / d.stride) * (d.w_in / d.stride);
  return static_cast<long long>(plan.split) * total_pix * d.cout * 4;
}

int conv_splitk_finish(const dmme_conv_desc& d, int split, cudaStream_t stream);  // conv_splitk.cu

// split-K with the slices of a tile as one thread-block cluster (see conv_tct_kernel, CLUSTER): grid = units
template <int CLPX>
static int launch_conv_tct_cluster(ConvTcParams& p, const ConvTctExtra& x, int smem, cudaStream_t stream) {
  auto kern = conv_tct_kernel<false, 0, false, true, false, true, CLPX>;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_tct (cluster split-K): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles * x.split;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(total);
  cfg.blockDim = dim3((2 + tct_epi_warps<false>()) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = x.split;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, x);
  return check_launch_err(e, "conv_tct_kernel (cluster split-K)");
}

template <bool WS, int CMOD, bool PAIR = false, bool SPLIT = false, bool NORM = false>
static int launch_conv_tct(ConvTcParams& p, const ConvTctExtra& x, int smem, cudaStream_t stream) {
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tct_kernel<WS, CMOD, PAIR, SPLIT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_tct: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles * (x.subpix ? 4 : 1) * (SPLIT ? x.split : 1);
  if (PAIR) {
    // one cluster of two CTAs (two SMs) per work unit
    const int pairs = total < g_sm_count_tc / 2 ? total : g_sm_count_tc / 2;
    cudaError_t e = launch_pdl_pair(conv_tct_kernel<WS, CMOD, PAIR>, dim3(2 * pairs), dim3((2 + tct_epi_warps<WS>()) * 32), smem, stream, p, x);
    return check_launch_err(e, "conv_tct_kernel (cta_group::2)");
  }
  int grid = total < g_sm_count_tc ? total : g_sm_count_tc;
  if (WS) grid -= grid % p.n_tiles;
  cudaError_t e = launch_pdl(conv_tct_kernel<WS, CMOD, PAIR, SPLIT, NORM>, dim3(grid), dim3((2 + tct_epi_warps<WS>()) * 32), smem, stream, p, x);
  return check_launch_err(e, SPLIT ? "conv_tct_kernel (split-K)" : (NORM ? "conv_tct_kernel (epilogue norm)" : "conv_tct_kernel"));
}

int conv_tc_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_tc_supported(d), DMME_E_SHAPE, "conv_tc: unsupported shape/layout");
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  g_sm_count_tc = device_sm_count();
  SplitPlan plan{0, 1};
  if (d.splitk_ws != nullptr) {
    plan = splitk_plan(d);
    if (plan.split > 1 && d.splitk_ws_bytes < static_cast<long long>(plan.split) * d.n * ho * wo * d.cout * 4) plan = SplitPlan{0, 1};
  }
  const bool has_out_norm = d.out_norm[0].out != nullptr || d.out_norm[1].out != nullptr;
  const bool epi_norm = has_out_norm && plan.split <= 1 && conv_tct_epilogue_norm(d);
  DMME_REQUIRE(!has_out_norm || plan.split > 1 || epi_norm, DMME_E_UNSUPPORTED,
               "conv_tc: out_norm needs the split-K path (ask dmme_conv2d_splitk_workspace and pass splitk_ws) or an 8x8 map "
               "on the transposed kernel (ask dmme_conv2d_epilogue_norm)");
  const int np = plan.split > 1 ? plan.np : tct_tile_pixels(d);
  const int tile_px = np ? np : 128;
  p.bw = wo < tile_px ? wo : tile_px;
  p.bh = ho < tile_px / p.bw ? ho : tile_px / p.bw;
  p.bni = tile_px / (p.bw * p.bh);
  p.tiles_x = wo / p.bw;
  p.tiles_y = ho / p.bh;
  const int m_tiles = p.tiles_x * p.tiles_y * ceil_div(d.n, p.bni);
  p.chunks0 = d.c0 / 64; p.chunks1 = d.c1 / 64;
  p.rchunks0 = d.rc0 / 64; p.rchunks1 = d.rc1 / 64;
  p.c0 = d.c0; p.c1 = d.c1;
  p.taps = d.upsample == 3 ? 4 : d.ksize * d.ksize; p.stride = d.stride;
  p.n = d.n; p.ho = ho; p.wo = wo;
  p.cout = d.cout;
  p.bias = d.bias; p.temb = d.temb; p.temb_rows = d.temb_rows; p.temb_ld = d.temb_ld;
  p.addend = static_cast<const __nv_bfloat16*>(d.addend);
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.out2 = static_cast<__nv_bfloat16*>(d.out2);
  p.out3 = static_cast<__nv_bfloat16*>(d.out3);
  p.out_mode = d.out_layout;
  p.stats = d.out_layout == DMME_OUT_NHWC ? d.stats : nullptr;
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_tc: null src0/weight/out");
  DMME_REQUIRE(d.c1 == 0 || d.src1, DMME_E_BADARG, "conv_tc: c1 > 0 but src1 is null");
  DMME_REQUIRE(d.rc0 == 0 || d.res0, DMME_E_BADARG, "conv_tc: rc0 > 0 but res0 is null");
  DMME_REQUIRE(d.rc1 == 0 || d.res1, DMME_E_BADARG, "conv_tc: rc1 > 0 but res1 is null");
  DMME_REQUIRE(d.temb == nullptr || (d.temb_ld % 4 == 0), DMME_E_SHAPE, "conv_tc: temb_ld must be a multiple of 4");
  DMME_REQUIRE(d.out_layout != DMME_OUT_QKV || (d.out2 && d.out3), DMME_E_BADARG, "conv_tc: QKV needs out2/out3");

  int rc;
  if ((rc = make_act_map(&p.a[0], d.src0, d.n, d.h_in, d.w_in, d.c0, d.stride, p.bw, p.bh, p.bni))) return rc;
  if (d.c1 && (rc = make_act_map(&p.a[1], d.src1, d.n, d.h_in, d.w_in, d.c1, d.stride, p.bw, p.bh, p.bni))) return rc;
  if (d.rc0 && (rc = make_act_map(&p.a[2], d.res0, d.n, ho, wo, d.rc0, 1, p.bw, p.bh, p.bni))) return rc;
  if (d.rc1 && (rc = make_act_map(&p.a[3], d.res1, d.n, ho, wo, d.rc1, 1, p.bw, p.bh, p.bni))) return rc;

  const uint64_t ktot = (uint64_t)p.taps * (d.c0 + d.c1) + d.rc0 + d.rc1;
  if (np) {
    // transposed kernel: 128 output channels x np pixels per unit
    uint64_t dims[2] = {ktot, (uint64_t)d.cout * (d.upsample == 3 ? 4 : 1)};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64u, 128u};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
    p.m_tiles = m_tiles;
    p.n_tiles = d.cout / 128;
    ConvTctExtra x;
    x.np = np;
    x.total_pix = static_cast<long long>(d.n) * ho * wo;
    DMME_REQUIRE(x.total_pix + 256 < (1ll << 31), DMME_E_SHAPE, "conv_tct: more than 2^31 output pixels");
    x.trace = g_conv_trace;
    x.log2_l = 0;
    while ((1 << x.log2_l) < ho * wo) ++x.log2_l;
    x.subpix = d.upsample == 3 ? 1 : 0;
    x.log2_w = 0;
    while ((1 << x.log2_w) < wo) ++x.log2_w;
    x.split = 1; x.partial = nullptr; x.split_stride = 0;
    memset(x.no, 0, sizeof(x.no));
    const int budget = 226 * 1024 - 1024;
    if (plan.split > 1) {
      x.split = plan.split;
      x.partial = static_cast<float*>(d.splitk_ws);
      x.split_stride = x.total_pix * d.cout;
      p.bias = nullptr; p.temb = nullptr; p.addend = nullptr; p.stats = nullptr;  // applied by the finishing pass
      x.stage_bytes = np * 128 + 128 * 128;
      const int st = budget / x.stage_bytes;
      x.stages = st > kTctMaxStages ? kTctMaxStages : st;
      const int smem_s = x.stages * x.stage_bytes + 1024;
      if (splitk_cluster_ok(d, plan) &&
          static_cast<long long>(plan.split - 1) * (np / plan.split) * 512 <= static_cast<long long>(x.stages) * x.stage_bytes) {
        // the slices meet inside a cluster: the finishing pass's terms stay with the kernel
        p.bias = d.bias; p.temb = d.temb; p.addend = static_cast<const __nv_bfloat16*>(d.addend); p.stats = d.stats;
        for (int k = 0; k < 2; ++k) {
          const dmme_out_norm& sn = d.out_norm[k];
          if (sn.out == nullptr) continue;
          DMME_REQUIRE(sn.cpg >= 1 && sn.cpg <= 32 && 32 % sn.cpg == 0, DMME_E_SHAPE,
                       "conv_tct: out_norm channels per group must divide 32 (got %d)", sn.cpg);
          DMME_REQUIRE(sn.out != d.out, DMME_E_BADARG, "conv_tct: out_norm[%d].out aliases out", k);
          DMME_REQUIRE(sn.scale == nullptr || sn.shift != nullptr, DMME_E_BADARG, "conv_tct: out_norm scale without shift");
          x.no[k].out = static_cast<__nv_bfloat16*>(sn.out);
          x.no[k].gamma = sn.gamma; x.no[k].beta = sn.beta; x.no[k].scale = sn.scale; x.no[k].shift = sn.shift;
          x.no[k].ss_rows = sn.ss_rows; x.no[k].ss_ld = sn.ss_ld; x.no[k].cpg = sn.cpg; x.no[k].silu = sn.silu; x.no[k].eps = sn.eps;
        }
        return ho * wo == 16 ? launch_conv_tct_cluster<16>(p, x, smem_s, stream) : launch_conv_tct_cluster<64>(p, x, smem_s, stream);
      }
      if ((rc = launch_conv_tct<false, 0, false, true>(p, x, smem_s, stream))) return rc;
      return conv_splitk_finish(d, plan.split, stream);
    }
    const long long slab = static_cast<long long>(ktot / 64) * 128 * 128;
    const bool ws = p.taps == 1 && slab + 2 * np * 128 <= budget && p.n_tiles <= g_sm_count_tc;
    // cta_group::2: 256 output channels x 256 pixels per CTA pair when that still gives most SM pairs a unit
    const bool pair = g_tct_pair_mode != 0 && !ws && !x.subpix && np == 256 && d.cout % 256 == 0 && d.out_layout == DMME_OUT_NHWC &&
                      (static_cast<long long>(m_tiles) * (d.cout / 256) >= 56 || g_tct_pair_mode == 2);
#ifdef DMME_EXPERIMENTAL
    if (pair) {
      // half-tile boxes: split along the tile's slowest dimension
      const int hbni = p.bni >= 2 ? p.bni / 2 : p.bni, hbh = p.bni >= 2 ? p.bh : p.bh / 2;
      if ((rc = make_act_map(&p.a[0], d.src0, d.n, d.h_in, d.w_in, d.c0, d.stride, p.bw, hbh, hbni))) return rc;
      if (d.c1 && (rc = make_act_map(&p.a[1], d.src1, d.n, d.h_in, d.w_in, d.c1, d.stride, p.bw, hbh, hbni))) return rc;
      if (d.rc0 && (rc = make_act_map(&p.a[2], d.res0, d.n, ho, wo, d.rc0, 1, p.bw, hbh, hbni))) return rc;
      if (d.rc1 && (rc = make_act_map(&p.a[3], d.res1, d.n, ho, wo, d.rc1, 1, p.bw, hbh, hbni))) return rc;
      p.n_tiles = d.cout / 256;
      x.stage_bytes = (np / 2) * 128 + 128 * 128;
      int st = budget / x.stage_bytes;
      x.stages = st > kTctMaxStages ? kTctMaxStages : st;
      const int smem2 = x.stages * x.stage_bytes + 1024;
      return d.cout == 256 ? launch_conv_tct<false, 256, true>(p, x, smem2, stream)
                           : launch_conv_tct<false, 0, true>(p, x, smem2, stream);
    }
#else
    (void)pair;  // the cta_group::2 variant measured no faster: built only with -DDMME_EXPERIMENTAL
#endif
    x.stage_bytes = np * 128 + (ws ? 0 : 128 * 128);
    int stages = static_cast<int>((budget - (ws ? slab : 0)) / x.stage_bytes);
    x.stages = stages > kTctMaxStages ? kTctMaxStages : stages;
    const int smem = x.stages * x.stage_bytes + (ws ? static_cast<int>(slab) : 0) + 1024;
    const int cmod = d.out_layout == DMME_OUT_QKV ? d.cout / 3 : d.cout;
    if (epi_norm) {
      for (int k = 0; k < 2; ++k) {
        const dmme_out_norm& sn = d.out_norm[k];
        if (sn.out == nullptr) continue;
        DMME_REQUIRE(sn.cpg >= 1 && sn.cpg <= 32 && 32 % sn.cpg == 0, DMME_E_SHAPE,
                     "conv_tct: out_norm channels per group must divide 32 (got %d)", sn.cpg);
        DMME_REQUIRE(sn.out != d.out, DMME_E_BADARG, "conv_tct: out_norm[%d].out aliases out", k);
        DMME_REQUIRE(sn.scale == nullptr || sn.shift != nullptr, DMME_E_BADARG, "conv_tct: out_norm scale without shift");
        x.no[k].out = static_cast<__nv_bfloat16*>(sn.out);
        x.no[k].gamma = sn.gamma; x.no[k].beta = sn.beta; x.no[k].scale = sn.scale; x.no[k].shift = sn.shift;
        x.no[k].ss_rows = sn.ss_rows; x.no[k].ss_ld = sn.ss_ld; x.no[k].cpg = sn.cpg; x.no[k].silu = sn.silu; x.no[k].eps = sn.eps;
      }
      DMME_REQUIRE(!ws, DMME_E_UNSUPPORTED, "conv_tct: epilogue norm on a weight-stationary launch");
      return cmod == 128 ? launch_conv_tct<false, 128, false, false, true>(p, x, smem, stream)
                         : launch_conv_tct<false, 256, false, false, true>(p, x, smem, stream);
    }
    if (cmod == 128) return ws ? launch_conv_tct<true, 128>(p, x, smem, stream) : launch_conv_tct<false, 128>(p, x, smem, stream);
    if (cmod == 256) return ws ? launch_conv_tct<true, 256>(p, x, smem, stream) : launch_conv_tct<false, 256>(p, x, smem, stream);
    // the IDDPM multi-head qkv projection writes one NHWC tensor of 3C channels (models/iddpm.py:38-39)
    if (cmod == 384 && ws) return launch_conv_tct<true, 384>(p, x, smem, stream);
    if (cmod == 768 && ws) return launch_conv_tct<true, 768>(p, x, smem, stream);
    return ws ? launch_conv_tct<true, 0>(p, x, smem, stream) : launch_conv_tct<false, 0>(p, x, smem, stream);
  }
  // widest N tile that divides cout (and, for q/k/v splitting, the per-tensor width) while the persistent grid still
  // has a tile for every SM: wider tiles re-read the activation tile less often
  int unit = d.out_layout == DMME_OUT_QKV ? d.cout / 3 : d.cout;
  int bn = 64;
  if (unit % 128 == 0 && (long long)m_tiles * (d.cout / 128) >= 120) bn = 128;
  if (unit % 256 == 0 && (long long)m_tiles * (d.cout / 256) >= 120) bn = 256;
  // 1x1 convs whose [bn][K] weight slab fits beside the activation ring keep it resident (weight stationary)
  const bool one_by_one = p.taps == 1 && d.rc0 + d.rc1 == 0;
  while (one_by_one && bn > 64 && ktot * bn * 2 > (uint64_t)kWsSlabMax) bn >>= 1;
  const bool ws = one_by_one && ktot * bn * 2 <= (uint64_t)kWsSlabMax && d.cout / bn <= 148;
  {
    uint64_t dims[2] = {ktot, (uint64_t)d.cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64u, (uint32_t)bn};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
  }
  // one CTA per SM: the TMA ring takes the whole shared memory (192 KB each)
  if (ws) {
    switch (bn) {
      case 256: return launch_conv_tc<256, kWsStages, true>(p, m_tiles, stream);
      case 128: return launch_conv_tc<128, kWsStages, true>(p, m_tiles, stream);
      default: return launch_conv_tc<64, kWsStages, true>(p, m_tiles, stream);
    }
  }
  switch (bn) {
    case 256: return launch_conv_tc<256, 4, false>(p, m_tiles, stream);
    case 128: return launch_conv_tc<128, 6, false>(p, m_tiles, stream);
    default: return launch_conv_tc<64, 8, false>(p, m_tiles, stream);
  }
}

}  // namespace dmme

// A/B measurement switch: 0 = never use the transposed tcgen05 kernel, 1 = default, 2 = wherever it is supported
extern "C" void dmme_set_conv_tct_mode(int mode) { dmme::g_tct_mode = mode; }
// A/B measurement switch: 0 = never split K, 1 = default (cost model), 2 = wherever supported (tests)
extern "C" void dmme_set_conv_splitk_mode(int mode) { dmme::g_splitk_mode = mode; }
// A/B and test switch: 1 (default) = 4x4 maps reduce their K slices inside a thread-block cluster, 0 = partial tiles + finishing pass
extern "C" void dmme_set_conv_splitk_cluster(int mode) { dmme::g_splitk_cluster = mode; }
// A/B measurements only: a fixed (tile pixels, slices) plan for split-K on 4x4 maps (0, 0 = the cost model's)
extern "C" void dmme_debug_force_splitk_4x4(int np, int split) { dmme::g_splitk_force_np = np; dmme::g_splitk_force_split = split; }
extern "C" int dmme_get_conv_tct_mode(void) { return dmme::g_tct_mode; }
// debugging: device buffer of 3 x 512 int64 that CTA 0 of the transposed kernel fills with clock64 timestamps
extern "C" void dmme_debug_set_conv_trace(long long* buf) { dmme::g_conv_trace = buf; }
// A/B measurement switch for the cta_group::2 (two-SM) variant of the transposed kernel: 0 off, 1 default, 2 forced
extern "C" void dmme_set_conv_pair_mode(int mode) { dmme::g_tct_pair_mode = mode; }
