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
  CUtensorMap a[4];  // src0, src1, res0, res1
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

template <int BN, int STAGES>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;      // one accumulator stage
  constexpr uint32_t kTmemCols = 2 * kAccCols;           // double-buffered: epilogue of tile i overlaps the MMAs of tile i+1
  constexpr int kStage = stage_bytes<BN>();

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

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      DMME_TILE_COORDS(t)
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], kStage);
        uint8_t* sa = ring + s * kStage;
        uint8_t* sb = sa + kABytes;
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
        tma_load_2d(sb, &p.b, &full_bar[s], kb * kBlockK, col0);
      }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      int it = 0, t_it = 0;
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
        const uint64_t bdesc = umma_desc_sw128(sa + kABytes);
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
        if (p.addend) {
          const uint4* ap = reinterpret_cast<const uint4*>(p.addend + pix * p.cout + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 a4 = __ldg(ap + j);
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
  if (d.upsample) return false;
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

template <int BN, int STAGES>
static int launch_conv_tc(const ConvTcParams& p, int m_tiles, cudaStream_t stream) {
  constexpr int smem = STAGES * stage_bytes<BN>() + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  ConvTcParams q = p;
  q.m_tiles = m_tiles;
  q.n_tiles = p.cout / BN;
  const int total = q.m_tiles * q.n_tiles;
  // persistent: one CTA per SM (it owns the whole TMEM: two accumulator stages), tiles dealt round-robin
  const int grid = total < sm_count ? total : sm_count;
  conv_tc_kernel<BN, STAGES><<<grid, kConvThreads, smem, stream>>>(q);
  return check_launch("conv_tc_kernel");
}

int conv_tc_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_tc_supported(d), DMME_E_SHAPE, "conv_tc: unsupported shape/layout");
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  p.bw = wo < 128 ? wo : 128;
  p.bh = ho < 128 / p.bw ? ho : 128 / p.bw;
  p.bni = 128 / (p.bw * p.bh);
  p.tiles_x = wo / p.bw;
  p.tiles_y = ho / p.bh;
  const int m_tiles = p.tiles_x * p.tiles_y * ceil_div(d.n, p.bni);
  p.chunks0 = d.c0 / 64; p.chunks1 = d.c1 / 64;
  p.rchunks0 = d.rc0 / 64; p.rchunks1 = d.rc1 / 64;
  p.c0 = d.c0; p.c1 = d.c1;
  p.taps = d.ksize * d.ksize; p.stride = d.stride;
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
  // widest N tile that divides cout (and, for q/k/v splitting, the per-tensor width) while the persistent grid still
  // has a tile for every SM: wider tiles re-read the activation tile less often
  int unit = d.out_layout == DMME_OUT_QKV ? d.cout / 3 : d.cout;
  int bn = 64;
  if (unit % 128 == 0 && (long long)m_tiles * (d.cout / 128) >= 120) bn = 128;
  if (unit % 256 == 0 && (long long)m_tiles * (d.cout / 256) >= 120) bn = 256;
  {
    uint64_t dims[2] = {ktot, (uint64_t)d.cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64u, (uint32_t)bn};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
  }
  // one CTA per SM: the TMA ring takes the whole shared memory (192 KB each)
  switch (bn) {
    case 256: return launch_conv_tc<256, 4>(p, m_tiles, stream);
    case 128: return launch_conv_tc<128, 6>(p, m_tiles, stream);
    default: return launch_conv_tc<64, 8>(p, m_tiles, stream);
  }
}

}  // namespace dmme
