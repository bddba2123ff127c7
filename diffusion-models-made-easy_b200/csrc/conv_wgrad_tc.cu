// Convolution weight gradient on tcgen05 tensor cores (stride-1 3x3 and 1x1 convs, bf16 NHWC activations).
//
//   dW[co][tap][ci] = sum over output pixels p of  G[p][co] * X[p + tap][ci]
//
// is a GEMM whose reduction dimension is the PIXEL axis, while both NHWC operands are contiguous along their channel
// axis: in UMMA terms both operands are MN-major.  A TMA box [64 channels][K pixels] with SWIZZLE_128B lands in shared
// memory as K rows of 128 bytes, which is exactly the canonical MN-major SWIZZLE_128B layout
// (((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in bf16 elements: 64 channels contiguous per row, 8-row groups SBO = 1024 B apart,
// 64-channel blocks LBO = K*128 B apart), so no transpose pass is needed: the instruction descriptor just sets the
// a_major / b_major bits.  A tap is a shifted TMA box of X whose out-of-bounds pixels are zero-filled by the hardware
// (= the conv padding), as in conv_tc.cu.
//
// Work item (one CTA): 128 output channels x N input channels (N = 128 or 64) x the taps of one filter row (3 taps,
// or the single tap of a 1x1 / fused-residual weight) x one slice of the pixel axis.  Per 64-pixel chunk the CTA loads
// one G tile (16 KB) and one X tile per tap; the accumulators (taps x N fp32 columns) stay in TMEM for the whole slice and
// are written once as an fp32 partial [slice][cout][K]; backward.cu's wgrad_reduce_kernel sums the slices in a fixed order
// (deterministic) and scatters into the OIHW gradient.  The bias gradient is a per-image pixel sum + column sum.
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct WgradTcParams {
  CUtensorMap g;       // grad_out [n][ho][wo][cout], box [64][bw][1][bh][bni]
  CUtensorMap x[4];    // src0, src1, res0, res1, same box geometry
  int c0, c1, rc0, rc1;
  int cout, kp;        // kp = row length of the partial (taps * (c0+c1) + rc0 + rc1 + 1)
  int taps;            // 9 or 1
  int bw, bh, bni, tiles_x, tiles_y;
  int chunks_total;    // 64-pixel chunks over the whole batch
  int chunks_per_slice;
  int ci_blocks;       // (c0 + c1) / N
  int res_blocks;      // (rc0 + rc1) / N
  int rows;            // filter rows: 3 or 1
  float* partial;      // [slices][cout][kp]
};

constexpr int kWgChunk = 64;                 // pixels per K chunk
constexpr int kWgTileBytes = kWgChunk * 128; // one [64 px][64 ch] box
constexpr int kWgThreads = 192;

// instruction descriptor: bf16 x bf16 -> fp32, both operands MN-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// shared-memory descriptor, MN-major, 128-byte swizzle: rows = K (pixels) of 128 B = 64 channels; 8-row groups 1024 B
// apart (SBO); 64-channel blocks `lbo_bytes` apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

template <int N>  // input channels per item: 128 or 64
__global__ void __launch_bounds__(kWgThreads) conv_wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
  constexpr int kXBytes = (N / 64) * kWgTileBytes;        // one tap's X tile
  constexpr int kGBytes = 2 * kWgTileBytes;               // G tile: 128 output channels
  constexpr int kStageBytes = kGBytes + 3 * kXBytes;      // up to 3 taps per item
  constexpr int kStages = N == 128 ? 3 : 4;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* ring = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  // ---- item decode: blockIdx.x = ((co_tile * blocks + block) * rows + row) ----
  const int blocks = p.ci_blocks + p.res_blocks;
  int item = blockIdx.x;
  const int row = item % p.rows; item /= p.rows;
  const int blk = item % blocks;
  const int co_tile = item / blocks;
  const bool is_res = blk >= p.ci_blocks;
  if (is_res && row != 0) return;  // residual weights have a single tap: only the first row's CTA works (uniform exit)
  const int ntaps = is_res ? 1 : (p.taps == 9 ? 3 : 1);
  int which, cc;                    // source tensor map and channel offset inside it
  int kcol;                         // first column of this item's taps in the partial row
  const int ctot = p.c0 + p.c1;
  if (!is_res) {
    const int ch = blk * N;
    which = ch < p.c0 ? 0 : 1;
    cc = which ? ch - p.c0 : ch;
    kcol = (row * (p.taps == 9 ? 3 : 1)) * ctot + ch;   // tap = row * 3 + s -> column tap * ctot + ch
  } else {
    const int ch = (blk - p.ci_blocks) * N;
    which = ch < p.rc0 ? 2 : 3;
    cc = which == 3 ? ch - p.rc0 : ch;
    kcol = p.taps * ctot + ch;
  }
  const int chunk0 = blockIdx.y * p.chunks_per_slice;
  int chunk1 = chunk0 + p.chunks_per_slice;
  if (chunk1 > p.chunks_total) chunk1 = p.chunks_total;
  const int nchunks = chunk1 - chunk0;  // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.g); tma_prefetch_desc(&p.x[which]); }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (int i = 0; i < nchunks; ++i) {
        const int s = i % kStages;
        mbar_wait(&empty_bar[s], ((i / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], kGBytes + ntaps * kXBytes);
        uint8_t* sg = ring + s * kStageBytes;
        const int c = chunk0 + i;
        const int tx = c % p.tiles_x, ty = (c / p.tiles_x) % p.tiles_y, ng = c / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = ng * p.bni;
        tma_load_5d(sg, &p.g, &full_bar[s], co_tile * 128, x0, 0, y0, n0);
        tma_load_5d(sg + kWgTileBytes, &p.g, &full_bar[s], co_tile * 128 + 64, x0, 0, y0, n0);
        for (int t = 0; t < ntaps; ++t) {
          const int dy = (!is_res && p.taps == 9) ? row - 1 : 0;
          const int dx = (!is_res && p.taps == 9) ? t - 1 : 0;
          uint8_t* sx = sg + kGBytes + t * kXBytes;
#pragma unroll
          for (int b = 0; b < N / 64; ++b)
            tma_load_5d(sx + b * kWgTileBytes, &p.x[which], &full_bar[s], cc + b * 64, x0 + dx, 0, y0 + dy, n0);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn(128, N);
      for (int i = 0; i < nchunks; ++i) {
        const int s = i % kStages;
        mbar_wait(&full_bar[s], (i / kStages) & 1);
        tc_fence_after();
        const uint32_t sg = smem_u32(ring + s * kStageBytes);
        for (int t = 0; t < ntaps; ++t) {
          const uint32_t sx = sg + kGBytes + t * kXBytes;
#pragma unroll
          for (int k = 0; k < kWgChunk / 16; ++k) {
            // 16 pixels further along K = 16 rows of 128 bytes
            const uint64_t adesc = umma_desc_mn_sw128(sg + k * 2048, kWgTileBytes);
            const uint64_t bdesc = umma_desc_mn_sw128(sx + k * 2048, kWgTileBytes);
            umma_bf16(tmem_base + t * N, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    // =========================== epilogue: TMEM -> fp32 partial ===========================
    const int q = warp & 3;
    const int co = co_tile * 128 + q * 32 + lane;
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    float* prow = p.partial + (static_cast<long long>(blockIdx.y) * p.cout + co) * p.kp;
    for (int t = 0; t < ntaps; ++t) {
      float* dst = prow + kcol + t * ctot;  // taps of one filter row are ctot columns apart (residual: single tap)
#pragma unroll 1
      for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(t * N + c), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[c + j] = __uint_as_float(v[j]);  // rows of the partial are kp floats: no 16 B alignment
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

bool conv_wgrad_tc_supported(const dmme_conv_desc& d) {
  if (d.act_dtype != DMME_BF16 || d.in_layout != DMME_IN_NHWC || d.out_layout != DMME_OUT_NHWC) return false;
  if (d.upsample || d.stride != 1 || (d.ksize != 1 && d.ksize != 3)) return false;
  if (d.cout % 128) return false;
  if (d.c0 <= 0 || d.c0 % 64 || d.c1 % 64 || d.rc0 % 64 || d.rc1 % 64) return false;
  if (!is_pow2(d.h_in) || !is_pow2(d.w_in)) return false;
  return true;
}

static int wg_block_n(const dmme_conv_desc& d) {
  return (d.c0 % 128 == 0 && d.c1 % 128 == 0 && d.rc0 % 128 == 0 && d.rc1 % 128 == 0) ? 128 : 64;
}

// Slices of the pixel axis.  Measured (IDDPM training step, batch 128, tools/prof_train.py --graph): slicing for 3 / 2 / 1
// waves of CTAs = 21.7 / 20.1 / 18.8 ms, exactly one wave rounded DOWN (148 / items slices) 17.6-18.3 ms, half a wave 18.7:
// every slice costs a prologue, a 196 KB partial tile written and re-read by the reduction, and a tail -- so the fewest
// slices that still give every SM a CTA win.  0 = that default; > 0: that many waves; < 0: 1 / |value| of a wave (A/B).
static int g_wgrad_waves = 0;

// slices of the pixel axis (see g_wgrad_waves), at least 4 chunks per slice
void conv_wgrad_tc_geometry(const dmme_conv_desc& d, int& items, int& chunks_total, int& chunks_per_slice, int& slices) {
  const int nb = wg_block_n(d);
  const int rows = d.ksize == 3 ? 3 : 1;
  const int blocks = (d.c0 + d.c1) / nb + (d.rc0 + d.rc1) / nb;
  items = (d.cout / 128) * blocks * rows;
  const int bw = d.w_in < kWgChunk ? d.w_in : kWgChunk;
  const int bh = d.h_in < kWgChunk / bw ? d.h_in : kWgChunk / bw;
  const int bni = kWgChunk / (bw * bh);
  chunks_total = (d.w_in / bw) * (d.h_in / bh) * ceil_div(d.n, bni);
  int want = g_wgrad_waves > 0 ? ceil_div(148 * g_wgrad_waves, items) : g_wgrad_waves == 0 ? 148 / items : 148 / (items * -g_wgrad_waves);
  if (want < 1) want = 1;
  int max_slices = chunks_total / 4;
  if (max_slices < 1) max_slices = 1;
  if (want > max_slices) want = max_slices;
  chunks_per_slice = ceil_div(chunks_total, want);
  slices = ceil_div(chunks_total, chunks_per_slice);
}

static int make_map(CUtensorMap* out, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bni) {
  uint64_t dims[5] = {(uint64_t)c, (uint64_t)w, 1, (uint64_t)h, (uint64_t)n};
  uint64_t strides[4] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
  uint32_t box[5] = {64u, (uint32_t)bw, 1u, (uint32_t)bh, (uint32_t)bni};
  return encode_map(out, ptr, 5, dims, strides, box);
}

template <int N>
static int launch_wgrad_tc(const WgradTcParams& p, int items, int slices, cudaStream_t stream) {
  constexpr int stages = N == 128 ? 3 : 4;
  constexpr int smem = stages * (2 * kWgTileBytes + 3 * (N / 64) * kWgTileBytes) + 1024;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("conv_wgrad_tc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  dim3 grid(items, slices);
  conv_wgrad_tc_kernel<N><<<grid, kWgThreads, smem, stream>>>(p);
  return check_launch("conv_wgrad_tc_kernel");
}

// fills partial[slices][cout][kp] for every column except the bias column (kp - 1)
int conv_wgrad_tc_partials(const dmme_conv_desc& d, const void* grad_out, float* partial, int& slices, cudaStream_t stream) {
  DMME_REQUIRE(conv_wgrad_tc_supported(d), DMME_E_SHAPE, "conv_wgrad_tc: unsupported shape/layout");
  WgradTcParams p;
  memset(&p, 0, sizeof(p));
  int items, chunks_total, cps;
  conv_wgrad_tc_geometry(d, items, chunks_total, cps, slices);
  const int nb = wg_block_n(d);
  p.c0 = d.c0; p.c1 = d.c1; p.rc0 = d.rc0; p.rc1 = d.rc1;
  p.cout = d.cout;
  p.taps = d.ksize * d.ksize;
  p.kp = p.taps * (d.c0 + d.c1) + d.rc0 + d.rc1 + 1;
  p.bw = d.w_in < kWgChunk ? d.w_in : kWgChunk;
  p.bh = d.h_in < kWgChunk / p.bw ? d.h_in : kWgChunk / p.bw;
  p.bni = kWgChunk / (p.bw * p.bh);
  p.tiles_x = d.w_in / p.bw; p.tiles_y = d.h_in / p.bh;
  p.chunks_total = chunks_total; p.chunks_per_slice = cps;
  p.ci_blocks = (d.c0 + d.c1) / nb; p.res_blocks = (d.rc0 + d.rc1) / nb;
  p.rows = d.ksize == 3 ? 3 : 1;
  p.partial = partial;
  int rc;
  if ((rc = make_map(&p.g, grad_out, d.n, d.h_in, d.w_in, d.cout, p.bw, p.bh, p.bni))) return rc;
  if ((rc = make_map(&p.x[0], d.src0, d.n, d.h_in, d.w_in, d.c0, p.bw, p.bh, p.bni))) return rc;
  if (d.c1 && (rc = make_map(&p.x[1], d.src1, d.n, d.h_in, d.w_in, d.c1, p.bw, p.bh, p.bni))) return rc;
  if (d.rc0 && (rc = make_map(&p.x[2], d.res0, d.n, d.h_in, d.w_in, d.rc0, p.bw, p.bh, p.bni))) return rc;
  if (d.rc1 && (rc = make_map(&p.x[3], d.res1, d.n, d.h_in, d.w_in, d.rc1, p.bw, p.bh, p.bni))) return rc;
  return nb == 128 ? launch_wgrad_tc<128>(p, items, slices, stream) : launch_wgrad_tc<64>(p, items, slices, stream);
}

}  // namespace dmme

// A/B switch: waves of CTAs the tcgen05 weight-gradient kernel slices the pixel axis for (default 0 = one wave, rounded down)
extern "C" void dmme_set_wgrad_waves(int waves) { dmme::g_wgrad_waves = waves; }
