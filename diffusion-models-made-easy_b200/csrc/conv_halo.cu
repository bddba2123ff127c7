// 3x3 stride-1 convolution on tcgen05 with the activation halo tile reused across the nine filter taps.
//
// conv_tc.cu fetches one [128 px][64 ch] A tile per tap: every activation byte crosses L2 -> SM nine times and
// the kernel is bound by that feed (measured ~8.7 TB/s of operand traffic at 25% tensor-pipe utilisation).
// Here the batch is viewed as one flat sequence of ZERO-PADDED pixels, G = n*(H+2)*(W+2) + (y+1)*(W+2) + (x+1).
// In that space a filter tap is a constant shift d = (r-1)*(W+2) + (s-1), so for a tile of 256 consecutive
// positions ONE halo tile [256 + 2(W+3) positions][64 ch] in shared memory serves all nine taps: the MMA's A
// descriptor is simply advanced by d rows (128 B each; the 128-byte-swizzle phase is a function of the absolute
// shared-memory address for both TMA writes and MMA reads, so row-granular shifts need no re-layout).  The halo tile is assembled by one TMA box per padded image row
// ([W+2 px][64 ch], out-of-bounds pixels / rows / images zero-filled = the conv padding).
// Outputs at padding positions are junk and masked in the epilogue (11% of the MMA work at 32x32, 21% at 16x16).
//
// Work unit: 128 output channels x 256 positions, computed TRANSPOSED: the weight tile [128 cout][64] is the
// MMA's M-side operand and the pixel rows are its N-side operand (D^T = W X^T, one M=128, N=256 instruction per
// 16 channels).  A 128x128 tile would make every MMA read 8 KB of shared memory per 64 cycles = the whole
// 128 B/cycle shared-memory bandwidth (measured: 35% tensor utilisation); 128x256 needs 96 B/cycle.  The
// accumulator therefore has TMEM lane = output channel and column = position; two accumulators (2 x 256
// columns) are double-buffered so the epilogue of unit i overlaps the MMAs of unit i+1.  Persistent CTAs (one
// per SM), static round-robin schedule.
// Warp roles: 0..7 = epilogue (two warps per TMEM lane quarter), 8 = halo-tile TMA producer, 9 = weight-tile
// TMA producer, 10 = MMA issuer / TMEM owner.  The single-thread issue loops sit in the HIGHEST warp ids because
// the SM's warp arbiter favours high ids: with the issuer in warp 1 the eight busy epilogue warps starved it
// (measured: 1186 cycles per tap iteration for 512 cycles of MMA work).
#include <cuda.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct ConvHaloParams {
  CUtensorMap a[4];  // src0, src1, res0, res1: box = one padded row [W+2 px][64 ch]
  CUtensorMap b;     // weights [cout][K] bf16, box [128][64]
  int chunks0, chunks1, rchunks0, rchunks1;
  int n, h, w, wp, pimg, nr;
  long long gtot;
  int m_tiles, n_tiles, cout;
  const float* bias;
  const float* temb;
  int temb_rows, temb_ld;
  const __nv_bfloat16* addend;
  __nv_bfloat16* out;
  long long* stats;
};

constexpr int kHaloTM = 256;
constexpr int kHaloBN = 128;
constexpr int kHaloASlot = 47 * 1024;  // >= nr * (W+2) * 128 bytes
constexpr int kHaloBSlot = kHaloBN * 128;
constexpr int kHaloAStages = 3;
constexpr int kHaloBStages = 4;
constexpr int kHaloSmem = kHaloAStages * kHaloASlot + kHaloBStages * kHaloBSlot + 1024;
constexpr int kHaloEpiWarps = 8;                       // two warps per TMEM lane quarter, alternating column chunks
constexpr int kHaloThreads = (kHaloEpiWarps + 3) * 32;
constexpr int kWarpProdA = kHaloEpiWarps, kWarpProdB = kHaloEpiWarps + 1, kWarpMma = kHaloEpiWarps + 2;

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kHaloAStages], a_empty[kHaloAStages];
  __shared__ __align__(8) uint64_t b_full[kHaloBStages], b_empty[kHaloBStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* abuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* bbuf = abuf + kHaloAStages * kHaloASlot;

  const int cchunks = p.chunks0 + p.chunks1;
  const int nck = cchunks + p.rchunks0 + p.rchunks1;
  const int units = p.m_tiles * p.n_tiles;
  const int row_bytes = p.wp * 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHaloAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kHaloBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kHaloEpiWarps * 32); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == kWarpProdA && lane == 0) {
    tma_prefetch_desc(&p.a[0]);
    if (p.chunks1) tma_prefetch_desc(&p.a[1]);
    if (p.rchunks0) tma_prefetch_desc(&p.a[2]);
    if (p.rchunks1) tma_prefetch_desc(&p.a[3]);
    tma_prefetch_desc(&p.b);
  }
  if (warp == kWarpMma) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == kWarpProdA) {
    // =========================== halo-tile producer ===========================
    if (lane == 0) {
      int a_it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int mt = u / p.n_tiles;
        const int g0 = mt * kHaloTM;
        for (int ck = 0; ck < nck; ++ck, ++a_it) {
          int which, cc, halo;
          if (ck < cchunks) {
            which = ck < p.chunks0 ? 0 : 1;
            cc = (which ? ck - p.chunks0 : ck) * 64;
            halo = p.wp + 1;
          } else {
            const int rk = ck - cchunks;
            which = rk < p.rchunks0 ? 2 : 3;
            cc = (which == 3 ? rk - p.rchunks0 : rk) * 64;
            halo = 0;
          }
          const int as = a_it % kHaloAStages;
          mbar_wait(&a_empty[as], ((a_it / kHaloAStages) & 1) ^ 1);
          mbar_expect_tx(&a_full[as], p.nr * row_bytes);
          const int pr0 = floor_div(g0 - halo, p.wp);
          uint8_t* dst = abuf + as * kHaloASlot;
          for (int i = 0; i < p.nr; ++i) {
            const int pr = pr0 + i;
            const int ni = floor_div(pr, p.h + 2);         // -1 or >= n: whole row out of bounds -> zeros
            const int yy = pr - ni * (p.h + 2) - 1;        // -1 or h: padding row -> zeros
            tma_load_5d(dst + i * row_bytes, &p.a[which], &a_full[as], cc, -1, 0, yy, ni);
          }
        }
      }
    }
  } else if (warp == kWarpProdB) {
    // =========================== weight-tile producer ===========================
    if (lane == 0) {
      int b_it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int nt = u % p.n_tiles;
        const int col0 = nt * kHaloBN;
        for (int ck = 0; ck < nck; ++ck) {
          const bool is_conv = ck < cchunks;
          const int ntaps = is_conv ? 9 : 1;
          const int kb0 = is_conv ? ck : 9 * cchunks + (ck - cchunks);
          const int kbs = is_conv ? cchunks : 0;
          for (int tap = 0; tap < ntaps; ++tap, ++b_it) {
            const int bs = b_it % kHaloBStages;
            mbar_wait(&b_empty[bs], ((b_it / kHaloBStages) & 1) ^ 1);
            mbar_expect_tx(&b_full[bs], kHaloBSlot);
            tma_load_2d(bbuf + bs * kHaloBSlot, &p.b, &b_full[bs], (kb0 + tap * kbs) * 64, col0);
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer ===========================
    // the whole warp walks the loop (converged waits); one elected lane issues the MMAs and their commits
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kHaloBN, kHaloTM);  // M = 128 output channels, N = 256 positions
      int a_it = 0, b_it = 0, u_it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++u_it) {
        const int mt = u / p.n_tiles;
        const int g0 = mt * kHaloTM;
        const int stage = u_it & 1;
        mbar_wait(&acc_empty[stage], ((u_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dtm = tmem_base + stage * kHaloTM;
        for (int ck = 0; ck < nck; ++ck, ++a_it) {
          const bool is_conv = ck < cchunks;
          const int halo = is_conv ? p.wp + 1 : 0;
          const int ntaps = is_conv ? 9 : 1;
          const int as = a_it % kHaloAStages;
          mbar_wait(&a_full[as], (a_it / kHaloAStages) & 1);
          tc_fence_after();
          const int pr0 = floor_div(g0 - halo, p.wp);
          const int rowbase = g0 - pr0 * p.wp;  // row of position g0 inside the halo slot
          const uint32_t a_addr = smem_u32(abuf + as * kHaloASlot);
          for (int tap = 0; tap < ntaps; ++tap, ++b_it) {
            const int bs = b_it % kHaloBStages;
            mbar_wait(&b_full[bs], (b_it / kHaloBStages) & 1);
            tc_fence_after();
            const int d = is_conv ? (tap / 3 - 1) * p.wp + (tap % 3 - 1) : 0;
            const uint64_t wdesc = umma_desc_sw128(smem_u32(bbuf + bs * kHaloBSlot));  // [128 cout][64]: M side
            // pixel rows [256][64] shifted by the tap: N side.  The start address is 128-byte (one pixel row)
            // aligned, not 1024: measured on B200, the MMA derives the swizzle phase from the absolute shared-memory
            // address exactly as TMA did when writing the rows, so the descriptor's base-offset field stays 0
            // (setting it to (addr >> 7) & 7 produces wrong results).
            const uint64_t xdesc = umma_desc_sw128(a_addr + static_cast<uint32_t>(rowbase + d) * 128u);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(dtm, wdesc + 2 * k, xdesc + 2 * k, idesc, (ck | tap | k) != 0 ? 1u : 0u);
              umma_commit(&b_empty[bs]);
              if (tap == ntaps - 1) {
                umma_commit(&a_empty[as]);
                if (ck == nck - 1) umma_commit(&acc_full[stage]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    // thread = one output channel (TMEM lane), columns = positions; warp (q, half) drains channels
    // [32q, 32q+32) x positions [128 half, 128 half + 128) of every unit.  Valid positions in increasing order map
    // to consecutive pixels, so per 32-position chunk one ballot gives the validity mask and the stores just walk
    // a pointer: ~10 instructions per position (the first version spent 66 on index arithmetic and was the
    // kernel's bottleneck).
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    const bool temb_per_image = p.temb && p.temb_rows != 1;
    int u_it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++u_it) {
      const int mt = u / p.n_tiles, nt = u - mt * p.n_tiles;
      const int ch = nt * kHaloBN + q * 32 + lane;
      const float bias_c = p.bias ? __ldg(p.bias + ch) : 0.f;
      const long long gs = static_cast<long long>(mt) * kHaloTM + half * 128;
      const int n_a = static_cast<int>(gs / p.pimg);                       // image of the first position
      const int boundary = static_cast<int>(static_cast<long long>(n_a + 1) * p.pimg - gs);  // first position of image n_a+1
      float bt = bias_c;
      if (p.temb && n_a < p.n) bt += __ldg(p.temb + static_cast<long long>(temb_per_image ? n_a : 0) * p.temb_ld + ch);
      float s1 = 0.f, s2 = 0.f, s1a = 0.f, s2a = 0.f;
      const int stage = u_it & 1;
      mbar_wait(&acc_full[stage], (u_it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        // lane l classifies position c + l; the ballot is the chunk's validity mask
        const long long g = gs + c + lane;
        const int n = static_cast<int>(g / p.pimg);
        const int rem = static_cast<int>(g - static_cast<long long>(n) * p.pimg);
        const int yy = rem / p.wp - 1, xx = rem - (yy + 1) * p.wp - 1;
        const bool ok = n < p.n && yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + static_cast<uint32_t>(stage * kHaloTM + half * 128 + c), v);
        tmem_ld_wait();
        const int brel = boundary - c;  // chunk-relative index at which the next image starts (may be out of range)
        if (mask == 0u) {  // warp-uniform: a chunk of pure padding; it may still contain the image boundary
          if (brel >= 0 && brel < 32) {
            s1a = s1; s2a = s2; s1 = 0.f; s2 = 0.f;
            if (temb_per_image && n_a + 1 < p.n) bt = bias_c + __ldg(p.temb + static_cast<long long>(n_a + 1) * p.temb_ld + ch);
          }
          continue;
        }
        const long long mypix = (static_cast<long long>(n) * p.h + yy) * p.w + xx;
        const long long pix0 = __shfl_sync(0xffffffffu, mypix, __ffs(mask) - 1);
        __nv_bfloat16* op = p.out + pix0 * p.cout + ch;
        const __nv_bfloat16* ap = p.addend ? p.addend + pix0 * p.cout + ch : nullptr;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i == brel) {  // crossed into image n_a + 1: park the first image's sums, switch the per-image temb
            s1a = s1; s2a = s2; s1 = 0.f; s2 = 0.f;
            if (temb_per_image && n_a + 1 < p.n) bt = bias_c + __ldg(p.temb + static_cast<long long>(n_a + 1) * p.temb_ld + ch);
          }
          if (mask & (1u << i)) {
            float val = __uint_as_float(v[i]) + bt;
            if (ap) { val += __bfloat162float(*ap); ap += p.cout; }
            const __nv_bfloat16 r = __float2bfloat16_rn(val);
            *op = r;
            op += p.cout;
            const float rf = __bfloat162float(r);
            s1 += rf;
            s2 = fmaf(rf, rf, s2);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[stage]);
      if (p.stats) {
        const bool crossed = boundary < 128;    // sums of image n_a are parked in (s1a, s2a), (s1, s2) belong to n_a + 1
        float fa1 = crossed ? s1a : s1, fa2 = crossed ? s2a : s2;
        float fb1 = crossed ? s1 : 0.f, fb2 = crossed ? s2 : 0.f;
        // micro-group = 4 adjacent channels = 4 adjacent lanes
        fa1 += __shfl_xor_sync(0xffffffffu, fa1, 1); fa2 += __shfl_xor_sync(0xffffffffu, fa2, 1);
        fb1 += __shfl_xor_sync(0xffffffffu, fb1, 1); fb2 += __shfl_xor_sync(0xffffffffu, fb2, 1);
        fa1 += __shfl_xor_sync(0xffffffffu, fa1, 2); fa2 += __shfl_xor_sync(0xffffffffu, fa2, 2);
        fb1 += __shfl_xor_sync(0xffffffffu, fb1, 2); fb2 += __shfl_xor_sync(0xffffffffu, fb2, 2);
        if ((lane & 3) == 0) {
          unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats);
          if (n_a < p.n) {
            unsigned long long* sa = st + (static_cast<long long>(n_a) * (p.cout >> 2) + (ch >> 2)) * 2;
            atomicAdd(sa, static_cast<unsigned long long>(__float2ll_rn(fa1 * kFix)));
            atomicAdd(sa + 1, static_cast<unsigned long long>(__float2ll_rn(fa2 * kFix)));
          }
          if (crossed && n_a + 1 < p.n) {
            unsigned long long* sb = st + (static_cast<long long>(n_a + 1) * (p.cout >> 2) + (ch >> 2)) * 2;
            atomicAdd(sb, static_cast<unsigned long long>(__float2ll_rn(fb1 * kFix)));
            atomicAdd(sb + 1, static_cast<unsigned long long>(__float2ll_rn(fb2 * kFix)));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int halo_rows(int wp) { return (kHaloTM + 3 * wp) / wp + 1; }

static int g_halo_mode = 1;     // 1: AUTO prefers this kernel, 0: AUTO never picks it (A/B measurements)
static int g_sm_count = 0;

bool conv_halo_supported(const dmme_conv_desc& d) {
  if (g_halo_mode == 0) return false;
  if (d.act_dtype != DMME_BF16 || d.in_layout != DMME_IN_NHWC || d.out_layout != DMME_OUT_NHWC) return false;
  if (d.upsample || d.ksize != 3 || d.stride != 1) return false;
  if (d.c0 <= 0 || d.c0 % 64 || d.c1 % 64 || d.rc0 % 64 || d.rc1 % 64 || d.cout % kHaloBN) return false;
  if (d.h_in < 16 || d.w_in < 16) return false;  // padded-position utilisation < 75% below 16x16
  const int wp = d.w_in + 2;
  if (wp > 256 || halo_rows(wp) * wp * 128 > kHaloASlot) return false;
  if (static_cast<long long>(d.n) * (d.h_in + 2) * wp > (1LL << 30)) return false;
  return true;
}

int conv_halo_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_halo_supported(d), DMME_E_SHAPE, "conv_halo: unsupported shape/layout");
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_halo: null src0/weight/out");
  DMME_REQUIRE(d.c1 == 0 || d.src1, DMME_E_BADARG, "conv_halo: c1 > 0 but src1 is null");
  DMME_REQUIRE(d.rc0 == 0 || d.res0, DMME_E_BADARG, "conv_halo: rc0 > 0 but res0 is null");
  DMME_REQUIRE(d.rc1 == 0 || d.res1, DMME_E_BADARG, "conv_halo: rc1 > 0 but res1 is null");
  DMME_REQUIRE(d.temb == nullptr || (d.temb_ld % 4 == 0), DMME_E_SHAPE, "conv_halo: temb_ld must be a multiple of 4");
  ConvHaloParams p;
  memset(&p, 0, sizeof(p));
  p.chunks0 = d.c0 / 64; p.chunks1 = d.c1 / 64; p.rchunks0 = d.rc0 / 64; p.rchunks1 = d.rc1 / 64;
  p.n = d.n; p.h = d.h_in; p.w = d.w_in; p.wp = d.w_in + 2; p.pimg = (d.h_in + 2) * p.wp;
  p.nr = halo_rows(p.wp);
  p.gtot = static_cast<long long>(d.n) * p.pimg;
  p.m_tiles = static_cast<int>((p.gtot + kHaloTM - 1) / kHaloTM);
  p.n_tiles = d.cout / kHaloBN;
  p.cout = d.cout;
  p.bias = d.bias; p.temb = d.temb; p.temb_rows = d.temb_rows; p.temb_ld = d.temb_ld;
  p.addend = static_cast<const __nv_bfloat16*>(d.addend);
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.stats = d.stats;

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
    uint32_t box[2] = {64u, (uint32_t)kHaloBN};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
  }
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem);
    if (e != cudaSuccess) {
      g_sm_count = 0;
      set_error("conv_halo: cudaFuncSetAttribute(%d bytes): %s", kHaloSmem, cudaGetErrorString(e));
      return (int)e;
    }
  }
  const int units = p.m_tiles * p.n_tiles;
  const int grid = units < g_sm_count ? units : g_sm_count;
  conv_halo_kernel<<<grid, kHaloThreads, kHaloSmem, stream>>>(p);
  return check_launch("conv_halo_kernel");
}

}  // namespace dmme

// A/B measurement switch: 0 = AUTO never uses the halo kernel, 1 = default
extern "C" void dmme_set_conv_halo_mode(int mode) { dmme::g_halo_mode = mode; }
extern "C" int dmme_get_conv_halo_mode(void) { return dmme::g_halo_mode; }
