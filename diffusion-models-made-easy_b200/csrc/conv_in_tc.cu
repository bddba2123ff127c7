// Input convolution (models/ddpm.py:219, Conv 3 -> 128 on the fp32 NCHW image) on tcgen05 for 32 x 32 images.
//
// K = 27 is too short for an implicit GEMM over 64-channel rows, so the operand is built: for every output position the
// 27 taps of the image patch are written as ONE 128-byte K-major row
//     [ hi(27 taps) | 5 zeros | lo(27 taps) | 5 zeros ]      hi = bf16(v), lo = bf16(v - hi)
// (the image keeps 16 mantissa bits; the weights are rounded to bf16 like every other conv's of the bf16 mode) against the
// weight rows [ w | 0 | w | 0 ], so one K = 64 block (four MMAs) is the whole convolution of a tile.  The product is
// computed transposed (M = 128 output channels, N = 128 positions = four image rows), which makes TMEM lane = channel:
// the epilogue is the halo kernel's (64 contiguous bytes per pixel and warp, GroupNorm sums of the stored values in the
// thread).  The FFMA kernel this replaces (conv_small.cu: conv_in_rows_kernel) is instruction-issue bound at 69 us for
// batch 256; here the tensor-core time is negligible and the kernel is paced by its epilogue stores.
//
// CIN = 6 is the same problem met in training: the data gradient of the IDDPM output conv (models/ddpm.py:277) is a 3x3
// conv of the 6-channel fp32 NCHW image-space gradient into 128 channels.  Its 54 taps fill a K = 64 block on their own,
// so hi and lo are two operand tiles (eight MMAs per unit) against the one weight tile [ w(54) | 0 ].
//
// Warps: 0 = TMA producer (zero-padded fp32 patches: rows -1 / 32 and columns -1 / 32 are out of bounds = zero fill),
// 1 = MMA issuer / TMEM owner, 2..9 = epilogue, 10..17 = operand builders (thread = one position x 8 taps).
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct ConvInTcParams {
  CUtensorMap x;        // image NCHW fp32 as (w, h, c, n): box (40, 6, CIN, 1) at (-4, 4 t - 1, 0, image)
  const float* weight;  // fp32 [9 CIN][128], k = tap * CIN + ci (dmme_pack_conv_weight, DMME_CONV_GENERIC)
  const float* bias;
  __nv_bfloat16* out;   // NHWC
  long long* stats;
  int n, units;         // units = n * 8 (four image rows each)
};

constexpr int kInW = 32;
constexpr int kInCout = 128;
constexpr int kInPatchX0 = 4;                        // patch column of image column 0 (TMA: 16-byte aligned box start)
constexpr int kInPatchCols = 40;                     // image columns -4 .. 35
constexpr int kInPatchStages = 4;
constexpr int kInBTile = 128 * 128;                  // [128 positions][64 k] bf16
constexpr int kInBStages = 3;
template <int CIN> struct InGeom {
  static constexpr int kTaps = 9 * CIN;                        // 27: hi and lo halves of one tile; 54: a tile each
  static constexpr int kTiles = CIN == 3 ? 1 : 2;
  static constexpr int kGroups = CIN == 3 ? 4 : 8;             // 8-tap groups of a position
  static constexpr int kPatchFloats = CIN * 6 * kInPatchCols;
  static constexpr int kPatchSlot = (kPatchFloats * 4 + 1023) / 1024 * 1024;
  static constexpr int kSmem = kInCout * 128 + kInBStages * kTiles * kInBTile + kInPatchStages * kPatchSlot + 1024;
};
constexpr int kInEpiWarps = 8, kInBuildWarps = 8;
constexpr int kInThreads = (2 + kInEpiWarps + kInBuildWarps) * 32;

template <int CIN>
__global__ void __launch_bounds__(kInThreads, 1) conv_in_tc_kernel(const __grid_constant__ ConvInTcParams p) {
  using G = InGeom<CIN>;
  constexpr int kStage = G::kTiles * kInBTile;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t patch_full[kInPatchStages], patch_empty[kInPatchStages];
  __shared__ __align__(8) uint64_t b_ready[kInBStages], b_empty[kInBStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t w_ready;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* wbuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // [128 channels][64 k] bf16, SWIZZLE_128B
  uint8_t* bbuf = wbuf + kInCout * 128;
  uint8_t* pbuf = bbuf + kInBStages * kStage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kInPatchStages; ++s) { mbar_init(&patch_full[s], 1); mbar_init(&patch_empty[s], kInBuildWarps * 32); }
    for (int s = 0; s < kInBStages; ++s) { mbar_init(&b_ready[s], kInBuildWarps * 32); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kInEpiWarps * 32); }
    mbar_init(&w_ready, kInBuildWarps * 32);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.x);
  if (warp == 1) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();

  // a CTA walks a contiguous range of units: the tiles of an image stay together (one flush of the statistics per image)
  const int u_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.units / gridDim.x);
  const int u_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.units / gridDim.x);

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      pdl_wait();
      int it = 0;
      for (int u = u_begin; u < u_end; ++u, ++it) {
        const int s = it % kInPatchStages;
        mbar_wait(&patch_empty[s], ((it / kInPatchStages) & 1) ^ 1);
        mbar_expect_tx(&patch_full[s], G::kPatchFloats * 4);
        // the innermost start coordinate must be a multiple of 16 bytes: the box starts at image column -4
        tma_load_4d(pbuf + s * G::kPatchSlot, &p.x, &patch_full[s], -kInPatchX0, (u & 7) * 4 - 1, 0, u >> 3);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint64_t wdesc = umma_desc_sw128(smem_u32(wbuf));
      mbar_wait(&w_ready, 0);
      tc_fence_after();
      int it = 0;
      for (int u = u_begin; u < u_end; ++u, ++it) {
        const int stage = it & 1, bs = it % kInBStages;
        mbar_wait(&acc_empty[stage], ((it >> 1) & 1) ^ 1);
        mbar_wait(&b_ready[bs], (it / kInBStages) & 1);
        tc_fence_after();
        const uint32_t dtm = tmem_base + stage * 128;
#pragma unroll
        for (int tl = 0; tl < G::kTiles; ++tl) {  // CIN = 6: the hi tile, then the lo tile, against the same weight tile
          const uint64_t xdesc = umma_desc_sw128(smem_u32(bbuf + bs * kStage + tl * kInBTile));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(dtm, wdesc + 2 * k, xdesc + 2 * k, idesc, (tl | k) != 0 ? 1u : 0u);
        }
        umma_commit(&b_empty[bs]);
        umma_commit(&acc_full[stage]);
      }
    }
  } else if (warp >= 2 + kInEpiWarps) {
    // =========================== operand builders ===========================
    const int bt = threadIdx.x - (2 + kInEpiWarps) * 32;  // 0..255
    pdl_wait();
    // weight rows: CIN = 3 [ w(27) | 0 | w(27) | 0 ] (unit g and unit g + 4 hold taps [8 g, +8)), CIN = 6 [ w(54) | 0 ]
    for (int idx = bt; idx < kInCout * G::kGroups; idx += kInBuildWarps * 32) {
      const int co = idx / G::kGroups, g = idx % G::kGroups;
      float wv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = g * 8 + e;
        wv[e] = k < G::kTaps ? __ldg(p.weight + k * kInCout + co) : 0.f;
      }
      uint4 o;
      o.x = pack_bf16x2(wv[0], wv[1]); o.y = pack_bf16x2(wv[2], wv[3]);
      o.z = pack_bf16x2(wv[4], wv[5]); o.w = pack_bf16x2(wv[6], wv[7]);
      uint8_t* row = wbuf + co * 128;
      *reinterpret_cast<uint4*>(row + ((g ^ (co & 7)) << 4)) = o;
      if (CIN == 3) *reinterpret_cast<uint4*>(row + (((g + 4) ^ (co & 7)) << 4)) = o;
    }
    fence_proxy_async();
    mbar_arrive(&w_ready);
    // this thread's taps: k = 8 g + e -> (ci, dy, dx) -> offset inside the patch [3][6][40] of position (row 0, x = 0)
    const int g = bt % G::kGroups;
    int koff[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g * 8 + e;
      const int tap = k / CIN, ci = k - tap * CIN, dy = tap / 3, dx = tap - dy * 3;
      koff[e] = k < G::kTaps ? ci * (6 * kInPatchCols) + dy * kInPatchCols + dx : -1;
    }
    constexpr int kPosStep = kInBuildWarps * 32 / G::kGroups;  // positions covered by one pass of the builder threads
    int it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
      const int s = it % kInPatchStages, bs = it % kInBStages;
      mbar_wait(&patch_full[s], (it / kInPatchStages) & 1);
      mbar_wait(&b_empty[bs], ((it / kInBStages) & 1) ^ 1);
      const float* patch = reinterpret_cast<const float*>(pbuf + s * G::kPatchSlot);
      uint8_t* tile = bbuf + bs * kStage;
#pragma unroll
      for (int j = 0; j < 128 / kPosStep; ++j) {
        const int pos = bt / G::kGroups + kPosStep * j;  // position inside the unit: row pos / 32, column pos % 32
        const int base = (pos >> 5) * kInPatchCols + (pos & 31) + kInPatchX0 - 1;  // tap dx = 0 reads image column x - 1
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = koff[e] >= 0 ? patch[base + koff[e]] : 0.f;
        uint4 hi, lo;
        float r[8];
        hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
        hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
        unpack_bf16x2(hi.x, r[0], r[1]); unpack_bf16x2(hi.y, r[2], r[3]);
        unpack_bf16x2(hi.z, r[4], r[5]); unpack_bf16x2(hi.w, r[6], r[7]);
        lo.x = pack_bf16x2(v[0] - r[0], v[1] - r[1]); lo.y = pack_bf16x2(v[2] - r[2], v[3] - r[3]);
        lo.z = pack_bf16x2(v[4] - r[4], v[5] - r[5]); lo.w = pack_bf16x2(v[6] - r[6], v[7] - r[7]);
        uint8_t* row = tile + pos * 128;
        *reinterpret_cast<uint4*>(row + ((g ^ (pos & 7)) << 4)) = hi;
        if (CIN == 3) *reinterpret_cast<uint4*>(row + (((g + 4) ^ (pos & 7)) << 4)) = lo;
        else *reinterpret_cast<uint4*>(row + kInBTile + ((g ^ (pos & 7)) << 4)) = lo;
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the MMA's async-proxy reads
      mbar_arrive(&b_ready[bs]);
      mbar_arrive(&patch_empty[s]);
    }
  } else {
    // =========================== epilogue: thread = output channel ===========================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int ch = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    pdl_wait();
    const float bias_c = p.bias ? __ldg(p.bias + ch) : 0.f;
    float s1 = 0.f, s2 = 0.f;
    int cur_n = -1;
    auto flush_stats = [&]() {  // warp-uniform: per-image sums of this lane's channel -> micro-group atomics
      if (p.stats && cur_n >= 0) {
        float a1 = s1, a2 = s2;
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
        if ((lane & 3) == 0) {
          unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                   (static_cast<long long>(cur_n) * (kInCout >> 2) + (ch >> 2)) * 2;
          atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(a1 * kFix)));
          atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(a2 * kFix)));
        }
      }
      s1 = 0.f; s2 = 0.f;
    };
    int it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
      const int stage = it & 1;
      const int img = u >> 3;
      if (img != cur_n) {
        flush_stats();
        cur_n = img;
      }
      mbar_wait(&acc_full[stage], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int rr = 2 * half; rr < 2 * half + 2; ++rr) {
        const int yy = (u & 7) * 4 + rr;
        __nv_bfloat16* op = p.out + ((static_cast<long long>(img) * kInW + yy) * kInW) * kInCout + ch;
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + static_cast<uint32_t>(stage * 128 + rr * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const __nv_bfloat16 r = __float2bfloat16_rn(__uint_as_float(v[i]) + bias_c);
          op[i * kInCout] = r;
          const float rf = __bfloat162float(r);
          s1 += rf;
          s2 = fmaf(rf, rf, s2);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[stage]);
    }
    flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int g_in_tc_mode = 1;  // 0 = the input conv stays on the FFMA kernel, 1 = default, 2 = every batch size

// Below 64 images the FFMA kernel's shorter prologue wins (measured inside the step: 1.119 vs 1.127 ms at 32 images,
// 3.25 vs 3.20 ms at 256)
constexpr int kInTcMinBatch = 64;

bool conv_in_tc_supported(const dmme_conv_desc& d) {
  const bool plain = d.ksize == 3 && d.stride == 1 && !d.upsample && d.c1 == 0 && d.rc0 == 0 && d.rc1 == 0 && !d.temb &&
                     !d.addend && d.act_dtype == DMME_BF16 && d.in_layout == DMME_IN_NCHW_F32 && d.out_layout == DMME_OUT_NHWC;
  return g_in_tc_mode != 0 && plain && (d.c0 == 3 || d.c0 == 6) && d.cout == kInCout && d.w_in == kInW && d.h_in == kInW &&
         d.n >= (g_in_tc_mode == 2 ? 1 : kInTcMinBatch) && d.n <= (1 << 24);
}

template <int CIN>
static int launch_in_tc(const ConvInTcParams& p, int grid, cudaStream_t stream) {
  constexpr int smem = InGeom<CIN>::kSmem;
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_in_tc_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("conv_in_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  cudaError_t e = launch_pdl(conv_in_tc_kernel<CIN>, dim3(grid), dim3(kInThreads), smem, stream, p);
  return check_launch_err(e, "conv_in_tc_kernel");
}

int conv_in_tc_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_in_tc_supported(d), DMME_E_SHAPE, "conv_in_tc: unsupported shape/layout");
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_in_tc: null src0/weight/out");
  ConvInTcParams p;
  memset(&p, 0, sizeof(p));
  p.weight = static_cast<const float*>(d.weight);
  p.bias = d.bias;
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.stats = d.stats;
  p.n = d.n;
  p.units = d.n * (kInW / 4);
  {
    uint64_t dims[4] = {(uint64_t)kInW, (uint64_t)kInW, (uint64_t)d.c0, (uint64_t)d.n};
    uint64_t strides[3] = {(uint64_t)kInW * 4, (uint64_t)kInW * kInW * 4, (uint64_t)d.c0 * kInW * kInW * 4};
    uint32_t box[4] = {(uint32_t)kInPatchCols, 6u, (uint32_t)d.c0, 1u};
    int rc = encode_map_f32(&p.x, d.src0, 4, dims, strides, box);
    if (rc) return rc;
  }
  const int sm_count = device_sm_count();
  const int grid = p.units < sm_count ? p.units : sm_count;
  return d.c0 == 3 ? launch_in_tc<3>(p, grid, stream) : launch_in_tc<6>(p, grid, stream);
}

}  // namespace dmme

// A/B measurement switch: 0 = the input conv stays on the FFMA kernel, 1 = tcgen05 on 32x32 images from 64 images up
// (default), 2 = at every batch size
extern "C" void dmme_set_conv_in_tc_mode(int mode) { dmme::g_in_tc_mode = mode; }
