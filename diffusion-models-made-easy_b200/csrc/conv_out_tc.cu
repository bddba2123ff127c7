// Output convolution (models/ddpm.py:277, GN -> SiLU -> Conv 128 -> 3, IDDPM -> 6) on tcgen05: 3x3 stride-1 conv with a
// handful of output channels, NHWC bf16 in, NCHW fp32 (image space) out.
//
// With cout = 3 the weight side of the GEMM is tiny and the activation side is everything, so the roles of conv_halo.cu
// are swapped: the halo tile of zero-padded pixel rows (same layout, same "a filter tap is a shift of the operand
// descriptor by d rows of 128 bytes" trick) is the MMA's M side -- 128 consecutive positions of the padded-row space =
// 128 TMEM lanes -- and the weights, padded to 16 output channels by the TMA's out-of-bounds zero fill, are the N side
// (M = 128, N = 16, K = 16 per instruction).  All 9 x cin/64 weight tiles (2 KB each) stay in shared memory for the
// whole kernel.  The accumulator has lane = position, column = output channel, so the epilogue is one 4/8-column
// tcgen05.ld per thread and cout coalesced fp32 stores (consecutive lanes = consecutive x of an image row).
// The FFMA kernel this replaces (conv_out128_kernel) needs 906 M FMAs at batch 256: 126 us against 25 us of FFMA issue.
//
// Work unit = RT whole padded rows (3 rows of 34 at 32x32, 7 of 18 at 16x16, 12 of 10 at 8x8), persistent CTAs, static
// round-robin.  Warps: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..5 = epilogue (one per TMEM lane quarter).
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "sampler.cuh"
#include "tmap.cuh"

namespace dmme {

static int g_out_tc_mode = 1;  // 0 = FFMA kernel, 1 = default, 2 = the per-tap kernel on every size (A/B)

struct ConvOutTcParams {
  CUtensorMap a;  // activations: box = one padded row [W+2 px][64 ch]
  CUtensorMap b;  // weights [cout][K] bf16, box [16][64] (rows >= cout zero-filled)
  int chunks;     // cin / 64
  int n, h, w, wp;
  int rt;          // padded rows per tile
  int total_rows;  // n * (h + 2)
  int units;
  int cout;
  const float* bias;
  float* out;  // [n][cout][h][w] fp32, or null when the sampler update below consumes eps in the epilogue
  // fused sampler update (dmme_sampler_epilogue): x_t <- x_{t-1} in place from the eps / v still in registers
  int samp_kind;
  float* x;
  const float* noise;
  const float* beta; const float* alpha; const float* alpha_bar;
  const int64_t* t_ptr; const int64_t* tau;
  int table_len, tau_len;
  unsigned long long seed, goff;  // goff: Philox group (4 elements) of x[0] inside the whole sample batch
};

constexpr int kOutSlot = 24 * 1024;  // >= (1 + (rt + 2) * (W + 2)) * 128 bytes
constexpr int kOutStages = 5;
constexpr int kOutWTile = 16 * 128;  // [16 cout][64 ch] bf16
constexpr int kOutThreads = 6 * 32;
constexpr int kOutN = 16;

template <int NCOL>  // accumulator columns read by the epilogue: 4 (cout <= 4) or 8
__global__ void __launch_bounds__(kOutThreads, 1) conv_out_tc_kernel(const __grid_constant__ ConvOutTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kOutStages], a_empty[kOutStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* abuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* wbuf = abuf + kOutStages * kOutSlot;
  const int row_bytes = p.wp * 128;
  const int nr = p.rt + 2;
  const int nwt = 9 * p.chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kOutStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    mbar_init(&w_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a);
    tma_prefetch_desc(&p.b);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();  // after the TMEM allocation (see common.cuh)

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      pdl_wait();
      // weight tile (tap, chunk) = columns [(tap * chunks + chunk) * 64, +64) of the tap-major packed [cout][K] matrix
      mbar_expect_tx(&w_full, static_cast<uint32_t>(nwt) * kOutWTile);
      for (int i = 0; i < nwt; ++i) tma_load_2d(wbuf + i * kOutWTile, &p.b, &w_full, i * 64, 0);
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int pr0 = u * p.rt - 1;  // first halo row (padded-row index, may be -1)
        for (int ck = 0; ck < p.chunks; ++ck, ++it) {
          const int s = it % kOutStages;
          mbar_wait(&a_empty[s], ((it / kOutStages) & 1) ^ 1);
          mbar_expect_tx(&a_full[s], nr * row_bytes);
          uint8_t* dst = abuf + s * kOutSlot + 128;  // 128 bytes of slack: tap (-1,-1) of position 0 reaches one row back
          for (int i = 0; i < nr; ++i) {
            const int pr = pr0 + i;
            int ni, yy;
            if (pr < 0) { ni = -1; yy = 0; }  // before the first image: whole row out of bounds -> zeros
            else { ni = pr / (p.h + 2); yy = pr - ni * (p.h + 2) - 1; }  // yy = -1 or h: padding row -> zeros
            tma_load_5d(dst + i * row_bytes, &p.a, &a_full[s], ck * 64, -1, 0, yy, ni);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kOutN);  // M = 128 positions, N = 16 (padded) output channels
      mbar_wait(&w_full, 0);
      int it = 0, u_it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++u_it) {
        const int stage = u_it & 1;
        mbar_wait(&acc_empty[stage], ((u_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dtm = tmem_base + stage * kOutN;
        for (int ck = 0; ck < p.chunks; ++ck, ++it) {
          const int s = it % kOutStages;
          mbar_wait(&a_full[s], (it / kOutStages) & 1);
          tc_fence_after();
          // position 0 of the tile = first pixel slot of the tile's first row = halo row 1
          const uint32_t x0_addr = smem_u32(abuf + s * kOutSlot) + 128u + row_bytes;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int d = (tap / 3 - 1) * p.wp + (tap % 3 - 1);
            const uint64_t xdesc = umma_desc_sw128(x0_addr + static_cast<uint32_t>(d * 128));  // shifted pixel rows: M side
            const uint64_t wdesc = umma_desc_sw128(smem_u32(wbuf + (tap * p.chunks + ck) * kOutWTile));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(dtm, xdesc + 2 * k, wdesc + 2 * k, idesc, (ck | tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[stage]);
      }
    }
  } else {
    // =========================== epilogue: thread = position ===========================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int pos = q * 32 + lane;
    const int rr = pos / p.wp;
    const int xx = pos - rr * p.wp - 1;
    const bool in_tile = rr < p.rt && xx >= 0 && xx < p.w;
    pdl_wait();
    float bias[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) bias[c] = (p.bias && c < p.cout) ? __ldg(p.bias + c) : 0.f;
    const long long plane = static_cast<long long>(p.h) * p.w;
    // per-step scalars of the fused sampler update: read once, after the grid dependency resolved (the step counter and
    // x_t were written by earlier kernels of the stream)
    DdpmScalars sd{};
    DdimScalars si{};
    IddpmScalars sv{};
    if (p.samp_kind == DMME_SAMPLER_DDPM) sd = ddpm_scalars(p.beta, p.alpha, p.alpha_bar, p.t_ptr, p.table_len);
    else if (p.samp_kind == DMME_SAMPLER_DDIM) si = ddim_scalars(p.alpha_bar, p.tau, p.t_ptr, p.table_len, p.tau_len);
    else if (p.samp_kind == DMME_SAMPLER_IDDPM) sv = iddpm_scalars(p.beta, p.alpha, p.alpha_bar, p.t_ptr, p.table_len);
    const int img_c = p.samp_kind == DMME_SAMPLER_IDDPM ? p.cout >> 1 : p.cout;  // channels of x_t
    int u_it = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++u_it) {
      const int stage = u_it & 1;
      const int pr = u * p.rt + rr;
      const int n = pr / (p.h + 2);
      const int yy = pr - n * (p.h + 2) - 1;
      const bool valid = in_tile && pr < p.total_rows && yy >= 0 && yy < p.h;
      mbar_wait(&acc_full[stage], (u_it >> 1) & 1);
      tc_fence_after();
      uint32_t v[NCOL];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(stage * kOutN);
      if constexpr (NCOL == 4) tmem_ld4(taddr, v);
      else tmem_ld8(taddr, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[stage]);
      if (valid) {
        if (p.out) {
          float* op = p.out + (static_cast<long long>(n) * p.cout * p.h + yy) * p.w + xx;
#pragma unroll
          for (int c = 0; c < NCOL; ++c)
            if (c < p.cout) op[c * plane] = __uint_as_float(v[c]) + bias[c];
        }
        if (p.samp_kind != DMME_SAMPLER_NONE) {
          // element e of x_t (NCHW): the noise of element e is lane e % 4 of Philox group e / 4, exactly what the
          // stand-alone kernels draw (a thread owns one pixel of every channel, so it computes a group per channel and
          // uses one of its four normals)
          const long long e0 = (static_cast<long long>(n) * img_c * p.h + yy) * p.w + xx;
#pragma unroll
          for (int c = 0; c < NCOL; ++c) {
            if (c < img_c) {
              const long long e = e0 + c * plane;
              const float eps = __uint_as_float(v[c]) + bias[c];
              const float xi = p.x[e];
              float z = 0.f;
              const bool need_z = p.samp_kind == DMME_SAMPLER_DDPM ? !sd.last : (p.samp_kind == DMME_SAMPLER_IDDPM ? !sv.last : false);
              if (need_z) {
                if (p.noise) {
                  z = p.noise[e];
                } else {
                  const unsigned long long t = static_cast<unsigned long long>(p.samp_kind == DMME_SAMPLER_DDPM ? sd.t : sv.t);
                  const float4 z4 = philox_normal4(p.seed, t, p.goff + static_cast<unsigned long long>(e >> 2));
                  const int j = static_cast<int>(e & 3);
                  z = j == 0 ? z4.x : (j == 1 ? z4.y : (j == 2 ? z4.z : z4.w));
                }
              } else if (p.noise && p.samp_kind != DMME_SAMPLER_DDIM) {
                z = p.noise[e];  // t == 1: drawn and discarded, like the reference; the value does not matter
              }
              float r;
              if (p.samp_kind == DMME_SAMPLER_DDPM) r = ddpm_update(xi, eps, z, sd);
              else if (p.samp_kind == DMME_SAMPLER_DDIM) r = ddim_update(xi, eps, si);
              else {
                // v = channel img_c + c of the network output; NCOL = 8 covers cout <= 8
                float vv = 0.f;
#pragma unroll
                for (int k = 0; k < NCOL; ++k)
                  if (k == img_c + c) vv = __uint_as_float(v[k]) + bias[k];
                r = iddpm_update(xi, eps, vv, z, sv);
              }
              p.x[e] = r;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// ------------------------------------------------------------------------------------------------
// 32 x 32 maps (the CIFAR-10 models' output conv): the horizontal taps move out of the K loop into the N dimension
// and the GroupNorm + SiLU in front of the conv (models/ddpm.py:277) is applied to the operand tile in shared memory.
//   Z[pos][3 co + dx] = sum_dy sum_c X[pos + (dy - 1) W][c] * w[co][dy][dx][c]    3 * cin / 16 MMAs per 128 positions
//   out[pos][co]      = sum_dx Z[pos + dx - 1][3 co + dx]                         two lane shuffles per channel
// A work unit is four whole image rows (128 positions = 128 TMEM lanes, lane = x within a warp's row) of one image; its
// operand is ONE TMA box of six unpadded rows per 64-channel chunk (rows -1 / 32 are out of bounds = zero fill = the
// conv's vertical padding), a vertical tap is a shift of the operand descriptor by 32 rows of 128 bytes (a whole number
// of swizzle atoms), and the horizontal padding is the shuffle's edge lanes.  Against the kernel above (one MMA group
// per tap over padded rows: 72 MMAs per 96 pixels, four epilogue warps that each draw a whole Philox group per output):
// 24 MMAs per 128 pixels, eight epilogue warps, one Philox group per lane.
// Warps: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..9 = epilogue (two per TMEM lane quarter, alternating units),
// 10..17 = GroupNorm + SiLU of the operand tile (thread = one 16-byte unit column x every 32nd position).
struct ConvOutDxParams {
  CUtensorMap a;  // activations NHWC as (c, w, h, n): box (64, 32, 6, 1)
  CUtensorMap b;  // weights [cout][dy][dx][cin] bf16 as (c, dx, co, dy): box (64, 3, 8, 1) = rows 3 co + dx
  int chunks;     // cin / 64
  int n, units;   // units = n * 8
  int cout, ncols;  // ncols = the MMA's N: 16 (cout <= 4) or 32
  const float* bias;
  float* out;
  const float2* gn_ab;  // [n][cin] (a, b) of the fused GroupNorm of the input, or null
  int gn_silu;
  int samp_kind;
  float* x;
  const float* noise;
  const float* beta; const float* alpha; const float* alpha_bar;
  const int64_t* t_ptr; const int64_t* tau;
  int table_len, tau_len;
  unsigned long long seed, goff;
};

constexpr int kDxW = 32;
constexpr int kDxSlot = 6 * kDxW * 128;  // six rows of 32 positions x 64 channels
constexpr int kDxStages = 6;
constexpr int kDxWTile = 32 * 128;       // 24 rows written by TMA (8 co x 3 dx), padded to the MMA's N = 32
constexpr int kDxEpiWarps = 8, kDxXfWarps = 8;
constexpr int kDxThreads = (2 + kDxEpiWarps + kDxXfWarps) * 32;

template <int NLD>  // accumulator columns the epilogue reads: 16 (cout <= 4) or 32 (cout <= 8)
__global__ void __launch_bounds__(kDxThreads, 1) conv_out_dx_kernel(const __grid_constant__ ConvOutDxParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kDxStages], a_ready[kDxStages], a_empty[kDxStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_slot;
  constexpr int kCo = NLD == 16 ? 4 : 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* abuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* wbuf = abuf + kDxStages * kDxSlot;
  const bool fused_gn = p.gn_ab != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDxStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_ready[s], kDxXfWarps * 32);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    mbar_init(&w_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a);
    tma_prefetch_desc(&p.b);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      pdl_wait();
      mbar_expect_tx(&w_full, static_cast<uint32_t>(3 * p.chunks) * (24 * 128));
      for (int dy = 0; dy < 3; ++dy)
        for (int ck = 0; ck < p.chunks; ++ck) tma_load_4d(wbuf + (dy * p.chunks + ck) * kDxWTile, &p.b, &w_full, ck * 64, 0, 0, dy);
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int img = u >> 3, y0 = (u & 7) * 4 - 1;
        for (int ck = 0; ck < p.chunks; ++ck, ++it) {
          const int s = it % kDxStages;
          mbar_wait(&a_empty[s], ((it / kDxStages) & 1) ^ 1);
          mbar_expect_tx(&a_full[s], kDxSlot);
          tma_load_4d(abuf + s * kDxSlot, &p.a, &a_full[s], ck * 64, 0, y0, img);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.ncols);
      mbar_wait(&w_full, 0);
      int it = 0, u_it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++u_it) {
        const int stage = u_it & 1;
        mbar_wait(&acc_empty[stage], ((u_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dtm = tmem_base + stage * 32;
        for (int ck = 0; ck < p.chunks; ++ck, ++it) {
          const int s = it % kDxStages;
          mbar_wait(fused_gn ? &a_ready[s] : &a_full[s], (it / kDxStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(abuf + s * kDxSlot);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t xdesc = umma_desc_sw128(a_addr + static_cast<uint32_t>(dy * kDxW * 128));  // positions + (dy - 1) W
            const uint64_t wdesc = umma_desc_sw128(smem_u32(wbuf + (dy * p.chunks + ck) * kDxWTile));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(dtm, xdesc + 2 * k, wdesc + 2 * k, idesc, (ck | dy | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[stage]);
      }
    }
  } else if (warp >= 2 + kDxEpiWarps) {
    // =========================== GroupNorm + SiLU of the operand tile ===========================
    // Same coefficient pairs, fma and SiLU as gn_apply_kernel (bit-identical operand).  Rows outside the image were
    // zero-filled by TMA and stay zero: the reference pads after the activation.
    if (fused_gn) {
      const int xt = threadIdx.x - (2 + kDxEpiWarps) * 32;
      const int pu = xt & 7;               // physical 16-byte unit inside the 128-byte row
      const int r0 = xt >> 3;              // position r0 of every slot row
      const int ul = pu ^ (r0 & 7);        // SWIZZLE_128B: logical unit = physical unit ^ (position & 7); slots are 1024-aligned
      const uint32_t off0 = static_cast<uint32_t>(r0) * 128u + static_cast<uint32_t>(pu) * 16u;
      pdl_wait();
      float4 cur[4], nxt[4];
      auto load_coeff = [&](float4 (&dst)[4], int u, int ck) {
        const float4* g = reinterpret_cast<const float4*>(p.gn_ab + static_cast<long long>(u >> 3) * (p.chunks * 64) + ck * 64 + ul * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = __ldg(g + j);
      };
      if (static_cast<int>(blockIdx.x) < p.units) load_coeff(cur, blockIdx.x, 0);
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int y0 = (u & 7) * 4 - 1;
        for (int ck = 0; ck < p.chunks; ++ck, ++it) {
          const int s = it % kDxStages;
          // the next stage's coefficients travel while this one is transformed
          const int un = ck + 1 < p.chunks ? u : u + static_cast<int>(gridDim.x);
          const int ckn = ck + 1 < p.chunks ? ck + 1 : 0;
          if (un < p.units) load_coeff(nxt, un, ckn);
          mbar_wait(&a_full[s], (it / kDxStages) & 1);
          uint8_t* tile = abuf + s * kDxSlot + off0;
          uint4 v[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int yy = y0 + j;
            if (yy >= 0 && yy < kDxW) v[j] = *reinterpret_cast<const uint4*>(tile + j * (kDxW * 128));
          }
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int yy = y0 + j;
            if (yy < 0 || yy >= kDxW) continue;
            float f[8];
            unpack_bf16x2(v[j].x, f[0], f[1]); unpack_bf16x2(v[j].y, f[2], f[3]);
            unpack_bf16x2(v[j].z, f[4], f[5]); unpack_bf16x2(v[j].w, f[6], f[7]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              f[2 * k] = fmaf(f[2 * k], cur[k].x, cur[k].y);
              f[2 * k + 1] = fmaf(f[2 * k + 1], cur[k].z, cur[k].w);
            }
            if (p.gn_silu) {
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] = silu_f(f[k]);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(tile + j * (kDxW * 128)) = o;
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the MMA's async-proxy reads
          mbar_arrive(&a_ready[s]);
#pragma unroll
          for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
        }
      }
    }
  } else {
    // =========================== epilogue: thread = pixel (warp = image row, lane = x) ===========================
    const int q = warp & 3;                // TMEM lane quarter = row of the unit
    const int half = (warp - 2) >> 2;      // units with (u_it & 1) == half, accumulator stage = half
    pdl_wait();
    float bias[kCo];
#pragma unroll
    for (int c = 0; c < kCo; ++c) bias[c] = (p.bias && c < p.cout) ? __ldg(p.bias + c) : 0.f;
    constexpr long long plane = kDxW * kDxW;
    DdpmScalars sd{};
    DdimScalars si{};
    IddpmScalars sv{};
    if (p.samp_kind == DMME_SAMPLER_DDPM) sd = ddpm_scalars(p.beta, p.alpha, p.alpha_bar, p.t_ptr, p.table_len);
    else if (p.samp_kind == DMME_SAMPLER_DDIM) si = ddim_scalars(p.alpha_bar, p.tau, p.t_ptr, p.table_len, p.tau_len);
    else if (p.samp_kind == DMME_SAMPLER_IDDPM) sv = iddpm_scalars(p.beta, p.alpha, p.alpha_bar, p.t_ptr, p.table_len);
    const int img_c = p.samp_kind == DMME_SAMPLER_IDDPM ? p.cout >> 1 : p.cout;  // channels of x_t (<= 4)
    const bool need_z = p.samp_kind == DMME_SAMPLER_DDPM ? !sd.last : (p.samp_kind == DMME_SAMPLER_IDDPM ? !sv.last : false);
    int u_it = half;
    for (int u = blockIdx.x + half * gridDim.x; u < p.units; u += 2 * gridDim.x, u_it += 2) {
      const int img = u >> 3, yy = (u & 7) * 4 + q;
      // x_t is fetched before the accumulator is awaited: the step's traffic has pushed it out of L2 and the DRAM round
      // trip would otherwise sit between the accumulator and the store
      const long long e0 = (static_cast<long long>(img) * img_c * kDxW + yy) * kDxW + lane;
      float xin[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.samp_kind != DMME_SAMPLER_NONE) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < img_c) xin[c] = p.x[e0 + c * plane];
      }
      mbar_wait(&acc_full[half], (u_it >> 1) & 1);
      tc_fence_after();
      uint32_t v[NLD];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(half * 32);
      if constexpr (NLD == 16) tmem_ld16(taddr, v);
      else tmem_ld32(taddr, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[half]);
      float o[kCo];
#pragma unroll
      for (int c = 0; c < kCo; ++c) {
        o[c] = 0.f;
        if (c < p.cout) {  // warp-uniform
          float fl = __shfl_up_sync(0xffffffffu, __uint_as_float(v[3 * c]), 1);        // dx = 0 term of the pixel to the left
          float fr = __shfl_down_sync(0xffffffffu, __uint_as_float(v[3 * c + 2]), 1);  // dx = 2 term of the pixel to the right
          if (lane == 0) fl = 0.f;         // horizontal zero padding
          if (lane == kDxW - 1) fr = 0.f;
          o[c] = ((__uint_as_float(v[3 * c + 1]) + fl) + fr) + bias[c];
        }
      }
      if (p.out) {
        float* op = p.out + (static_cast<long long>(img) * p.cout * kDxW + yy) * kDxW + lane;
#pragma unroll
        for (int c = 0; c < kCo; ++c)
          if (c < p.cout) op[c * plane] = o[c];
      }
      if (p.samp_kind != DMME_SAMPLER_NONE) {
        // element e of x_t (NCHW); its noise is lane e % 4 of Philox group e / 4, exactly what the stand-alone kernels
        // draw.  The four lanes of a quad share the groups of their row: lane j draws the group of channel j once and
        // the normals are handed round with shuffles.
        float zc[4] = {0.f, 0.f, 0.f, 0.f};
        if (need_z && !p.noise) {  // warp-uniform
          const int j = lane & 3;
          float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < img_c) {
            const unsigned long long t = static_cast<unsigned long long>(p.samp_kind == DMME_SAMPLER_DDPM ? sd.t : sv.t);
            const long long ej = e0 - j + j * plane;  // first element of the quad in channel j
            z4 = philox_normal4(p.seed, t, p.goff + static_cast<unsigned long long>(ej >> 2));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < img_c) {
              const int src = (lane & ~3) + c;
              const float a0 = __shfl_sync(0xffffffffu, z4.x, src), a1 = __shfl_sync(0xffffffffu, z4.y, src);
              const float a2 = __shfl_sync(0xffffffffu, z4.z, src), a3 = __shfl_sync(0xffffffffu, z4.w, src);
              zc[c] = j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < img_c) {
            const long long e = e0 + c * plane;
            const float xi = xin[c];
            float z = zc[c];
            if (p.noise && p.samp_kind != DMME_SAMPLER_DDIM) z = p.noise[e];  // t == 1: drawn and discarded, like the reference
            float r;
            if (p.samp_kind == DMME_SAMPLER_DDPM) r = ddpm_update(xi, o[c], z, sd);
            else if (p.samp_kind == DMME_SAMPLER_DDIM) r = ddim_update(xi, o[c], si);
            else {
              float vv = 0.f;
#pragma unroll
              for (int k = 0; k < kCo; ++k)
                if (k == img_c + c) vv = o[k];
              r = iddpm_update(xi, o[c], vv, z, sv);
            }
            p.x[e] = r;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}


bool conv_out_tc_supported(const dmme_conv_desc& d) {
  if (g_out_tc_mode == 0) return false;
  if (d.act_dtype != DMME_BF16 || d.in_layout != DMME_IN_NHWC || d.out_layout != DMME_OUT_NCHW_F32) return false;
  if (d.ksize != 3 || d.stride != 1 || d.upsample || d.c1 || d.rc0 || d.rc1 || d.temb || d.addend) return false;
  if (d.sampler && d.sampler->kind == DMME_SAMPLER_IDDPM && (d.cout & 1)) return false;
  if (d.c0 <= 0 || d.c0 % 64 || d.c0 > 256) return false;  // 9 * cin/64 resident weight tiles of 2 KB
  if (d.cout < 1 || d.cout > 8) return false;
  if (d.w_in != 8 && d.w_in != 16 && d.w_in != 32) return false;
  if (d.h_in != d.w_in) return false;
  if (static_cast<long long>(d.n) * (d.h_in + 2) > (1 << 24)) return false;
  return true;
}

// the row-tile kernel with the horizontal taps in N (and the fused GroupNorm of the input): 32 x 32 maps
bool conv_out_dx_supported(const dmme_conv_desc& d) {
  return g_out_tc_mode == 1 && conv_out_tc_supported(d) && d.w_in == kDxW && d.h_in == kDxW;
}

template <int NLD>
static int launch_out_dx(const ConvOutDxParams& p, int smem, int grid, cudaStream_t stream) {
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_out_dx_kernel<NLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_out_dx: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  cudaError_t e = launch_pdl(conv_out_dx_kernel<NLD>, dim3(grid), dim3(kDxThreads), smem, stream, p);
  return check_launch_err(e, "conv_out_dx_kernel");
}

static int conv_out_dx_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  ConvOutDxParams p;
  memset(&p, 0, sizeof(p));
  p.chunks = d.c0 / 64;
  p.n = d.n;
  p.units = d.n * (kDxW / 4);
  p.cout = d.cout;
  p.ncols = d.cout <= 4 ? 16 : 32;
  p.bias = d.bias;
  p.out = static_cast<float*>(d.out);
  p.gn_ab = reinterpret_cast<const float2*>(d.gn_ab);
  p.gn_silu = d.gn_silu;
  if (d.sampler && d.sampler->kind != DMME_SAMPLER_NONE) {
    const dmme_sampler_epilogue& s = *d.sampler;
    p.samp_kind = s.kind; p.x = s.x; p.noise = s.noise;
    p.beta = s.beta; p.alpha = s.alpha; p.alpha_bar = s.alpha_bar; p.t_ptr = s.t_ptr; p.tau = s.tau;
    p.table_len = s.table_len; p.tau_len = s.tau_len; p.seed = s.seed; p.goff = s.noise_offset / 4;
  }
  int rc;
  {
    const uint64_t row = (uint64_t)d.c0 * 2;
    uint64_t dims[4] = {(uint64_t)d.c0, (uint64_t)kDxW, (uint64_t)kDxW, (uint64_t)d.n};
    uint64_t strides[3] = {row, row * kDxW, row * kDxW * kDxW};
    uint32_t box[4] = {64u, (uint32_t)kDxW, 6u, 1u};
    if ((rc = encode_map(&p.a, d.src0, 4, dims, strides, box))) return rc;
  }
  {
    // packed [cout][K], K = (dy * 3 + dx) * cin + c, viewed as (c, dx, co, dy): a box is the rows 3 co + dx of one dy
    const uint64_t cin2 = (uint64_t)d.c0 * 2;
    uint64_t dims[4] = {(uint64_t)d.c0, 3, (uint64_t)d.cout, 3};
    uint64_t strides[3] = {cin2, 9 * cin2, 3 * cin2};
    uint32_t box[4] = {64u, 3u, 8u, 1u};
    if ((rc = encode_map(&p.b, d.weight, 4, dims, strides, box))) return rc;
  }
  const int sm_count = device_sm_count();
  const int smem = kDxStages * kDxSlot + 3 * p.chunks * kDxWTile + 1024;
  const int grid = p.units < sm_count ? p.units : sm_count;
  return p.ncols == 16 ? launch_out_dx<16>(p, smem, grid, stream) : launch_out_dx<32>(p, smem, grid, stream);
}

template <int NCOL>
static int launch_out_tc(const ConvOutTcParams& p, int smem, int grid, cudaStream_t stream) {
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_out_tc_kernel<NCOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_out_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  cudaError_t e = launch_pdl(conv_out_tc_kernel<NCOL>, dim3(grid), dim3(kOutThreads), smem, stream, p);
  return check_launch_err(e, "conv_out_tc_kernel");
}

int conv_out_tc_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_out_tc_supported(d), DMME_E_SHAPE, "conv_out_tc: unsupported shape/layout");
  DMME_REQUIRE(d.src0 && d.weight && (d.out || d.sampler), DMME_E_BADARG, "conv_out_tc: null src0/weight/out");
  ConvOutTcParams p;
  memset(&p, 0, sizeof(p));
  p.chunks = d.c0 / 64;
  p.n = d.n; p.h = d.h_in; p.w = d.w_in; p.wp = d.w_in + 2;
  p.rt = 128 / p.wp;
  p.total_rows = d.n * (d.h_in + 2);
  p.units = (p.total_rows + p.rt - 1) / p.rt;
  p.cout = d.cout;
  p.bias = d.bias;
  p.out = static_cast<float*>(d.out);
  if (d.sampler && d.sampler->kind != DMME_SAMPLER_NONE) {
    const dmme_sampler_epilogue& s = *d.sampler;
    DMME_REQUIRE(s.kind >= DMME_SAMPLER_DDPM && s.kind <= DMME_SAMPLER_IDDPM, DMME_E_BADARG, "conv_out_tc: unknown sampler kind %d", s.kind);
    DMME_REQUIRE(s.x && s.alpha_bar && s.t_ptr && s.table_len > 0, DMME_E_BADARG, "conv_out_tc: sampler epilogue needs x, alpha_bar, t_ptr");
    DMME_REQUIRE(s.kind == DMME_SAMPLER_DDIM ? (s.tau && s.tau_len > 0) : (s.beta && s.alpha), DMME_E_BADARG,
                 "conv_out_tc: sampler epilogue tables missing");
    DMME_REQUIRE(s.noise_offset % 4 == 0 && (static_cast<long long>(d.h_in) * d.w_in) % 4 == 0, DMME_E_BADARG,
                 "conv_out_tc: noise_offset / image size must be multiples of 4");
    p.samp_kind = s.kind; p.x = s.x; p.noise = s.noise;
    p.beta = s.beta; p.alpha = s.alpha; p.alpha_bar = s.alpha_bar; p.t_ptr = s.t_ptr; p.tau = s.tau;
    p.table_len = s.table_len; p.tau_len = s.tau_len; p.seed = s.seed; p.goff = s.noise_offset / 4;
  }
  if (conv_out_dx_supported(d)) return conv_out_dx_forward(d, stream);
  DMME_REQUIRE(d.gn_ab == nullptr, DMME_E_UNSUPPORTED, "conv_out_tc: a fused GroupNorm of the input needs a 32x32 map");
  // the halo tile is (rt + 2) padded rows behind one slack row.  The MMA's 128 positions reach up to 2 * (W + 2) + 130
  // rows from the slot start: positions past the tile's rt rows are junk lanes that are never stored, and what they
  // read (the next slot or the resident weights) lies inside this CTA's shared memory
  DMME_REQUIRE((1 + (p.rt + 2) * p.wp) * 128 <= kOutSlot &&
                   (2 * p.wp + 130) * 128 <= kOutSlot + 9 * p.chunks * kOutWTile,
               DMME_E_SHAPE, "conv_out_tc: halo tile does not fit its shared-memory slot");
  int rc;
  {
    uint64_t dims[5] = {(uint64_t)d.c0, (uint64_t)d.w_in, 1, (uint64_t)d.h_in, (uint64_t)d.n};
    uint64_t strides[4] = {(uint64_t)d.c0 * 2, (uint64_t)d.w_in * d.c0 * 2, (uint64_t)d.w_in * d.c0 * 2,
                           (uint64_t)d.h_in * d.w_in * d.c0 * 2};
    uint32_t box[5] = {64u, (uint32_t)p.wp, 1u, 1u, 1u};
    if ((rc = encode_map(&p.a, d.src0, 5, dims, strides, box))) return rc;
  }
  {
    const uint64_t ktot = 9ull * d.c0;
    uint64_t dims[2] = {ktot, (uint64_t)d.cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64u, 16u};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
  }
  const int sm_count = device_sm_count();
  const int smem = kOutStages * kOutSlot + 9 * p.chunks * kOutWTile + 1024;
  const int grid = p.units < sm_count ? p.units : sm_count;
  return d.cout <= 4 ? launch_out_tc<4>(p, smem, grid, stream) : launch_out_tc<8>(p, smem, grid, stream);
}

}  // namespace dmme

// A/B measurement switch: 0 = the output conv stays on the FFMA kernel, 1 = default, 2 = tcgen05 without the 32x32 row-tile kernel
extern "C" void dmme_set_conv_out_tc_mode(int mode) { dmme::g_out_tc_mode = mode; }
