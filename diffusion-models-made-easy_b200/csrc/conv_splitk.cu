// Finishing pass of the split-K convolutions (conv_tc.cu, conv_tct_kernel<.., SPLIT>).
//
// The split-K GEMM leaves `split` fp32 partial tensors [pixel][cout].  One CTA owns one (image, 32-channel slab): it sums
// the slices, adds bias + timestep embedding + addend (ResBlock.forward models/ddpm.py:129,131), stores the raw bf16
// output and its GroupNorm micro-group statistics exactly like the conv epilogues do -- and, because it holds whole
// images, finishes the GroupNorm(+SiLU) of up to two consumers in the same launch (norm_act_drop_conv
// models/ddpm.py:25-35: the next conv of the chain, and the up-path ResBlock that later reads this tensor as the skip
// half of its concat).  At the 4x4 / 8x8 levels this replaces conv + gn_apply (+ gn_coeff) launches by GEMM + finish.
#include "common.cuh"

namespace dmme {

struct NormOut {
  __nv_bfloat16* out;
  const float* gamma; const float* beta;
  const float* scale; const float* shift;
  int ss_rows, ss_ld, cpg, silu;
  float eps;
};

struct SplitFinishParams {
  const float* partial;
  long long split_stride;
  int split;
  int n, hw, cout;
  const float* bias;
  const float* temb;
  int temb_rows, temb_ld;
  const __nv_bfloat16* addend;
  __nv_bfloat16* out;
  long long* stats;
  NormOut no[2];
};

constexpr int kFinishWarps = 16;
constexpr int kMaxSplit = 16;  // conv_tc.cu splitk_plan never splits further

// PPT: pixels per thread (hw <= 16 * PPT).  A warp covers the 32 channels of one pixel (128 contiguous bytes of every K
// slice); its pixels are px = warp, warp + 16, ...  All K slices of a pixel are requested before the first is used (one
// L2 round trip per pixel instead of one per slice pair: the pass is pure latency), and the finished values stay in
// registers for the GroupNorm pass.
template <int PPT>
__global__ void __launch_bounds__(kFinishWarps * 32) splitk_finish_kernel(const SplitFinishParams p) {
  __shared__ float red1[kFinishWarps][32], red2[kFinishWarps][32];
  __shared__ float ch1[32], ch2[32];
  const int slabs = p.cout >> 5;
  const int n = blockIdx.x / slabs;
  const int slab = blockIdx.x - n * slabs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = slab * 32 + lane;  // this thread's channel
  pdl_trigger();
  pdl_wait();  // the partial tiles come from the split-K GEMM launched just before

  float add = p.bias ? __ldg(p.bias + c) : 0.f;
  if (p.temb) add += __ldg(p.temb + static_cast<long long>(p.temb_rows == 1 ? 0 : n) * p.temb_ld + c);
  const long long img0 = static_cast<long long>(n) * p.hw;

  // ---- pass 1: sum the K slices (fixed order: deterministic), epilogue terms, raw bf16 store, sums of the STORED values
  float rf[PPT];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int px = warp + i * kFinishWarps;
    rf[i] = 0.f;
    if (px < p.hw) {
      const long long o = (img0 + px) * p.cout + c;
      const float* __restrict__ pp = p.partial + o;
      float a[kMaxSplit];
#pragma unroll
      for (int j = 0; j < kMaxSplit; ++j) a[j] = j < p.split ? __ldg(pp + j * p.split_stride) : 0.f;
      const float ad = p.addend ? __bfloat162float(p.addend[o]) : 0.f;
      float v = add;
#pragma unroll
      for (int j = 0; j < kMaxSplit; j += 4) v += (a[j] + a[j + 1]) + (a[j + 2] + a[j + 3]);
      v += ad;
      const __nv_bfloat16 r = __float2bfloat16_rn(v);
      p.out[o] = r;
      rf[i] = __bfloat162float(r);
      s1 += rf[i];
      s2 = fmaf(rf[i], rf[i], s2);
    }
  }
  red1[warp][lane] = s1;
  red2[warp][lane] = s2;
  __syncthreads();
  if (warp == 0) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < kFinishWarps; ++w) { t1 += red1[w][lane]; t2 += red2[w][lane]; }
    ch1[lane] = t1;
    ch2[lane] = t2;
    if (p.stats) {
      // [n][cout/4][2] fixed-point micro-group sums, the format every conv epilogue writes (dmme_conv_desc.stats)
      float m1 = t1, m2 = t2;
      m1 += __shfl_xor_sync(0xffffffffu, m1, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
      m1 += __shfl_xor_sync(0xffffffffu, m1, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
      if ((lane & 3) == 0) {
        const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
        unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                 (static_cast<long long>(n) * (p.cout >> 2) + (c >> 2)) * 2;
        atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(m1 * kFix)));
        atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(m2 * kFix)));
      }
    }
  }
  if (p.no[0].out == nullptr && p.no[1].out == nullptr) return;
  __syncthreads();

  // ---- pass 2: the consumers' GroupNorm(+SiLU) of the stored tensor ----
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const NormOut& q = p.no[k];
    if (q.out == nullptr) continue;  // uniform
    const int g0 = (lane / q.cpg) * q.cpg;
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < q.cpg; ++j) { t1 += ch1[g0 + j]; t2 += ch2[g0 + j]; }
    const float inv_cnt = 1.0f / (static_cast<float>(p.hw) * q.cpg);
    const float mean = t1 * inv_cnt;
    const float var = fmaxf(t2 * inv_cnt - mean * mean, 0.f);
    const float rs = rsqrtf(var + q.eps);
    const float ga = q.gamma ? __ldg(q.gamma + c) : 1.f, be = q.beta ? __ldg(q.beta + c) : 0.f;
    float aa = rs * ga, bb = be - mean * rs * ga;
    if (q.scale) {
      const long long r = static_cast<long long>(q.ss_rows == 1 ? 0 : n) * q.ss_ld;
      const float sc = 1.f + __ldg(q.scale + r + c), sh = __ldg(q.shift + r + c);
      aa *= sc;
      bb = bb * sc + sh;
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int px = warp + i * kFinishWarps;
      if (px < p.hw) {
        float y = fmaf(rf[i], aa, bb);
        if (q.silu) y = silu_f(y);
        q.out[(img0 + px) * p.cout + c] = __float2bfloat16_rn(y);
      }
    }
  }
}

// 4x4 maps (hw <= 16): one WARP owns one (image, 32-channel slab) -- lane = channel, every pixel of the image in that lane's
// registers -- so the per-channel sums need neither shared memory nor a block barrier, and a CTA of eight warps finishes
// eight slabs.  (The block-per-slab kernel above runs 512 threads for 512 outputs behind two __syncthreads: 2048 CTAs and
// 21 us per launch at batch 256 for 4 MB of partial tiles per slice.)  Same summation order as splitk_finish_kernel<1>
// (slices in fours, pixels 0..hw-1, group channels in ascending order): identical bits.
constexpr int kFinishSmallWarps = 8;
static int g_finish_small = 1;  // A/B (tests: bit-equality with the block-per-slab kernel): 0 = never take the warp-per-slab kernel
__global__ void __launch_bounds__(kFinishSmallWarps * 32) splitk_finish_small_kernel(const SplitFinishParams p) {
  constexpr int HW = 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slabs = p.cout >> 5;
  const int unit = blockIdx.x * kFinishSmallWarps + warp;
  pdl_trigger();
  pdl_wait();  // the partial tiles come from the split-K GEMM launched just before
  if (unit >= p.n * slabs) return;
  const int n = unit / slabs;
  const int c = (unit - n * slabs) * 32 + lane;
  float add = p.bias ? __ldg(p.bias + c) : 0.f;
  if (p.temb) add += __ldg(p.temb + static_cast<long long>(p.temb_rows == 1 ? 0 : n) * p.temb_ld + c);
  const long long img0 = static_cast<long long>(n) * p.hw;

  float rf[HW];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int b = 0; b < HW; b += 4) {
    // four pixels x every K slice requested before the first value is used
    float a[4][kMaxSplit], ad[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int px = b + i;
      const long long o = (img0 + px) * p.cout + c;
#pragma unroll
      for (int j = 0; j < kMaxSplit; ++j) a[i][j] = (px < p.hw && j < p.split) ? __ldg(p.partial + o + j * p.split_stride) : 0.f;
      ad[i] = (px < p.hw && p.addend) ? __bfloat162float(p.addend[o]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int px = b + i;
      rf[px] = 0.f;
      if (px < p.hw) {
        float v = add;
#pragma unroll
        for (int j = 0; j < kMaxSplit; j += 4) v += (a[i][j] + a[i][j + 1]) + (a[i][j + 2] + a[i][j + 3]);
        v += ad[i];
        const __nv_bfloat16 r = __float2bfloat16_rn(v);
        p.out[(img0 + px) * p.cout + c] = r;
        rf[px] = __bfloat162float(r);
        s1 += rf[px];
        s2 = fmaf(rf[px], rf[px], s2);
      }
    }
  }
  if (p.stats) {
    float m1 = s1, m2 = s2;
    m1 += __shfl_xor_sync(0xffffffffu, m1, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
    m1 += __shfl_xor_sync(0xffffffffu, m1, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
    if ((lane & 3) == 0) {
      const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
      unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                               (static_cast<long long>(n) * (p.cout >> 2) + (c >> 2)) * 2;
      atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(m1 * kFix)));
      atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(m2 * kFix)));
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const NormOut& q = p.no[k];
    if (q.out == nullptr) continue;  // uniform
    const int g0 = (lane / q.cpg) * q.cpg;
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < q.cpg; ++j) {  // ascending channel order, as the shared-memory version
      t1 += __shfl_sync(0xffffffffu, s1, g0 + j);
      t2 += __shfl_sync(0xffffffffu, s2, g0 + j);
    }
    const float inv_cnt = 1.0f / (static_cast<float>(p.hw) * q.cpg);
    const float mean = t1 * inv_cnt;
    const float var = fmaxf(t2 * inv_cnt - mean * mean, 0.f);
    const float rs = rsqrtf(var + q.eps);
    const float ga = q.gamma ? __ldg(q.gamma + c) : 1.f, be = q.beta ? __ldg(q.beta + c) : 0.f;
    float aa = rs * ga, bb = be - mean * rs * ga;
    if (q.scale) {
      const long long r = static_cast<long long>(q.ss_rows == 1 ? 0 : n) * q.ss_ld;
      const float sc = 1.f + __ldg(q.scale + r + c), sh = __ldg(q.shift + r + c);
      aa *= sc;
      bb = bb * sc + sh;
    }
#pragma unroll
    for (int px = 0; px < HW; ++px) {
      if (px < p.hw) {
        float y = fmaf(rf[px], aa, bb);
        if (q.silu) y = silu_f(y);
        q.out[(img0 + px) * p.cout + c] = __float2bfloat16_rn(y);
      }
    }
  }
}

int conv_splitk_finish(const dmme_conv_desc& d, int split, cudaStream_t stream) {
  SplitFinishParams p;
  memset(&p, 0, sizeof(p));
  const int ho = d.h_in / d.stride, wo = d.w_in / d.stride;
  p.partial = static_cast<const float*>(d.splitk_ws);
  p.split = split;
  p.n = d.n; p.hw = ho * wo; p.cout = d.cout;
  p.split_stride = static_cast<long long>(d.n) * p.hw * d.cout;
  p.bias = d.bias; p.temb = d.temb; p.temb_rows = d.temb_rows; p.temb_ld = d.temb_ld;
  p.addend = static_cast<const __nv_bfloat16*>(d.addend);
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.stats = d.stats;
  for (int k = 0; k < 2; ++k) {
    const dmme_out_norm& s = d.out_norm[k];
    if (s.out == nullptr) continue;
    DMME_REQUIRE(s.cpg >= 1 && s.cpg <= 32 && 32 % s.cpg == 0, DMME_E_SHAPE,
                 "conv split-K finish: out_norm channels per group must divide 32 (got %d)", s.cpg);
    DMME_REQUIRE(s.out != d.out, DMME_E_BADARG, "conv split-K finish: out_norm[%d].out aliases out", k);
    p.no[k].out = static_cast<__nv_bfloat16*>(s.out);
    p.no[k].gamma = s.gamma; p.no[k].beta = s.beta; p.no[k].scale = s.scale; p.no[k].shift = s.shift;
    p.no[k].ss_rows = s.ss_rows; p.no[k].ss_ld = s.ss_ld; p.no[k].cpg = s.cpg; p.no[k].silu = s.silu; p.no[k].eps = s.eps;
    DMME_REQUIRE(s.scale == nullptr || s.shift != nullptr, DMME_E_BADARG, "conv split-K finish: scale without shift");
  }
  DMME_REQUIRE(split >= 2 && split <= kMaxSplit && p.hw <= 256, DMME_E_SHAPE, "conv split-K finish: split %d / %d pixels per image", split, p.hw);
  const dim3 grid(d.n * (d.cout / 32)), block(kFinishWarps * 32);
  cudaError_t e;
  // measured (tools/sweep_step.py, DDPM step): 3.73 -> 3.68 ms at batch 256, but 2.37 -> 2.42 ms at batch 128 -- with fewer
  // than ~2000 slabs the warps' four sequential load batches are exposed; g_finish_small == 2 forces the kernel (tests)
  if (p.hw <= 16 && (g_finish_small == 2 || (g_finish_small == 1 && d.n * (d.cout / 32) >= 2048)))
    e = launch_pdl(splitk_finish_small_kernel, dim3(ceil_div(d.n * (d.cout / 32), kFinishSmallWarps)), dim3(kFinishSmallWarps * 32), 0, stream, p);
  else if (p.hw <= 16) e = launch_pdl(splitk_finish_kernel<1>, grid, block, 0, stream, p);
  else if (p.hw <= 64) e = launch_pdl(splitk_finish_kernel<4>, grid, block, 0, stream, p);
  else e = launch_pdl(splitk_finish_kernel<16>, grid, block, 0, stream, p);
  return check_launch_err(e, "splitk_finish_kernel");
}

}  // namespace dmme

// A/B switch: 0 = 4x4 maps take the block-per-slab finishing kernel too, 1 = warp-per-slab kernel (default)
extern "C" void dmme_set_splitk_finish_small(int mode) { dmme::g_finish_small = mode; }
