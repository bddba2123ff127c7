// The two image-space ends of the UNet, where one GEMM dimension is 3 (or 6) and tensor cores cannot help:
//   conv_in : x_t NCHW fp32 (<= 4 channels) --3x3--> NHWC bf16 (input_conv, models/ddpm.py:219-221); K = 27
//   conv_out: NHWC bf16 --3x3--> eps NCHW fp32 (<= 8 channels)  (output_conv[2], models/ddpm.py:277-279)
// Both are memory-bound (67 MB written / read at batch 256) and are laid out for coalesced 16-byte
// accesses on the NHWC side; weights (fp32, the generic [K][cout] packing) live in shared memory.
#include "common.cuh"

namespace dmme {

struct ConvSmallParams {
  const void* src; void* out;
  const float* weight; const float* bias;
  int n, h, w, cin, cout;
  long long* stats;  // conv_in only: GroupNorm micro-group sums of the stored output (see dmme_conv_desc.stats)
};

constexpr int kInPixPerBlock = 64;

// ---- input conv: one thread = one pixel x 8 output channels; a CTA owns 64 consecutive pixels of one image ----
template <int CIN>
__global__ void __launch_bounds__(256) conv_in_kernel(const ConvSmallParams p) {
  extern __shared__ float wsm[];  // [9*CIN][cout], bias [cout], reduction scratch
  constexpr int K = 9 * CIN;
  for (int i = threadIdx.x; i < K * p.cout; i += blockDim.x) wsm[i] = p.weight[i];
  float* bsm = wsm + K * p.cout;
  for (int i = threadIdx.x; i < p.cout; i += blockDim.x) bsm[i] = p.bias ? p.bias[i] : 0.f;
  float* red = bsm + p.cout;  // [pixels per iteration][cout/4][2]
  __syncthreads();
  const int cg = p.cout / 8;            // threads per pixel
  const int ppi = blockDim.x / cg;      // pixels per iteration
  const int g = threadIdx.x % cg, slot = threadIdx.x / cg;
  const float* x = static_cast<const float*>(p.src);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
  const long long plane = static_cast<long long>(p.h) * p.w;
  const long long npix = static_cast<long long>(p.n) * plane;
  const long long base = static_cast<long long>(blockIdx.x) * kInPixPerBlock;
  float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
  for (int it = 0; it < kInPixPerBlock; it += ppi) {
    const long long pix = base + it + slot;
    if (pix >= npix || slot >= ppi) continue;
    const int xx = static_cast<int>(pix % p.w);
    const int yy = static_cast<int>((pix / p.w) % p.h);
    const long long ni = pix / plane;
    float in[K];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int iy = yy + r - 1, ix = xx + q - 1;
        const bool ok = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
          in[(r * 3 + q) * CIN + ci] = ok ? __ldg(x + (ni * CIN + ci) * plane + static_cast<long long>(iy) * p.w + ix) : 0.f;
      }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bsm[g * 8 + j];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4* wr = reinterpret_cast<const float4*>(wsm + k * p.cout + g * 8);
      const float4 w0 = wr[0], w1 = wr[1];
      const float v = in[k];
      acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
      acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
      acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
      acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + pix * p.cout + g * 8) = o;
    if (p.stats) {
      unpack_bf16x2(o.x, acc[0], acc[1]); unpack_bf16x2(o.y, acc[2], acc[3]);
      unpack_bf16x2(o.z, acc[4], acc[5]); unpack_bf16x2(o.w, acc[6], acc[7]);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        s1[m] += (acc[4 * m] + acc[4 * m + 1]) + (acc[4 * m + 2] + acc[4 * m + 3]);
        s2[m] += (acc[4 * m] * acc[4 * m] + acc[4 * m + 1] * acc[4 * m + 1]) +
                 (acc[4 * m + 2] * acc[4 * m + 2] + acc[4 * m + 3] * acc[4 * m + 3]);
      }
    }
  }
  if (p.stats) {  // the 64 pixels of a CTA belong to one image (host checks plane % 64 == 0)
    const int mg = p.cout / 4;
    if (slot < ppi) {
      red[(slot * mg + g * 2) * 2 + 0] = s1[0]; red[(slot * mg + g * 2) * 2 + 1] = s2[0];
      red[(slot * mg + g * 2 + 1) * 2 + 0] = s1[1]; red[(slot * mg + g * 2 + 1) * 2 + 1] = s2[1];
    }
    __syncthreads();
    const long long ni = base / plane;
    if (base < npix) {
      for (int e = threadIdx.x; e < mg * 2; e += blockDim.x) {
        float t = 0.f;
        for (int sl = 0; sl < ppi; ++sl) t += red[sl * mg * 2 + e];
        atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + ni * mg * 2 + e,
                  static_cast<unsigned long long>(__float2ll_rn(t * static_cast<float>(1 << DMME_STATS_FRAC_BITS))));
      }
    }
  }
}

// ---- output conv: 8 lanes = one pixel, each lane owns every 8th 16-byte channel unit ---------------
template <int COUT>
__global__ void __launch_bounds__(256) conv_out_kernel(const ConvSmallParams p) {
  extern __shared__ float wsm[];  // [COUT][9*cin] (transposed for conflict-free float4 reads)
  const int K = 9 * p.cin;
  for (int i = threadIdx.x; i < K * COUT; i += blockDim.x) {
    const int k = i / COUT, co = i - k * COUT;
    wsm[co * K + k] = p.weight[i];
  }
  __syncthreads();
  const int lane8 = threadIdx.x & 7;
  const int units = p.cin / 64;  // 16-byte units per lane per tap
  const long long npix = static_cast<long long>(p.n) * p.h * p.w;
  const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.src);
  float* out = static_cast<float*>(p.out);
  const long long plane = static_cast<long long>(p.h) * p.w;
  const long long pstride = static_cast<long long>(gridDim.x) * (blockDim.x / 8);
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x / 8) + (threadIdx.x >> 3);
       pix < ((npix + 3) / 4) * 4; pix += pstride) {  // whole warps iterate together (shuffles below)
    const bool ok = pix < npix;
    const long long pp = ok ? pix : 0;
    const int xx = static_cast<int>(pp % p.w);
    const int yy = static_cast<int>((pp / p.w) % p.h);
    const long long ni = pp / plane;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    for (int u = 0; u < units; ++u) {
      const int ch = (u * 8 + lane8) * 8;
      uint4 v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {  // nine independent 16-byte loads in flight
        const int iy = yy + t / 3 - 1, ix = xx + t % 3 - 1;
        const bool in = ok && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        v[t] = in ? __ldg(reinterpret_cast<const uint4*>(src + ((ni * p.h + iy) * p.w + ix) * p.cin + ch))
                  : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float f[8];
        unpack_bf16x2(v[t].x, f[0], f[1]); unpack_bf16x2(v[t].y, f[2], f[3]);
        unpack_bf16x2(v[t].z, f[4], f[5]); unpack_bf16x2(v[t].w, f[6], f[7]);
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float4* wr = reinterpret_cast<const float4*>(wsm + c * K + t * p.cin + ch);
          const float4 w0 = wr[0], w1 = wr[1];
          acc[c] = fmaf(f[0], w0.x, acc[c]); acc[c] = fmaf(f[1], w0.y, acc[c]);
          acc[c] = fmaf(f[2], w0.z, acc[c]); acc[c] = fmaf(f[3], w0.w, acc[c]);
          acc[c] = fmaf(f[4], w1.x, acc[c]); acc[c] = fmaf(f[5], w1.y, acc[c]);
          acc[c] = fmaf(f[6], w1.z, acc[c]); acc[c] = fmaf(f[7], w1.w, acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
    }
    if (ok && lane8 < COUT) {
      float v0 = acc[0];
#pragma unroll
      for (int c = 1; c < COUT; ++c)
        if (lane8 == c) v0 = acc[c];
      v0 += p.bias ? p.bias[lane8] : 0.f;
      out[(ni * COUT + lane8) * plane + static_cast<long long>(yy) * p.w + xx] = v0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Faster variants for the production shapes (weights in registers, inputs broadcast from shared memory / slid through
// a register window).  Both are FFMA-bound by construction: 27 (resp. 9 * cin) FMAs per output element and nothing
// else in the inner loop -- 906 MFMA at batch 256 = 25 us at the SM's FFMA rate.
// ------------------------------------------------------------------------------------------------
// input conv: thread = one output-channel PAIR (54 weights in registers); a CTA owns RB image rows whose zero-padded
// fp32 patch sits in shared memory; the 32 lanes of a warp work on the same 4 pixels, so every shared-memory read is a
// broadcast and each warp-wide store writes 128 contiguous bytes.
template <int CIN>
__global__ void __launch_bounds__(256) conv_in_rows_kernel(const ConvSmallParams p, int rb) {
  extern __shared__ __align__(16) float sm_in[];  // [CIN][rb + 2][w + 8]: image column x at padded column x + 4
  __shared__ float red[8][64][2];   // per-warp statistics partials (warp, lane pair)
  constexpr int K = 9 * CIN;
  pdl_trigger();
  pdl_wait();
  const int pairs = p.cout >> 1;
  const int pair = threadIdx.x % pairs, slot = threadIdx.x / pairs, slots = blockDim.x / pairs;
  const int ws = p.w + 8;
  const int blocks_per_img = p.h / rb;
  const int n = blockIdx.x / blocks_per_img, y0 = (blockIdx.x - n * blocks_per_img) * rb;
  const float* x = static_cast<const float*>(p.src);
  const long long plane = static_cast<long long>(p.h) * p.w;
  for (int i = threadIdx.x; i < CIN * (rb + 2) * ws; i += blockDim.x) {
    const int col = i % ws, row = (i / ws) % (rb + 2), ci = i / (ws * (rb + 2));
    const int iy = y0 + row - 1, ix = col - 4;
    sm_in[i] = (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) ? __ldg(x + (static_cast<long long>(n) * CIN + ci) * plane + static_cast<long long>(iy) * p.w + ix) : 0.f;
  }
  float w0[K], w1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float2 wv = __ldg(reinterpret_cast<const float2*>(p.weight + static_cast<long long>(k) * p.cout) + pair);
    w0[k] = wv.x; w1[k] = wv.y;
  }
  const float b0 = p.bias ? p.bias[2 * pair] : 0.f, b1 = p.bias ? p.bias[2 * pair + 1] : 0.f;
  __syncthreads();
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
  float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
  const int quads = rb * (p.w >> 2);  // groups of 4 consecutive pixels in a row
  for (int q = slot; q < quads; q += slots) {
    const int ry = q / (p.w >> 2), x0 = (q - ry * (p.w >> 2)) << 2;
    float acc0[4] = {b0, b0, b0, b0}, acc1[4] = {b1, b1, b1, b1};
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* rowp = sm_in + (ci * (rb + 2) + ry + r) * ws + x0 + 3;  // image column x0 - 1
        const float4 mid = *reinterpret_cast<const float4*>(rowp + 1);
        const float v[6] = {rowp[0], mid.x, mid.y, mid.z, mid.w, rowp[5]};
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int k = (r * 3 + s) * CIN + ci;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc0[j] = fmaf(v[j + s], w0[k], acc0[j]);
            acc1[j] = fmaf(v[j + s], w1[k], acc1[j]);
          }
        }
      }
    const long long pix = (static_cast<long long>(n) * p.h + y0 + ry) * p.w + x0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t o = pack_bf16x2(acc0[j], acc1[j]);
      *reinterpret_cast<uint32_t*>(out + (pix + j) * p.cout + 2 * pair) = o;
      float lo, hi;
      unpack_bf16x2(o, lo, hi);
      s1a += lo; s2a = fmaf(lo, lo, s2a);
      s1b += hi; s2b = fmaf(hi, hi, s2b);
    }
  }
  if (p.stats) {  // micro-group of 4 channels = two adjacent pairs
    float s1 = s1a + s1b, s2 = s2a + s2b;
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
    float* mine = &red[0][0][0] + (static_cast<long long>(slot) * (pairs >> 1) + (pair >> 1)) * 2;  // [slot][micro-group][2]
    if ((pair & 1) == 0) { mine[0] = s1; mine[1] = s2; }
    __syncthreads();
    const int mg = pairs >> 1;
    for (int e = threadIdx.x; e < mg * 2; e += blockDim.x) {
      float t = 0.f;
      for (int sl = 0; sl < slots; ++sl) t += (&red[0][0][0])[static_cast<long long>(sl) * mg * 2 + e];
      atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + static_cast<long long>(n) * mg * 2 + e,
                static_cast<unsigned long long>(__float2ll_rn(t * static_cast<float>(1 << DMME_STATS_FRAC_BITS))));
    }
  }
}

// output conv, cin = 128: a warp produces 8 consecutive pixels of a row; lane l owns input channels [4l, 4l+4) with its
// 9 x 3 x 4 weights in registers; the 3 x 10 input window is read once (8-byte loads, 256 contiguous bytes per warp);
// the cross-lane sum is a halving butterfly (27 shuffles per 8 pixels).
__global__ void __launch_bounds__(256) conv_out128_kernel(const ConvSmallParams p, int co_off) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.src);
  float* out = static_cast<float*>(p.out);
  float wr[9][3][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
      for (int j = 0; j < 4; ++j) wr[t][co][j] = __ldg(p.weight + (static_cast<long long>(t) * 128 + 4 * lane + j) * p.cout + co_off + co);
  const int groups_per_row = p.w >> 3;
  const long long groups = static_cast<long long>(p.n) * p.h * groups_per_row;
  const long long plane = static_cast<long long>(p.h) * p.w;
  for (long long g = warp_global; g < groups; g += nwarps) {
    const int x0 = static_cast<int>(g % groups_per_row) << 3;
    const int y = static_cast<int>((g / groups_per_row) % p.h);
    const long long n = g / (static_cast<long long>(groups_per_row) * p.h);
    float acc[8][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = y + r - 1;
      if (iy < 0 || iy >= p.h) continue;  // warp-uniform
      const __nv_bfloat16* rowp = src + ((n * p.h + iy) * p.w) * 128 + 4 * lane;
      float v[10][4];
#pragma unroll
      for (int c = 0; c < 10; ++c) {
        const int ix = x0 + c - 1;
        uint2 u = make_uint2(0u, 0u);
        if (ix >= 0 && ix < p.w) u = __ldg(reinterpret_cast<const uint2*>(rowp + static_cast<long long>(ix) * 128));
        unpack_bf16x2(u.x, v[c][0], v[c][1]);
        unpack_bf16x2(u.y, v[c][2], v[c][3]);
      }
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int co = 0; co < 3; ++co)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][co] = fmaf(v[i + s][j], wr[r * 3 + s][co][j], acc[i][co]);
    }
    // halving butterfly: after the three exchanges lane bits (4,3,2) select the pixel, then a full sum over bits 1,0
    float h1[4][3], h2[2][3], h3[3];
    {
      const bool up = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const float keep = up ? acc[4 + i][co] : acc[i][co], give = up ? acc[i][co] : acc[4 + i][co];
          h1[i][co] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
        }
    }
    {
      const bool up = (lane & 8) != 0;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const float keep = up ? h1[2 + i][co] : h1[i][co], give = up ? h1[i][co] : h1[2 + i][co];
          h2[i][co] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
        }
    }
    {
      const bool up = (lane & 4) != 0;
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        const float keep = up ? h2[1][co] : h2[0][co], give = up ? h2[0][co] : h2[1][co];
        h3[co] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
      }
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      h3[co] += __shfl_xor_sync(0xffffffffu, h3[co], 2);
      h3[co] += __shfl_xor_sync(0xffffffffu, h3[co], 1);
    }
    if ((lane & 3) == 0) {
      const int px = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
      for (int co = 0; co < 3; ++co)
        out[(n * p.cout + co_off + co) * plane + static_cast<long long>(y) * p.w + x0 + px] = h3[co] + (p.bias ? p.bias[co_off + co] : 0.f);
    }
  }
}

static bool plain(const dmme_conv_desc& d) {
  return d.ksize == 3 && d.stride == 1 && !d.upsample && d.c1 == 0 && d.rc0 == 0 && d.rc1 == 0 && !d.temb &&
         !d.addend && d.act_dtype == DMME_BF16;
}
bool conv_in_supported(const dmme_conv_desc& d) {
  return plain(d) && d.in_layout == DMME_IN_NCHW_F32 && d.out_layout == DMME_OUT_NHWC && d.c0 == 3 &&
         d.cout % 8 == 0 && d.cout <= 256 && 256 % (d.cout / 8) == 0;
}
bool conv_out_supported(const dmme_conv_desc& d) {
  return plain(d) && d.in_layout == DMME_IN_NHWC && d.out_layout == DMME_OUT_NCHW_F32 && d.c0 % 64 == 0 &&
         d.c0 <= 512 && (d.cout == 3 || d.cout == 6);
}

int conv_small_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_small: null src0/weight/out");
  ConvSmallParams p;
  p.src = d.src0; p.out = d.out; p.weight = static_cast<const float*>(d.weight); p.bias = d.bias;
  p.n = d.n; p.h = d.h_in; p.w = d.w_in; p.cin = d.c0; p.cout = d.cout;
  p.stats = nullptr;
  if (conv_in_supported(d) && d.cout % 64 == 0 && 512 % d.cout == 0 && d.w_in % 4 == 0) {
    // production path: whole-row CTAs, weights in registers
    int rb = 256 / d.w_in;
    if (rb < 1) rb = 1;
    if (rb > d.h_in) rb = d.h_in;
    while (d.h_in % rb) --rb;
    const size_t smem = sizeof(float) * 3 * (rb + 2) * (d.w_in + 8);
    if (smem <= 48 * 1024 && (d.cout / 4) * (256 / (d.cout / 2)) <= 8 * 64) {
      p.stats = d.stats;  // a CTA never straddles two images
      return check_launch_err(launch_pdl(conv_in_rows_kernel<3>, dim3(d.n * (d.h_in / rb)), dim3(256), smem, stream, p, rb),
                              "conv_in_rows_kernel");
    }
  }
  if (conv_in_supported(d)) {
    const int cg = d.cout / 8, ppi = 256 / cg;
    const size_t smem = sizeof(float) * (static_cast<size_t>(27) * d.cout + d.cout + static_cast<size_t>(ppi) * (d.cout / 4) * 2);
    static DeviceOnce once_;
  bool& configured = once_.here();
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_in_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e != cudaSuccess) { set_error("conv_in: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
      configured = true;
    }
    const long long npix = static_cast<long long>(d.n) * d.h_in * d.w_in;
    if ((static_cast<long long>(d.h_in) * d.w_in) % kInPixPerBlock == 0) p.stats = d.stats;
    const int grid = static_cast<int>(ceil_div_ll(npix, kInPixPerBlock));
    conv_in_kernel<3><<<grid, 256, smem, stream>>>(p);
    return check_launch("conv_in_kernel");
  }
  DMME_REQUIRE(conv_out_supported(d), DMME_E_SHAPE, "conv_small: unsupported shape");
  if (d.c0 == 128 && d.w_in % 8 == 0) {
    const long long groups = static_cast<long long>(d.n) * d.h_in * (d.w_in / 8);
    long long blocks = ceil_div_ll(groups, 8 * 4);  // >= 4 pixel groups per warp: the 108 weight loads amortise
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    for (int co_off = 0; co_off < d.cout; co_off += 3) {
      int rc = check_launch_err(launch_pdl(conv_out128_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, p, co_off),
                                "conv_out128_kernel");
      if (rc) return rc;
    }
    return 0;
  }
  const size_t smem = sizeof(float) * static_cast<size_t>(9) * d.c0 * d.cout;
  const long long npix = static_cast<long long>(d.n) * d.h_in * d.w_in;
  const long long blocks = ceil_div_ll(npix, 32);
  const int grid = static_cast<int>(blocks < 148 * 4 ? blocks : 148 * 4);
  cudaError_t e;
  if (d.cout == 3) {
    e = cudaFuncSetAttribute(conv_out_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    if (e != cudaSuccess) { set_error("conv_out: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    conv_out_kernel<3><<<grid, 256, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(conv_out_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    if (e != cudaSuccess) { set_error("conv_out: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    conv_out_kernel<6><<<grid, 256, smem, stream>>>(p);
  }
  return check_launch("conv_out_kernel");
}

}  // namespace dmme
