// GroupNorm (+ scale/shift) (+ SiLU) (+ channel-dropout mask) over NHWC activations.
//
// Fast path (bf16, C % 32 == 0, channels-per-group in {1,2,4,8,16,32}, HW <= 1024): one CTA owns one
// (image, 32-channel slab) = 64 bytes per pixel.  The slab is read ONCE with 16-byte vector loads into
// registers, mean and variance are computed in two register passes (fp32, same biased variance as
// torch.native_group_norm), and the normalised/activated slab is written once: algorithmic traffic
// = 2 B read + 2 B written per element.  The skip concat is two source pointers.
// Generic path: any C / groups / HW / storage type; one CTA per (image, group), three global passes.
#include "common.cuh"

namespace dmme {

struct GnParams {
  const void* src0; const void* src1; int c0, c1;
  int n, hw, groups; float eps;
  const float* gamma; const float* beta;
  const float* scale; const float* shift; int ss_rows, ss_ld;
  const float* mask;  // [n][C] or null
  int silu;
  void* out;
};

// --------------------------------------------------------------------------------------------
// fast path
// --------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256, 2) gn_slab_kernel(const GnParams p) {
  __shared__ float red[8][32];
  __shared__ float chan_stat[32];
  __shared__ float gmean[32], grstd[32];

  const int C = p.c0 + p.c1;
  const int slabs = C / 32;
  const int n = blockIdx.x / slabs;
  const int slab = blockIdx.x - n * slabs;
  const int cbase = slab * 32;
  const int cpg = C / p.groups;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int chunk = tid & 3;  // 8-channel vector inside the slab
  const int nvec = p.hw * 4;  // 16-byte vectors in the slab

  const __nv_bfloat16* src;
  int csrc, coff;
  if (cbase < p.c0) { src = static_cast<const __nv_bfloat16*>(p.src0); csrc = p.c0; coff = cbase; }
  else { src = static_cast<const __nv_bfloat16*>(p.src1); csrc = p.c1; coff = cbase - p.c0; }
  const __nv_bfloat16* base = src + static_cast<long long>(n) * p.hw * csrc + coff + chunk * 8;

  uint4 v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) v[i] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(vi >> 2) * csrc));
    else v[i] = make_uint4(0, 0, 0, 0);
  }

  // ---- pass 1: per-channel sums -> group means ----
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    float lo, hi;
    unpack_bf16x2(v[i].x, lo, hi); s[0] += lo; s[1] += hi;
    unpack_bf16x2(v[i].y, lo, hi); s[2] += lo; s[3] += hi;
    unpack_bf16x2(v[i].z, lo, hi); s[4] += lo; s[5] += hi;
    unpack_bf16x2(v[i].w, lo, hi); s[6] += lo; s[7] += hi;
  }
  // lanes with equal (lane & 3) hold the same channels
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 4);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
  }
  if (lane < 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = s[j];
  }
  __syncthreads();
  if (tid < 32) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w][tid];
    chan_stat[tid] = t;
  }
  __syncthreads();
  const float inv_cnt = 1.0f / (static_cast<float>(p.hw) * cpg);
  if (tid < 32) {
    // group of channel tid inside the slab: channels [g0, g0 + cpg)
    const int g0 = (tid / cpg) * cpg;
    float t = 0.f;
    for (int j = 0; j < cpg; ++j) t += chan_stat[g0 + j];
    gmean[tid] = t * inv_cnt;
  }
  __syncthreads();
  float mean[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mean[j] = gmean[chunk * 8 + j];

  // ---- pass 2: centred second moment ----
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) {
      float lo, hi, d;
      unpack_bf16x2(v[i].x, lo, hi); d = lo - mean[0]; s[0] += d * d; d = hi - mean[1]; s[1] += d * d;
      unpack_bf16x2(v[i].y, lo, hi); d = lo - mean[2]; s[2] += d * d; d = hi - mean[3]; s[3] += d * d;
      unpack_bf16x2(v[i].z, lo, hi); d = lo - mean[4]; s[4] += d * d; d = hi - mean[5]; s[5] += d * d;
      unpack_bf16x2(v[i].w, lo, hi); d = lo - mean[6]; s[6] += d * d; d = hi - mean[7]; s[7] += d * d;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 4);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
  }
  if (lane < 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = s[j];
  }
  __syncthreads();
  if (tid < 32) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w][tid];
    chan_stat[tid] = t;
  }
  __syncthreads();
  if (tid < 32) {
    const int g0 = (tid / cpg) * cpg;
    float t = 0.f;
    for (int j = 0; j < cpg; ++j) t += chan_stat[g0 + j];
    grstd[tid] = rsqrtf(t * inv_cnt + p.eps);
  }
  __syncthreads();

  // ---- per-channel affine folded into a*x + b ----
  float a[8], b[8], m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cbase + chunk * 8 + j;
    const float rs = grstd[chunk * 8 + j];
    float ga = p.gamma ? p.gamma[c] : 1.f, be = p.beta ? p.beta[c] : 0.f;
    float aa = rs * ga, bb = be - mean[j] * rs * ga;
    if (p.scale) {
      const long long r = static_cast<long long>(p.ss_rows == 1 ? 0 : n) * p.ss_ld;
      const float sc = 1.f + p.scale[r + c], sh = p.shift[r + c];
      aa *= sc;
      bb = bb * sc + sh;
    }
    a[j] = aa; b[j] = bb;
    m[j] = p.mask ? p.mask[static_cast<long long>(n) * C + c] : 1.f;
  }

  __nv_bfloat16* obase = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(n) * p.hw * C + cbase + chunk * 8;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) {
      float f[8];
      unpack_bf16x2(v[i].x, f[0], f[1]);
      unpack_bf16x2(v[i].y, f[2], f[3]);
      unpack_bf16x2(v[i].z, f[4], f[5]);
      unpack_bf16x2(v[i].w, f[6], f[7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float yv = fmaf(f[j], a[j], b[j]);
        if (p.silu) yv = silu_f(yv);
        f[j] = yv * m[j];
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(obase + static_cast<long long>(vi >> 2) * C) = o;
    }
  }
}

// --------------------------------------------------------------------------------------------
// generic path: one CTA per (image, group)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_generic_kernel(const GnParams p) {
  __shared__ float scratch[8];
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const int n = blockIdx.x / p.groups;
  const int g = blockIdx.x - n * p.groups;
  const int cnt = p.hw * cpg;
  const T* s0 = static_cast<const T*>(p.src0);
  const T* s1 = static_cast<const T*>(p.src1);

  auto load = [&](int e) -> float {
    const int l = e / cpg, c = g * cpg + (e - l * cpg);
    if (c < p.c0) return ld_act<T>(s0 + (static_cast<long long>(n) * p.hw + l) * p.c0 + c);
    return ld_act<T>(s1 + (static_cast<long long>(n) * p.hw + l) * p.c1 + (c - p.c0));
  };

  float s = 0.f;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) s += load(e);
  const float mean = block_sum(s, scratch) / cnt;
  float q = 0.f;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) { const float d = load(e) - mean; q += d * d; }
  const float rstd = rsqrtf(block_sum(q, scratch) / cnt + p.eps);

  T* out = static_cast<T*>(p.out);
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int l = e / cpg, c = g * cpg + (e - l * cpg);
    float y = (load(e) - mean) * rstd;
    if (p.gamma) y = y * p.gamma[c] + (p.beta ? p.beta[c] : 0.f);
    if (p.scale) {
      const long long r = static_cast<long long>(p.ss_rows == 1 ? 0 : n) * p.ss_ld;
      y = y * (p.scale[r + c] + 1.f) + p.shift[r + c];
    }
    if (p.silu) y = silu_precise(y);
    if (p.mask) y *= p.mask[static_cast<long long>(n) * C + c];
    st_act<T>(out + (static_cast<long long>(n) * p.hw + l) * C + c, y);
  }
}

// --------------------------------------------------------------------------------------------
// streaming path: statistics were accumulated by the producing convolution's epilogue
// (dmme_conv_desc.stats), so GroupNorm is one elementwise pass: 16-byte loads, y = silu(a*x + b), 16-byte stores
// --------------------------------------------------------------------------------------------
struct GnApplyParams {
  GnParams g;
  const long long* stats0; const long long* stats1;
  int chunks;  // CTAs per image
};

// y = a * x + b of channel c of image n: GroupNorm statistics (fixed-point micro-group sums of the producing conv) folded
// with gamma / beta and the optional IDDPM scale / shift.  One function for the streaming kernel and for the coefficient
// kernel of the conv-fused path, so both produce the same bits.
__device__ __forceinline__ void gn_coefficients(const GnApplyParams& q, int n, int c, float& aa, float& bb) {
  const GnParams& p = q.g;
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const float inv_cnt = 1.0f / (static_cast<float>(p.hw) * cpg);
  const double unfix = 1.0 / static_cast<double>(1 << DMME_STATS_FRAC_BITS);
  const int g0 = (c / cpg) * cpg;  // first channel of this channel's group
  long long s1 = 0, s2 = 0;
  for (int cc = g0; cc < g0 + cpg; cc += 4) {
    const long long* st = cc < p.c0 ? q.stats0 + (static_cast<long long>(n) * (p.c0 >> 2) + (cc >> 2)) * 2
                                    : q.stats1 + (static_cast<long long>(n) * (p.c1 >> 2) + ((cc - p.c0) >> 2)) * 2;
    s1 += st[0];
    s2 += st[1];
  }
  const float mean = static_cast<float>(static_cast<double>(s1) * unfix) * inv_cnt;
  const float ex2 = static_cast<float>(static_cast<double>(s2) * unfix) * inv_cnt;
  const float var = fmaxf(ex2 - mean * mean, 0.f);
  const float rs = rsqrtf(var + p.eps);
  const float ga = p.gamma ? p.gamma[c] : 1.f, be = p.beta ? p.beta[c] : 0.f;
  aa = rs * ga;
  bb = be - mean * rs * ga;
  if (p.scale) {
    const long long r = static_cast<long long>(p.ss_rows == 1 ? 0 : n) * p.ss_ld;
    const float sc = 1.f + p.scale[r + c], sh = p.shift[r + c];
    aa *= sc;
    bb = bb * sc + sh;
  }
}

// coefficients only, interleaved (a, b) per (image, channel): consumed by the halo conv kernel's fused GroupNorm stage
__global__ void gn_coeff_kernel(const GnApplyParams q, float2* __restrict__ ab) {
  pdl_trigger();
  pdl_wait();
  const int C = q.g.c0 + q.g.c1;
  const long long total = static_cast<long long>(q.g.n) * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / C), c = static_cast<int>(i - static_cast<long long>(n) * C);
    float aa, bb;
    gn_coefficients(q, n, c, aa, bb);
    ab[i] = make_float2(aa, bb);
  }
}

// Persistent layout: the whole tensor is one range of 16-byte vectors split evenly over the grid (one wave of CTAs, four
// per SM); a CTA walks its range image by image and rebuilds the per-channel coefficients only when the image changes.
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const GnApplyParams q) {
  extern __shared__ float ab[];  // a[C], b[C], m[C]
  const GnParams& p = q.g;
  const int C = p.c0 + p.c1;
  float* sa = ab;
  float* sb = ab + C;
  float* sm = ab + 2 * C;
  const int cv = C >> 3;  // 16-byte vectors per pixel
  const long long nvec = static_cast<long long>(p.hw) * cv;      // per image
  const long long total = nvec * p.n;
  long long per = (total + gridDim.x - 1) / gridDim.x;
  per = (per + 255) / 256 * 256;
  const long long r0 = blockIdx.x * per;
  const long long r1 = r0 + per < total ? r0 + per : total;
  pdl_trigger();
  pdl_wait();  // statistics and activations come from the producing conv

  for (long long seg = r0; seg < r1;) {
    const int n = static_cast<int>(seg / nvec);
    const long long v0 = seg - static_cast<long long>(n) * nvec;
    const long long v1 = (r1 - static_cast<long long>(n) * nvec) < nvec ? (r1 - static_cast<long long>(n) * nvec) : nvec;
    __syncthreads();  // the previous image's coefficients are no longer read
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float aa, bb;
      gn_coefficients(q, n, c, aa, bb);
      sa[c] = aa;
      sb[c] = bb;
      sm[c] = p.mask ? p.mask[static_cast<long long>(n) * C + c] : 1.f;
    }
    __syncthreads();

    const __nv_bfloat16* s0 = static_cast<const __nv_bfloat16*>(p.src0) + static_cast<long long>(n) * p.hw * p.c0;
    const __nv_bfloat16* s1p = p.c1 ? static_cast<const __nv_bfloat16*>(p.src1) + static_cast<long long>(n) * p.hw * p.c1 : nullptr;
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(n) * p.hw * C;
    // four independent 16-byte loads in flight per thread, then the arithmetic, then four stores
    auto apply = [&](const uint4& v, int c, long long o) {
      float f[8];
      unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
      unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
      const float4 a0 = *reinterpret_cast<const float4*>(sa + c), a1 = *reinterpret_cast<const float4*>(sa + c + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sb + c), b1 = *reinterpret_cast<const float4*>(sb + c + 4);
      f[0] = fmaf(f[0], a0.x, b0.x); f[1] = fmaf(f[1], a0.y, b0.y); f[2] = fmaf(f[2], a0.z, b0.z); f[3] = fmaf(f[3], a0.w, b0.w);
      f[4] = fmaf(f[4], a1.x, b1.x); f[5] = fmaf(f[5], a1.y, b1.y); f[6] = fmaf(f[6], a1.z, b1.z); f[7] = fmaf(f[7], a1.w, b1.w);
      if (p.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
      }
      if (p.mask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= sm[c + j];
      }
      uint4 o4;
      o4.x = pack_bf16x2(f[0], f[1]); o4.y = pack_bf16x2(f[2], f[3]);
      o4.z = pack_bf16x2(f[4], f[5]); o4.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(out + o) = o4;
    };
    // (pixel, vector-in-pixel) of this thread's current vector; successive vectors are 256 apart, so the pair advances
    // by a constant step with one conditional carry -- no division in the streaming loop
    const int step_p = 256 / cv, step_c = 256 - step_p * cv;
    int pixel = static_cast<int>((v0 + threadIdx.x) / cv);
    int cidx = static_cast<int>((v0 + threadIdx.x) - static_cast<long long>(pixel) * cv);
    auto src_of = [&](int& c, long long& o) -> const uint4* {
      c = cidx << 3;
      o = static_cast<long long>(pixel) * C + c;
      const uint4* ptr = c < p.c0 ? reinterpret_cast<const uint4*>(s0 + static_cast<long long>(pixel) * p.c0 + c)
                                  : reinterpret_cast<const uint4*>(s1p + static_cast<long long>(pixel) * p.c1 + (c - p.c0));
      pixel += step_p;
      cidx += step_c;
      if (cidx >= cv) { cidx -= cv; ++pixel; }
      return ptr;
    };
    constexpr int U = 4;
    long long vi = v0 + threadIdx.x;
    if (step_c == 0) {
      // 256 is a multiple of the vectors per pixel (C = 128, 256, 512): this thread meets the SAME eight channels in
      // every vector, so their coefficients are read from shared memory once per image instead of per vector (the
      // four 16-byte coefficient reads per vector kept the shared-memory pipe 53% busy in the ncu capture)
      const int c = cidx << 3;
      float ca[8], cb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { ca[j] = sa[c + j]; cb[j] = sb[c + j]; }
      const bool from0 = c < p.c0;
      const __nv_bfloat16* sp = from0 ? s0 + c : s1p + (c - p.c0);
      const int sc = from0 ? p.c0 : p.c1;
      auto apply_fixed = [&](const uint4& v, int px) {
        float f[8];
        unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
        unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], ca[j], cb[j]);
        if (p.silu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
        }
        if (p.mask) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] *= sm[c + j];  // dropout masks: training only, not worth eight registers
        }
        uint4 o4;
        o4.x = pack_bf16x2(f[0], f[1]); o4.y = pack_bf16x2(f[2], f[3]);
        o4.z = pack_bf16x2(f[4], f[5]); o4.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(out + static_cast<long long>(px) * C + c) = o4;
      };
      for (; vi + (U - 1) * 256 < v1; vi += U * 256) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(sp + static_cast<long long>(pixel + u * step_p) * sc));
#pragma unroll
        for (int u = 0; u < U; ++u) apply_fixed(v[u], pixel + u * step_p);
        pixel += U * step_p;
      }
      for (; vi < v1; vi += 256) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(sp + static_cast<long long>(pixel) * sc));
        apply_fixed(v, pixel);
        pixel += step_p;
      }
      seg = static_cast<long long>(n) * nvec + v1;
      continue;
    }
    for (; vi + (U - 1) * 256 < v1; vi += U * 256) {
      uint4 v[U];
      int c[U];
      long long o[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldg(src_of(c[u], o[u]));
#pragma unroll
      for (int u = 0; u < U; ++u) apply(v[u], c[u], o[u]);
    }
    for (; vi < v1; vi += 256) {
      int c;
      long long o;
      const uint4 v = __ldg(src_of(c, o));
      apply(v, c, o);
    }
    seg = static_cast<long long>(n) * nvec + v1;
  }
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_groupnorm_fwd(const void* src0, const void* src1, int c0, int c1, int n, int hw, int groups,
                                  float eps, const float* gamma, const float* beta, const float* scale,
                                  const float* shift, int ss_rows, int ss_ld, const float* chan_mask, int apply_silu,
                                  void* out, int act_dtype, const long long* stats0, const long long* stats1,
                                  void* stream) {
  DMME_REQUIRE(src0 && out, DMME_E_BADARG, "groupnorm: null src0/out");
  DMME_REQUIRE(n > 0 && hw > 0 && c0 > 0 && c1 >= 0 && groups > 0, DMME_E_BADARG, "groupnorm: bad sizes");
  DMME_REQUIRE(c1 == 0 || src1, DMME_E_BADARG, "groupnorm: c1 > 0 but src1 is null");
  const int C = c0 + c1;
  DMME_REQUIRE(C % groups == 0, DMME_E_SHAPE, "groupnorm: C=%d not divisible by groups=%d", C, groups);
  DMME_REQUIRE((scale == nullptr) == (shift == nullptr), DMME_E_BADARG, "groupnorm: scale and shift come together");
  GnParams p;
  p.src0 = src0; p.src1 = src1; p.c0 = c0; p.c1 = c1; p.n = n; p.hw = hw; p.groups = groups; p.eps = eps;
  p.gamma = gamma; p.beta = beta; p.scale = scale; p.shift = shift; p.ss_rows = ss_rows; p.ss_ld = ss_ld;
  p.mask = chan_mask; p.silu = apply_silu; p.out = out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  const bool have_stats = stats0 != nullptr && (c1 == 0 || stats1 != nullptr);
  if (have_stats && act_dtype == DMME_BF16 && c0 % 8 == 0 && c1 % 8 == 0 && cpg % 4 == 0 && c0 % cpg == 0 && C <= 4096) {
    GnApplyParams q;
    q.g = p; q.stats0 = stats0; q.stats1 = stats1;
    const long long total_vec = static_cast<long long>(n) * hw * (C >> 3);
    long long grid = ceil_div_ll(total_vec, 256);  // at least one vector per thread
    if (grid > 148 * 4) grid = 148 * 4;            // one wave, four CTAs per SM
    q.chunks = 0;
    cudaError_t e = launch_pdl(gn_apply_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), sizeof(float) * 3 * C, st, q);
    return check_launch_err(e, "gn_apply_kernel");
  }
  const bool fast = act_dtype == DMME_BF16 && C % 32 == 0 && c0 % 32 == 0 && (32 % cpg == 0) && hw <= 1024 &&
                    (hw * 4) % 32 == 0;
  if (fast) {
    const int nvec = hw * 4;
    const int threads = nvec < 256 ? nvec : 256;
    const int maxv = ceil_div(nvec, threads);
    const int blocks = n * (C / 32);
    if (maxv <= 1) gn_slab_kernel<1><<<blocks, threads, 0, st>>>(p);
    else if (maxv <= 4) gn_slab_kernel<4><<<blocks, threads, 0, st>>>(p);
    else gn_slab_kernel<16><<<blocks, threads, 0, st>>>(p);
    return check_launch("gn_slab_kernel");
  }
  if (act_dtype == DMME_BF16) gn_generic_kernel<__nv_bfloat16><<<n * groups, 256, 0, st>>>(p);
  else gn_generic_kernel<float><<<n * groups, 256, 0, st>>>(p);
  return check_launch("gn_generic_kernel");
}

// (a, b) coefficients of y = a * x + b per (image, channel), interleaved fp32 pairs [n][c0 + c1][2], from the statistics
// the producing convolutions wrote: the input of the conv-fused GroupNorm path (dmme_conv_desc.gn_ab)
extern "C" int dmme_groupnorm_coeff(const long long* stats0, const long long* stats1, int c0, int c1, int n, int hw,
                                    int groups, float eps, const float* gamma, const float* beta, const float* scale,
                                    const float* shift, int ss_rows, int ss_ld, float* ab_out, void* stream) {
  DMME_REQUIRE(stats0 && ab_out && n > 0 && hw > 0 && c0 > 0 && c1 >= 0 && groups > 0, DMME_E_BADARG, "groupnorm_coeff: bad arguments");
  DMME_REQUIRE(c1 == 0 || stats1, DMME_E_BADARG, "groupnorm_coeff: c1 > 0 but stats1 is null");
  const int C = c0 + c1;
  DMME_REQUIRE(C % groups == 0, DMME_E_SHAPE, "groupnorm_coeff: C=%d not divisible by groups=%d", C, groups);
  const int cpg = C / groups;
  DMME_REQUIRE(cpg % 4 == 0 && c0 % cpg == 0, DMME_E_SHAPE, "groupnorm_coeff: groups must be whole 4-channel micro-groups of one source");
  GnApplyParams q;
  memset(&q, 0, sizeof(q));
  q.g.c0 = c0; q.g.c1 = c1; q.g.n = n; q.g.hw = hw; q.g.groups = groups; q.g.eps = eps;
  q.g.gamma = gamma; q.g.beta = beta; q.g.scale = scale; q.g.shift = shift; q.g.ss_rows = ss_rows; q.g.ss_ld = ss_ld;
  q.stats0 = stats0; q.stats1 = stats1;
  const long long total = static_cast<long long>(n) * C;
  const int grid = static_cast<int>(ceil_div_ll(total, 256) < 148 * 4 ? ceil_div_ll(total, 256) : 148 * 4);
  cudaError_t e = launch_pdl(gn_coeff_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), q,
                             reinterpret_cast<float2*>(ab_out));
  return check_launch_err(e, "gn_coeff_kernel");
}
