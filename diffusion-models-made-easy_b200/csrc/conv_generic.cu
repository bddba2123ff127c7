// Shape-agnostic convolution on CUDA cores (fp32 FFMA implicit GEMM).
//
// This is the path for everything the tensor-core kernel does not take: channel counts that are not
// multiples of 64 (the reference's test fixture uses 3/4/8/16/32/64 channels), the 3-channel image
// input conv (K = 27), the 3/6-channel output conv, nearest-x2 upsampling folded into the gather,
// and the fp32 parity mode (activations stored fp32, rel-L2 <= 1e-4 against the reference).
// Same fused semantics as conv_tc.cu: two-source concat, fused 1x1 residual conv, bias/temb/addend
// epilogue.  Weights are packed fp32 [K][cout].
#include "common.cuh"

namespace dmme {

struct ConvGenParams {
  const void* src0; const void* src1; int c0, c1;
  const void* res0; const void* res1; int rc0, rc1;
  int n, h_in, w_in, ho, wo;
  int ksize, stride, upsample;
  int cout;
  const float* weight;  // [K][cout]
  const float* bias; const float* temb; int temb_rows, temb_ld;
  const void* addend;
  void* out; void* out2; void* out3;
  int in_layout, out_layout;
};

constexpr int GT_M = 64, GT_N = 64, GT_K = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv_generic_kernel(const ConvGenParams p) {
  __shared__ float As[GT_K][GT_M + 4];
  __shared__ float Bs[GT_K][GT_N + 4];

  const int tid = threadIdx.x;
  const int ctot = p.c0 + p.c1;
  const int kconv = p.ksize * p.ksize * ctot;
  const int ktot = kconv + p.rc0 + p.rc1;
  const long long mtot = static_cast<long long>(p.n) * p.ho * p.wo;
  const long long m0 = static_cast<long long>(blockIdx.x) * GT_M;
  const int n0 = blockIdx.y * GT_N;
  const int pad = p.ksize / 2;
  const int hin_eff = p.upsample ? 2 * p.h_in : p.h_in;
  const int win_eff = p.upsample ? 2 * p.w_in : p.w_in;

  // A-load mapping: kk = tid % 16 (consecutive channels -> contiguous in NHWC), 4 pixels per thread
  const int a_kk = tid & 15;
  int a_n[4], a_y[4], a_x[4];
  bool a_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long m = m0 + (tid >> 4) + 16 * j;
    a_ok[j] = m < mtot;
    const long long mm = a_ok[j] ? m : 0;
    a_x[j] = static_cast<int>(mm % p.wo);
    a_y[j] = static_cast<int>((mm / p.wo) % p.ho);
    a_n[j] = static_cast<int>(mm / (static_cast<long long>(p.wo) * p.ho));
  }
  // B-load mapping: nn = tid % 64, kk = tid / 64 + 4 j
  const int b_nn = tid & 63;

  const int tx = tid & 15, ty = tid >> 4;  // compute mapping: pixels ty*4.., couts tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < ktot; k0 += GT_K) {
    // ---- gather A ----
    {
      const int k = k0 + a_kk;
      int tap = 0, ci = 0;
      bool is_res = false;
      if (k < kconv) {
        tap = k / ctot;
        ci = k - tap * ctot;
      } else {
        is_res = true;
        ci = k - kconv;
      }
      const int r = tap / p.ksize, s = tap - r * p.ksize;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (a_ok[j] && k < ktot) {
          if (!is_res) {
            const int iy = a_y[j] * p.stride + r - pad;
            const int ix = a_x[j] * p.stride + s - pad;
            // upsample == 2: zero-dilated source (values at even coordinates only) = the data gradient of a stride-2 conv
            if (iy >= 0 && iy < hin_eff && ix >= 0 && ix < win_eff && (p.upsample != 2 || ((iy | ix) & 1) == 0)) {
              const int sy = p.upsample ? (iy >> 1) : iy;
              const int sx = p.upsample ? (ix >> 1) : ix;
              if (p.in_layout == DMME_IN_NCHW_F32) {
                const float* s0 = static_cast<const float*>(p.src0);
                v = s0[((static_cast<long long>(a_n[j]) * p.c0 + ci) * p.h_in + sy) * p.w_in + sx];
              } else {
                const long long pixi = (static_cast<long long>(a_n[j]) * p.h_in + sy) * p.w_in + sx;
                if (ci < p.c0) v = ld_act<T>(static_cast<const T*>(p.src0) + pixi * p.c0 + ci);
                else v = ld_act<T>(static_cast<const T*>(p.src1) + pixi * p.c1 + (ci - p.c0));
              }
            }
          } else {
            const long long pixo = (static_cast<long long>(a_n[j]) * p.ho + a_y[j]) * p.wo + a_x[j];
            if (ci < p.rc0) v = ld_act<T>(static_cast<const T*>(p.res0) + pixo * p.rc0 + ci);
            else v = ld_act<T>(static_cast<const T*>(p.res1) + pixo * p.rc1 + (ci - p.rc0));
          }
        }
        As[a_kk][(tid >> 4) + 16 * j] = v;
      }
    }
    // ---- load B ----
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = (tid >> 6) + 4 * j;
      const int k = k0 + kk;
      const int co = n0 + b_nn;
      Bs[kk][b_nn] = (k < ktot && co < p.cout) ? p.weight[static_cast<long long>(k) * p.cout + co] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GT_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int L = p.ho * p.wo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= mtot) continue;
    const int x = static_cast<int>(m % p.wo);
    const int y = static_cast<int>((m / p.wo) % p.ho);
    const int n = static_cast<int>(m / (static_cast<long long>(p.wo) * p.ho));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[co];
      if (p.temb) v += p.temb[static_cast<long long>(p.temb_rows == 1 ? 0 : n) * p.temb_ld + co];
      if (p.addend) v += ld_act<T>(static_cast<const T*>(p.addend) + m * p.cout + co);
      if (p.out_layout == DMME_OUT_NCHW_F32) {
        static_cast<float*>(p.out)[((static_cast<long long>(n) * p.cout + co) * p.ho + y) * p.wo + x] = v;
      } else if (p.out_layout == DMME_OUT_QKV) {
        const int c = p.cout / 3;
        const int which = co / c, cc = co - which * c;
        if (which == 0) st_act<T>(static_cast<T*>(p.out) + m * c + cc, v);
        else if (which == 1) st_act<T>(static_cast<T*>(p.out2) + m * c + cc, v);
        else st_act<T>(static_cast<T*>(p.out3) + (static_cast<long long>(n) * c + cc) * L + (y * p.wo + x), v);
      } else {
        st_act<T>(static_cast<T*>(p.out) + m * p.cout + co, v);
      }
    }
  }
}

int conv_generic_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_generic: null src0/weight/out");
  DMME_REQUIRE(d.ksize == 1 || d.ksize == 3, DMME_E_SHAPE, "conv_generic: ksize must be 1 or 3 (got %d)", d.ksize);
  DMME_REQUIRE(d.stride == 1 || d.stride == 2, DMME_E_SHAPE, "conv_generic: stride must be 1 or 2");
  DMME_REQUIRE(!(d.upsample && d.stride != 1), DMME_E_SHAPE, "conv_generic: upsample needs stride 1");
  DMME_REQUIRE(d.upsample >= 0 && d.upsample <= 2, DMME_E_BADARG, "conv_generic: upsample must be 0, 1 (nearest) or 2 (zero-dilated)");
  DMME_REQUIRE(d.n > 0 && d.h_in > 0 && d.w_in > 0 && d.c0 > 0 && d.cout > 0, DMME_E_BADARG, "conv_generic: bad sizes");
  DMME_REQUIRE(d.c1 == 0 || d.src1, DMME_E_BADARG, "conv_generic: c1 > 0 but src1 is null");
  DMME_REQUIRE(d.rc0 == 0 || d.res0, DMME_E_BADARG, "conv_generic: rc0 > 0 but res0 is null");
  DMME_REQUIRE(d.rc1 == 0 || d.res1, DMME_E_BADARG, "conv_generic: rc1 > 0 but res1 is null");
  DMME_REQUIRE(d.in_layout == DMME_IN_NHWC || (d.c1 == 0), DMME_E_SHAPE, "conv_generic: NCHW input has one source");
  DMME_REQUIRE(d.out_layout != DMME_OUT_QKV || (d.cout % 3 == 0 && d.out2 && d.out3), DMME_E_BADARG,
               "conv_generic: QKV output needs cout %% 3 == 0 and out2/out3");
  ConvGenParams p;
  p.src0 = d.src0; p.src1 = d.src1; p.c0 = d.c0; p.c1 = d.c1;
  p.res0 = d.res0; p.res1 = d.res1; p.rc0 = d.rc0; p.rc1 = d.rc1;
  p.n = d.n; p.h_in = d.h_in; p.w_in = d.w_in;
  const int hin_eff = d.upsample ? 2 * d.h_in : d.h_in, win_eff = d.upsample ? 2 * d.w_in : d.w_in;
  const int pad = d.ksize / 2;
  p.ho = (hin_eff + 2 * pad - d.ksize) / d.stride + 1;
  p.wo = (win_eff + 2 * pad - d.ksize) / d.stride + 1;
  p.ksize = d.ksize; p.stride = d.stride; p.upsample = d.upsample;
  p.cout = d.cout;
  p.weight = static_cast<const float*>(d.weight);
  p.bias = d.bias; p.temb = d.temb; p.temb_rows = d.temb_rows; p.temb_ld = d.temb_ld;
  p.addend = d.addend;
  p.out = d.out; p.out2 = d.out2; p.out3 = d.out3;
  p.in_layout = d.in_layout; p.out_layout = d.out_layout;
  const long long mtot = static_cast<long long>(p.n) * p.ho * p.wo;
  dim3 grid(static_cast<unsigned>(ceil_div_ll(mtot, GT_M)), ceil_div(d.cout, GT_N));
  if (d.act_dtype == DMME_BF16) conv_generic_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  else conv_generic_kernel<float><<<grid, 256, 0, stream>>>(p);
  return check_launch("conv_generic_kernel");
}

// ------------------------------------------------------------------------------------------------
// weight packing and layout helpers
// ------------------------------------------------------------------------------------------------
// k = tap * cin + ci (tap = r * ksize + s), residual rows appended: k = ksize^2 * cin + cr
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int ksize,
                                   const float* __restrict__ wres, int rc, void* __restrict__ packed, int tc) {
  const int taps = ksize * ksize;
  const long long ktot = static_cast<long long>(taps) * cin + rc;
  const long long total = ktot * cout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long k;
    int co;
    if (tc) { co = static_cast<int>(i / ktot); k = i - co * ktot; }   // [cout][K]
    else { k = i / cout; co = static_cast<int>(i - k * cout); }       // [K][cout]
    float v;
    if (k < static_cast<long long>(taps) * cin) {
      const int tap = static_cast<int>(k / cin), ci = static_cast<int>(k - static_cast<long long>(tap) * cin);
      v = w[(static_cast<long long>(co) * cin + ci) * taps + tap];
    } else {
      v = wres[static_cast<long long>(co) * rc + (k - static_cast<long long>(taps) * cin)];
    }
    if (tc) static_cast<__nv_bfloat16*>(packed)[i] = __float2bfloat16_rn(v);
    else static_cast<float*>(packed)[i] = v;
  }
}

// Data-gradient weights: the dgrad of conv(w) w.r.t. input channels [ci_off, ci_off + ci_cnt) is itself a convolution of
// grad_out with  w'[co' = ci - ci_off][k' = tap' * cout + co] = w[co][ci][taps - 1 - tap']  (spatial flip, in/out swapped).
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, int cout, int cin, int ksize, int ci_off, int ci_cnt,
                                         void* __restrict__ packed, int tc) {
  const int taps = ksize * ksize;
  const long long ktot = static_cast<long long>(taps) * cout;
  const long long total = ktot * ci_cnt;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long k;
    int cn;
    if (tc) { cn = static_cast<int>(i / ktot); k = i - cn * ktot; }     // [ci_cnt][K']
    else { k = i / ci_cnt; cn = static_cast<int>(i - k * ci_cnt); }     // [K'][ci_cnt]
    const int tap = static_cast<int>(k / cout), co = static_cast<int>(k - static_cast<long long>(tap) * cout);
    const float v = w[(static_cast<long long>(co) * cin + (ci_off + cn)) * taps + (taps - 1 - tap)];
    if (tc) static_cast<__nv_bfloat16*>(packed)[i] = __float2bfloat16_rn(v);
    else static_cast<float*>(packed)[i] = v;
  }
}

// Every bf16 tensor-core weight pack of a model in ONE launch (the graph-captured training step re-packs all weights after
// each optimizer step: 172 launches of the two kernels above, 0.9 ms of an 18 ms step).  items[] lives in device memory; a
// block of 256 threads converts kPackPerBlock consecutive packed elements of one item, found by bisection over first_block.
constexpr int kPackPerBlock = 2048;
__global__ void __launch_bounds__(256) pack_batch_kernel(const dmme_pack_item* __restrict__ items, int n_items) {
  int lo = 0, hi = n_items - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {  // the last item whose first_block <= b
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  const dmme_pack_item it = items[lo];
  const int taps = it.ksize * it.ksize;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(it.packed);
  const long long i0 = (b - it.first_block) * kPackPerBlock;
  if (taps == 9) {
    // 3x3 weights (all but a few per cent of the elements): a thread owns one (cout, cin) pair, reads its nine taps -- 36
    // contiguous bytes of the OIHW tensor -- and writes those of them that fall into the block's range of packed elements;
    // consecutive threads write consecutive packed elements.  (Element by element, as the 1x1 path below does, every load
    // fetched a 32-byte sector for four bytes, nine times over: 295 us per training step.)
    const int cin = it.cin, cout = it.cout;
    if (!it.dgrad) {
      const long long conv_k = 9ll * cin, ktot = conv_k + it.rc, total = ktot * cout;
      const long long i1 = i0 + kPackPerBlock < total ? i0 + kPackPerBlock : total;
      for (long long co = i0 / ktot; co * ktot < i1; ++co) {
        const long long row0 = co * ktot;
        const long long ka = (i0 > row0 ? i0 : row0) - row0, kb = (i1 < row0 + ktot ? i1 : row0 + ktot) - row0;
        const long long cb = kb < conv_k ? kb : conv_k;
        if (ka < cb) {
          for (int ci = threadIdx.x; ci < cin; ci += 256) {
            const float* src = it.w + (co * cin + ci) * 9;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const long long k = static_cast<long long>(tap) * cin + ci;
              if (k >= ka && k < cb) out[row0 + k] = __float2bfloat16_rn(src[tap]);
            }
          }
        }
        for (long long k = (ka > conv_k ? ka : conv_k) + threadIdx.x; k < kb; k += 256)
          out[row0 + k] = __float2bfloat16_rn(it.wres[co * it.rc + (k - conv_k)]);
      }
    } else {
      const long long ktot = 9ll * cout, total = ktot * it.ci_cnt;
      const long long i1 = i0 + kPackPerBlock < total ? i0 + kPackPerBlock : total;
      for (long long cn = i0 / ktot; cn * ktot < i1; ++cn) {
        const long long row0 = cn * ktot;
        const long long ka = (i0 > row0 ? i0 : row0) - row0, kb = (i1 < row0 + ktot ? i1 : row0 + ktot) - row0;
        for (int co = threadIdx.x; co < cout; co += 256) {
          const float* src = it.w + (static_cast<long long>(co) * cin + (it.ci_off + cn)) * 9;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const long long k = static_cast<long long>(tap) * cout + co;
            if (k >= ka && k < kb) out[row0 + k] = __float2bfloat16_rn(src[8 - tap]);
          }
        }
      }
    }
    return;
  }
  if (!it.dgrad) {
    const long long ktot = static_cast<long long>(taps) * it.cin + it.rc, total = ktot * it.cout;
    for (long long i = i0 + threadIdx.x; i < i0 + kPackPerBlock && i < total; i += 256) {
      const int co = static_cast<int>(i / ktot);
      const long long k = i - co * ktot;
      float v;
      if (k < static_cast<long long>(taps) * it.cin) {
        const int tap = static_cast<int>(k / it.cin), ci = static_cast<int>(k - static_cast<long long>(tap) * it.cin);
        v = it.w[(static_cast<long long>(co) * it.cin + ci) * taps + tap];
      } else {
        v = it.wres[static_cast<long long>(co) * it.rc + (k - static_cast<long long>(taps) * it.cin)];
      }
      out[i] = __float2bfloat16_rn(v);
    }
  } else {
    const long long ktot = static_cast<long long>(taps) * it.cout, total = ktot * it.ci_cnt;
    for (long long i = i0 + threadIdx.x; i < i0 + kPackPerBlock && i < total; i += 256) {
      const int cn = static_cast<int>(i / ktot);
      const long long k = i - cn * ktot;
      const int tap = static_cast<int>(k / it.cout), co = static_cast<int>(k - static_cast<long long>(tap) * it.cout);
      out[i] = __float2bfloat16_rn(it.w[(static_cast<long long>(co) * it.cin + (it.ci_off + cn)) * taps + (taps - 1 - tap)]);
    }
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int n, int c, int hw) {
  const long long total = static_cast<long long>(n) * c * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c);
    const long long pix = i / c;
    const int l = static_cast<int>(pix % hw);
    const long long ni = pix / hw;
    st_act<T>(dst + i, src[(ni * c + ci) * hw + l]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int n, int c, int hw) {
  const long long total = static_cast<long long>(n) * c * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = static_cast<int>(i % hw);
    const long long nc = i / hw;
    const int ci = static_cast<int>(nc % c);
    const long long ni = nc / c;
    dst[i] = ld_act<T>(src + (ni * hw + l) * c + ci);
  }
}
// nearest x2, NHWC; one thread per 16-byte (or single element) output vector
template <typename T, int VEC>
__global__ void upsample2x_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int h, int w, int c) {
  pdl_trigger();
  pdl_wait();
  const int cv = c / VEC;
  const long long total = static_cast<long long>(n) * (2 * h) * (2 * w) * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    long long pix = i / cv;
    const int x = static_cast<int>(pix % (2 * w));
    pix /= (2 * w);
    const int y = static_cast<int>(pix % (2 * h));
    const long long ni = pix / (2 * h);
    const T* s = src + ((ni * h + (y >> 1)) * w + (x >> 1)) * c + v * VEC;
    T* d = dst + i * VEC;
    if (VEC * sizeof(T) == 16) *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(s));
    else *d = *s;
  }
}

static int grid_for(long long total, int threads) {
  long long b = ceil_div_ll(total, threads);
  const long long cap = 148LL * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, const float* w_res, int rc,
                                     void* packed, int kernel, void* stream) {
  DMME_REQUIRE(w_oihw && packed, DMME_E_BADARG, "pack_conv_weight: null pointer");
  DMME_REQUIRE(cout > 0 && cin > 0 && (ksize == 1 || ksize == 3) && rc >= 0, DMME_E_BADARG, "pack_conv_weight: bad sizes");
  DMME_REQUIRE(rc == 0 || w_res, DMME_E_BADARG, "pack_conv_weight: rc > 0 but w_res is null");
  DMME_REQUIRE(kernel == DMME_CONV_TC || kernel == DMME_CONV_GENERIC, DMME_E_BADARG, "pack_conv_weight: kernel must be TC or GENERIC");
  const long long total = (static_cast<long long>(ksize) * ksize * cin + rc) * cout;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, cout, cin, ksize, w_res, rc, packed, kernel == DMME_CONV_TC ? 1 : 0);
  return check_launch("pack_weight_kernel");
}

extern "C" int dmme_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin, int ksize, int ci_off, int ci_cnt,
                                           void* packed, int kernel, void* stream) {
  DMME_REQUIRE(w_oihw && packed, DMME_E_BADARG, "pack_conv_weight_dgrad: null pointer");
  DMME_REQUIRE(cout > 0 && cin > 0 && (ksize == 1 || ksize == 3) && ci_off >= 0 && ci_cnt > 0 && ci_off + ci_cnt <= cin,
               DMME_E_BADARG, "pack_conv_weight_dgrad: bad sizes");
  DMME_REQUIRE(kernel == DMME_CONV_TC || kernel == DMME_CONV_GENERIC, DMME_E_BADARG, "pack_conv_weight_dgrad: kernel must be TC or GENERIC");
  const long long total = static_cast<long long>(ksize) * ksize * cout * ci_cnt;
  pack_weight_dgrad_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, cout, cin, ksize, ci_off, ci_cnt, packed, kernel == DMME_CONV_TC ? 1 : 0);
  return check_launch("pack_weight_dgrad_kernel");
}

extern "C" int dmme_pack_block_elems(void) { return kPackPerBlock; }

extern "C" int dmme_pack_conv_weights_batch(const dmme_pack_item* items_dev, int n_items, long long total_blocks, void* stream) {
  DMME_REQUIRE(items_dev && n_items > 0 && total_blocks > 0 && total_blocks < (1ll << 31), DMME_E_BADARG,
               "pack_conv_weights_batch: empty or oversized table");
  pack_batch_kernel<<<static_cast<unsigned>(total_blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items);
  return check_launch("pack_batch_kernel");
}

extern "C" int dmme_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int act_dtype, void* stream) {
  DMME_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, DMME_E_BADARG, "nchw_to_nhwc: bad arguments");
  const long long total = static_cast<long long>(n) * c * h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act_dtype == DMME_BF16) nchw_to_nhwc_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n, c, h * w);
  else nchw_to_nhwc_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(src, static_cast<float*>(dst), n, c, h * w);
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int dmme_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int act_dtype, void* stream) {
  DMME_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, DMME_E_BADARG, "nhwc_to_nchw: bad arguments");
  const long long total = static_cast<long long>(n) * c * h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act_dtype == DMME_BF16) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), dst, n, c, h * w);
  else nhwc_to_nchw_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const float*>(src), dst, n, c, h * w);
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int dmme_upsample2x_nhwc(const void* src, void* dst, int n, int h, int w, int c, int act_dtype, void* stream) {
  DMME_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, DMME_E_BADARG, "upsample2x: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t le = cudaSuccess;
  if (act_dtype == DMME_BF16) {
    if (c % 8 == 0) {
      const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
      le = launch_pdl(upsample2x_kernel<__nv_bfloat16, 8>, dim3(grid_for(total, 256)), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), n, h, w, c);
    } else {
      const long long total = static_cast<long long>(n) * 4 * h * w * c;
      le = launch_pdl(upsample2x_kernel<__nv_bfloat16, 1>, dim3(grid_for(total, 256)), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), n, h, w, c);
    }
  } else {
    if (c % 4 == 0) {
      const long long total = static_cast<long long>(n) * 4 * h * w * (c / 4);
      le = launch_pdl(upsample2x_kernel<float, 4>, dim3(grid_for(total, 256)), dim3(256), 0, st, static_cast<const float*>(src), static_cast<float*>(dst), n, h, w, c);
    } else {
      const long long total = static_cast<long long>(n) * 4 * h * w * c;
      le = launch_pdl(upsample2x_kernel<float, 1>, dim3(grid_for(total, 256)), dim3(256), 0, st, static_cast<const float*>(src), static_cast<float*>(dst), n, h, w, c);
    }
  }
  return check_launch_err(le, "upsample2x_kernel");
}
