// Backward kernels of the UNet hot path, shape-agnostic CUDA-core versions (fp32 math, fp32 or bf16 storage).
//
// The reference gets its backward pass from autograd through ATen (SURVEY par. 3.3: loss.backward() after
// DDPM.training_step diffusion_models/ddpm.py:53-81 / IDDPM.training_step diffusion_models/iddpm.py:62-116).
// Here every backward op is an explicit kernel:
//   conv dgrad        = the forward convolution kernels run on grad_out with flipped/transposed weights
//                       (dmme_pack_conv_weight_dgrad; stride-2 convs via zero-dilated gather, conv_generic.cu)
//   conv wgrad (+bias) = conv_wgrad_kernel: [cout] x [K+1] products reduced over pixel slices, deterministic
//                       two-stage reduction (partials -> OIHW gradients)
//   GroupNorm(+scale/shift)+SiLU+mask backward = gn_bwd_kernel (+ gn_bwd_finalize_kernel for the parameter grads)
//   attention backward = strided batched products (gemm_strided_kernel) + softmax / dS row kernels
//   timestep-MLP backward = the same strided products + silu_bwd_kernel
//   glue: pixel sums (temb gradient), 2x2 sum pooling (nearest-upsample backward), in-place adds, MSE loss.
// Tensor-core versions of the dominant products live in conv_wgrad_tc.cu; everything here is the any-shape path
// (the reference's test fixture trains a UNet with 4/8/16/32 channels and 2 groups, tests/test_ddpm.py:7-23).
#include <mma.h>

#include "common.cuh"

namespace dmme {

bool attn_tensor_core_multi_head_enabled();
bool attn_tc_mh_supported(int act_dtype, int heads, int L, int dh, int row_stride, int head_stride, long long batch_stride,
                          int v_transposed, const void* q, const void* k, const void* v, const void* out);
int attn_tc_mh_forward(const void* qkv, int n, int heads, int L, int dh, float scale, int swap, void* out, float* p_out,
                       cudaStream_t stream);
bool attn_mma_supported(int act_dtype, int heads, int L, int dh, int row_stride, int head_stride, long long batch_stride,
                        int v_transposed, const void* q, const void* k, const void* v, const void* out);
int attn_mma_forward(const void* q, const void* k, const void* v, long long batch_stride, int row_stride, int head_stride,
                     int n, int heads, int L, int dh, float scale, int swap, void* out, float* p_out, cudaStream_t stream);

__device__ __forceinline__ float ld_any(const void* p, long long i, int dtype) {
  return dtype == DMME_F32 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, long long i, int dtype, float v) {
  if (dtype == DMME_F32) static_cast<float*>(p)[i] = v;
  else static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = 1.0f / (1.0f + expf(-z));
  return s * (1.0f + z * (1.0f - s));
}
// bf16 paths: sigmoid(z) = 0.5 + 0.5 tanh(z / 2) with tanh.approx (one MUFU, rel. error 2^-11, far below the bf16 rounding
// of the gradient it multiplies).  exp + full-precision division cost ~25 instructions per call and made the GroupNorm
// backward slab kernel -- which evaluates it twice per element at 8 warps per SM -- instruction-bound (210 us for a 100 MB
// layer).
__device__ __forceinline__ float dsilu_fast(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  const float s = fmaf(0.5f, t, 0.5f);
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// ------------------------------------------------------------------------------------------------
// strided batched product  C[b](i,j) = alpha * sum_k A[b](i,k) B[b](k,j)  (+ C[b](i,j) when accumulate)
// b = bo * heads + h; every operand has its own (outer-batch, head, row, column) element strides.
// ------------------------------------------------------------------------------------------------
struct GemmOp { const void* p; int dtype; long long s_bo, s_h, s_r, s_c; };
struct GemmParams {
  GemmOp a, b;
  void* c; int c_dtype; long long c_bo, c_h, c_r, c_c;
  int M, N, K, heads;
  float alpha; int accumulate;
  int bf16_mma;  // operands may be rounded to bf16 and multiplied on the tensor cores (bf16 training mode only)
};

constexpr int SG_M = 64, SG_N = 64, SG_K = 16;

__global__ void __launch_bounds__(256) gemm_strided_kernel(const GemmParams p) {
  __shared__ float As[SG_K][SG_M + 4];
  __shared__ float Bs[SG_K][SG_N + 4];
  const int tid = threadIdx.x;
  const int bo = blockIdx.z / p.heads, h = blockIdx.z - bo * p.heads;
  const long long a0 = bo * p.a.s_bo + h * p.a.s_h, b0 = bo * p.b.s_bo + h * p.b.s_h;
  const long long c0 = bo * p.c_bo + h * p.c_h;
  const int m0 = blockIdx.x * SG_M, n0 = blockIdx.y * SG_N;
  const bool a_kfast = p.a.s_c == 1;   // k contiguous: consecutive threads walk k
  const bool b_jfast = p.b.s_c == 1;   // j contiguous: consecutive threads walk j
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += SG_K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk, ii;
      if (a_kfast) { kk = tid & 15; ii = (tid >> 4) + 16 * j; }
      else { ii = tid & 63; kk = (tid >> 6) + 4 * j; }
      const int i = m0 + ii, k = k0 + kk;
      As[kk][ii] = (i < p.M && k < p.K) ? ld_any(p.a.p, a0 + i * p.a.s_r + k * p.a.s_c, p.a.dtype) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk, jj;
      if (b_jfast) { jj = tid & 63; kk = (tid >> 6) + 4 * j; }
      else { kk = tid & 15; jj = (tid >> 4) + 16 * j; }
      const int jn = n0 + jj, k = k0 + kk;
      Bs[kk][jj] = (jn < p.N && k < p.K) ? ld_any(p.b.p, b0 + k * p.b.s_r + jn * p.b.s_c, p.b.dtype) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cn = n0 + tx * 4 + j;
      if (cn >= p.N) continue;
      const long long o = c0 + r * p.c_r + cn * p.c_c;
      float v = p.alpha * acc[i][j];
      if (p.accumulate) v += ld_any(p.c, o, p.c_dtype);
      st_any(p.c, o, p.c_dtype, v);
    }
  }
}

// The same product with bf16 operands on the tensor cores (mma.sync through the wmma API, fp32 accumulate): the
// any-stride attention products of the bf16 training path (Q K^T, P V and the four backward products per site) were
// FFMA-bound at ~17 TFLOP/s, 28% of a training step.  Operands are converted to bf16 while they are staged in shared
// memory (fp32 P / dS buffers included: P is rounded to bf16 in the forward kernels as well); fp32 mode never takes
// this path.  64 x 64 tile, K step 32, 8 warps x (16 x 32) accumulators.
constexpr int MG_K = 32;

__global__ void __launch_bounds__(256) gemm_strided_mma_kernel(const GemmParams p) {
  using namespace nvcuda;
  __shared__ __align__(32) __nv_bfloat16 As[SG_M][MG_K + 8];
  __shared__ __align__(32) __nv_bfloat16 Bs[MG_K][SG_N + 8];
  __shared__ __align__(32) float Cs[SG_M][SG_N + 4];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int bo = blockIdx.z / p.heads, h = blockIdx.z - bo * p.heads;
  const long long a0 = bo * p.a.s_bo + h * p.a.s_h, b0 = bo * p.b.s_bo + h * p.b.s_h;
  const long long c0 = bo * p.c_bo + h * p.c_h;
  const int m0 = blockIdx.x * SG_M, n0 = blockIdx.y * SG_N;
  const bool a_kfast = p.a.s_c == 1;
  const bool b_jfast = p.b.s_c == 1;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2];
  wmma::fill_fragment(acc[0], 0.f);
  wmma::fill_fragment(acc[1], 0.f);

  // vector path: one 16-byte (bf16) or two 16-byte (fp32) loads of eight elements along the operand's contiguous
  // dimension per thread and k-step instead of eight scalar loads (the scalar staging loop made the kernel load-bound:
  // 8.3 ms of a 33 ms training step); needs 8-element alignment of every offset, else the scalar path below
  const bool a_mfast = p.a.s_r == 1;
  const bool b_kfast = p.b.s_r == 1;
  auto aligned8 = [](const GemmOp& o, long long off0, bool row_contig) {
    const long long other = row_contig ? o.s_c : o.s_r;
    const int esz = o.dtype == DMME_F32 ? 4 : 2;
    return (other % 8 == 0) && (off0 % 8 == 0) && ((reinterpret_cast<uintptr_t>(o.p) % 16) == 0) &&
           (o.s_bo % 8 == 0) && (o.s_h % 8 == 0) && esz > 0;
  };
  const bool a_vec = (a_kfast || a_mfast) && p.M % 8 == 0 && p.K % 8 == 0 && aligned8(p.a, a0, a_mfast);
  const bool b_vec = (b_jfast || b_kfast) && p.N % 8 == 0 && p.K % 8 == 0 && aligned8(p.b, b0, b_kfast);
  auto load8 = [](const GemmOp& o, long long idx, float (&f)[8]) {
    if (o.dtype == DMME_F32) {
      const float4 x0 = *reinterpret_cast<const float4*>(static_cast<const float*>(o.p) + idx);
      const float4 x1 = *reinterpret_cast<const float4*>(static_cast<const float*>(o.p) + idx + 4);
      f[0] = x0.x; f[1] = x0.y; f[2] = x0.z; f[3] = x0.w; f[4] = x1.x; f[5] = x1.y; f[6] = x1.z; f[7] = x1.w;
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(o.p) + idx);
      unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
      unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
    }
  };

  for (int k0 = 0; k0 < p.K; k0 += MG_K) {
    if (a_vec) {
      float f[8];
      if (a_kfast) {  // eight consecutive k of one row
        const int ii = tid >> 2, kk = (tid & 3) << 3;
        const int i = m0 + ii, k = k0 + kk;
        if (i < p.M && k < p.K) load8(p.a, a0 + i * p.a.s_r + k, f);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(&As[ii][kk]) = u;
      } else {  // eight consecutive rows of one k
        const int kk = tid >> 3, ii = (tid & 7) << 3;
        const int i = m0 + ii, k = k0 + kk;
        if (i < p.M && k < p.K) load8(p.a, a0 + i + k * p.a.s_c, f);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) As[ii + e][kk] = __float2bfloat16_rn(f[e]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int kk, ii;
        if (a_kfast) { kk = tid & 31; ii = (tid >> 5) + 8 * j; }
        else { ii = tid & 63; kk = (tid >> 6) + 4 * j; }
        const int i = m0 + ii, k = k0 + kk;
        As[ii][kk] = __float2bfloat16_rn((i < p.M && k < p.K) ? ld_any(p.a.p, a0 + i * p.a.s_r + k * p.a.s_c, p.a.dtype) : 0.f);
      }
    }
    if (b_vec) {
      float f[8];
      if (b_jfast) {  // eight consecutive columns of one k
        const int kk = tid >> 3, jj = (tid & 7) << 3;
        const int jn = n0 + jj, k = k0 + kk;
        if (jn < p.N && k < p.K) load8(p.b, b0 + k * p.b.s_r + jn, f);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(&Bs[kk][jj]) = u;
      } else {  // eight consecutive k of one column
        const int jj = tid >> 2, kk = (tid & 3) << 3;
        const int jn = n0 + jj, k = k0 + kk;
        if (jn < p.N && k < p.K) load8(p.b, b0 + k + jn * p.b.s_c, f);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) Bs[kk + e][jj] = __float2bfloat16_rn(f[e]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int kk, jj;
        if (b_jfast) { jj = tid & 63; kk = (tid >> 6) + 4 * j; }
        else { kk = tid & 31; jj = (tid >> 5) + 8 * j; }
        const int jn = n0 + jj, k = k0 + kk;
        Bs[kk][jj] = __float2bfloat16_rn((jn < p.N && k < p.K) ? ld_any(p.b.p, b0 + k * p.b.s_r + jn * p.b.s_c, p.b.dtype) : 0.f);
      }
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < MG_K; ks += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fa;
      wmma::load_matrix_sync(fa, &As[wm][ks], MG_K + 8);
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;
        wmma::load_matrix_sync(fb, &Bs[ks][wn + 16 * f], SG_N + 8);
        wmma::mma_sync(acc[f], fa, fb, acc[f]);
      }
    }
    __syncthreads();
  }
  wmma::store_matrix_sync(&Cs[wm][wn], acc[0], SG_N + 4, wmma::mem_row_major);
  wmma::store_matrix_sync(&Cs[wm][wn + 16], acc[1], SG_N + 4, wmma::mem_row_major);
  __syncthreads();
  // consecutive threads walk whichever output index is contiguous in memory
  const bool c_jfast = p.c_c == 1;
  for (int e = tid; e < SG_M * SG_N; e += 256) {
    const int ii = c_jfast ? e >> 6 : e & 63, jj = c_jfast ? e & 63 : e >> 6;
    const int r = m0 + ii, cn = n0 + jj;
    if (r >= p.M || cn >= p.N) continue;
    const long long o = c0 + r * p.c_r + cn * p.c_c;
    float v = p.alpha * Cs[ii][jj];
    if (p.accumulate) v += ld_any(p.c, o, p.c_dtype);
    st_any(p.c, o, p.c_dtype, v);
  }
}

static int launch_gemm(const GemmParams& p, int batches, cudaStream_t st) {
  dim3 grid(ceil_div(p.M, SG_M), ceil_div(p.N, SG_N), batches);
  if (p.bf16_mma) {
    gemm_strided_mma_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("gemm_strided_mma_kernel");
  }
  gemm_strided_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("gemm_strided_kernel");
}

// plain row-major helper: C[M][N] (ldc) = alpha * op(A) op(B); fp32 everywhere
static int gemm_f32(const float* a, long long a_r, long long a_c, const float* b, long long b_r, long long b_c, float* c,
                    long long ldc, int M, int N, int K, float alpha, int accumulate, cudaStream_t st, int bf16_mma = 0) {
  GemmParams p;
  p.a = {a, DMME_F32, 0, 0, a_r, a_c};
  p.b = {b, DMME_F32, 0, 0, b_r, b_c};
  p.c = c; p.c_dtype = DMME_F32; p.c_bo = 0; p.c_h = 0; p.c_r = ldc; p.c_c = 1;
  p.M = M; p.N = N; p.K = K; p.heads = 1; p.alpha = alpha; p.accumulate = accumulate;
  p.bf16_mma = bf16_mma;
  return launch_gemm(p, 1, st);
}

// ------------------------------------------------------------------------------------------------
// small elementwise / reduction helpers
// ------------------------------------------------------------------------------------------------
static int grid_1d(long long total, int threads) {
  long long b = ceil_div_ll(total, threads);
  const long long cap = 148LL * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// out[c] (+)= sum_r in[r][c]  (column sums of a row-major fp32 matrix); one thread per column
__global__ void colsum_kernel(const float* __restrict__ in, int rows, int cols, long long ld, float* __restrict__ out,
                              int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += in[r * ld + c];
  out[c] = accumulate ? out[c] + s : s;
}

// g[i] *= silu'(z[i])
__global__ void silu_bwd_kernel(float* __restrict__ g, const float* __restrict__ z, long long numel) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    g[i] *= dsilu_f(z[i]);
}

// sinusoidal embedding rows [sin(t f) | cos(t f)] (models/ddpm.py:338-349)
__global__ void sin_embed_kernel(const int64_t* __restrict__ t, const float* __restrict__ freq, int rows, int half,
                                 float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * half) return;
  const int r = i / (2 * half), j = i - r * 2 * half;
  const float a = static_cast<float>(t[r]) * freq[j < half ? j : j - half];
  out[i] = j < half ? sinf(a) : cosf(a);
}

// dst = a + b (activation dtype; dst may alias a or b)
template <typename T>
__global__ void add_kernel(T* dst, const T* a, const T* b, long long numel) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    st_act<T>(dst + i, ld_act<T>(a + i) + ld_act<T>(b + i));
}

// out[n][c] = sum over pixels of g[n][px][c]  (gradient of the broadcast timestep-embedding add, models/ddpm.py:129)
// grid (ceil(c / 32), n), block (32, 8): threadIdx.x walks channels (coalesced), threadIdx.y strides pixels
template <typename T>
__global__ void pixel_sum_kernel(const T* __restrict__ g, int hw, int c, float* __restrict__ out, long long out_ld) {
  __shared__ float red[8][33];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  const int n = blockIdx.y;
  float s = 0.f;
  if (ch < c)
    for (int px = threadIdx.y; px < hw; px += 8) s += ld_act<T>(g + (static_cast<long long>(n) * hw + px) * c + ch);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    out[n * out_ld + ch] = t;
  }
}

// backward of nearest x2 upsampling (models/ddpm.py:161): out[n][y][x][c] = sum of the 2x2 block of g
template <typename T>
__global__ void pool2x_sum_kernel(const T* __restrict__ g, T* __restrict__ out, int n, int h, int w, int c) {
  const long long total = static_cast<long long>(n) * h * w * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    long long pix = i / c;
    const int x = static_cast<int>(pix % w);
    pix /= w;
    const int y = static_cast<int>(pix % h);
    const long long ni = pix / h;
    const T* s = g + ((ni * 2 * h + 2 * y) * (2 * w) + 2 * x) * c + ch;
    const float v = (ld_act<T>(s) + ld_act<T>(s + c)) + (ld_act<T>(s + 2LL * w * c) + ld_act<T>(s + 2LL * w * c + c));
    st_act<T>(out + i, v);
  }
}

// zero-dilation x2 (NHWC): dst[n][2y][2x] = src[n][y][x], all other pixels 0.  The gradient of a stride-2 conv
// (models/ddpm.py:147) w.r.t. its input / weights is the stride-1 data / weight gradient of this dilated grad_out, so
// both run on the stride-1 tensor-core kernels.
template <typename T, int VEC>
__global__ void dilate2x_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int h, int w, int c) {
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
    T* d = dst + i * VEC;
    if (((x | y) & 1) == 0) {
      const T* sp = src + ((ni * h + (y >> 1)) * w + (x >> 1)) * c + v * VEC;
      if (VEC * sizeof(T) == 16) *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(sp));
      else *d = *sp;
    } else {
      if (VEC * sizeof(T) == 16) *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
      else st_act<T>(d, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// convolution weight gradient
//   dW[co][k] = sum over output pixels of g[pix][co] * in[pix, k],  k = tap * (c0 + c1) + ci | residual channel | 1 (bias)
// grid (k tiles, cout tiles, pixel slices); each CTA writes its 64x64 partial tile; wgrad_reduce_kernel sums the
// slices in a fixed order (deterministic) and scatters into the OIHW / [cout][rc] / [cout] gradient tensors.
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  const void* g;  // grad_out: NHWC act dtype, or NCHW fp32 when g_nchw
  int g_nchw;
  const void* src0; const void* src1; int c0, c1;
  const void* res0; const void* res1; int rc0, rc1;
  int n, h_in, w_in, ho, wo, ksize, stride, upsample, in_nchw;
  int cout, kp;      // kp = ksize^2 (c0+c1) + rc0 + rc1 + 1
  float* partial;    // [slices][cout][kp]
  long long pix_per_slice;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradParams p) {
  __shared__ float Gs[SG_K][SG_M + 4];  // [pixel][cout]
  __shared__ float Xs[SG_K][SG_N + 4];  // [pixel][k]
  const int tid = threadIdx.x;
  const int ctot = p.c0 + p.c1;
  const int kconv = p.ksize * p.ksize * ctot;
  const int kres = kconv + p.rc0 + p.rc1;
  const int k0 = blockIdx.x * SG_N, co0 = blockIdx.y * SG_M;
  const long long mtot = static_cast<long long>(p.n) * p.ho * p.wo;
  const long long pbeg = blockIdx.z * p.pix_per_slice;
  const long long pend = pbeg + p.pix_per_slice < mtot ? pbeg + p.pix_per_slice : mtot;
  const int pad = p.ksize / 2;
  const int hin_eff = p.upsample ? 2 * p.h_in : p.h_in, win_eff = p.upsample ? 2 * p.w_in : p.w_in;

  // this thread's fixed k column and cout column in the loads
  const int lk = tid & 63, lp = tid >> 6;
  const int k = k0 + lk;
  int kind = 3, tap = 0, ci = 0;  // 0 conv, 1 residual, 2 bias, 3 out of range
  if (k < kconv) { kind = 0; tap = k / ctot; ci = k - tap * ctot; }
  else if (k < kres) { kind = 1; ci = k - kconv; }
  else if (k == kres) kind = 2;
  const int r = tap / p.ksize, s = tap - r * p.ksize;
  const int co_l = co0 + lk;

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pp0 = pbeg; pp0 < pend; pp0 += SG_K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pl = lp + 4 * j;
      const long long pix = pp0 + pl;
      float gv = 0.f, xv = 0.f;
      if (pix < pend) {
        const int x = static_cast<int>(pix % p.wo);
        const int y = static_cast<int>((pix / p.wo) % p.ho);
        const long long ni = pix / (static_cast<long long>(p.wo) * p.ho);
        if (co_l < p.cout) {
          if (p.g_nchw) gv = static_cast<const float*>(p.g)[((ni * p.cout + co_l) * p.ho + y) * p.wo + x];
          else gv = ld_act<T>(static_cast<const T*>(p.g) + pix * p.cout + co_l);
        }
        if (kind == 0) {
          const int iy = y * p.stride + r - pad, ix = x * p.stride + s - pad;
          if (iy >= 0 && iy < hin_eff && ix >= 0 && ix < win_eff) {
            const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
            if (p.in_nchw) {
              xv = static_cast<const float*>(p.src0)[((ni * p.c0 + ci) * p.h_in + sy) * p.w_in + sx];
            } else {
              const long long pixi = (ni * p.h_in + sy) * p.w_in + sx;
              xv = ci < p.c0 ? ld_act<T>(static_cast<const T*>(p.src0) + pixi * p.c0 + ci)
                             : ld_act<T>(static_cast<const T*>(p.src1) + pixi * p.c1 + (ci - p.c0));
            }
          }
        } else if (kind == 1) {
          xv = ci < p.rc0 ? ld_act<T>(static_cast<const T*>(p.res0) + pix * p.rc0 + ci)
                          : ld_act<T>(static_cast<const T*>(p.res1) + pix * p.rc1 + (ci - p.rc0));
        } else if (kind == 2) {
          xv = 1.f;
        }
      }
      Gs[pl][lk] = gv;
      Xs[pl][lk] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Gs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* part = p.partial + static_cast<long long>(blockIdx.z) * p.cout * p.kp;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kc = k0 + tx * 4 + j;
      if (kc < p.kp) part[static_cast<long long>(co) * p.kp + kc] = acc[i][j];
    }
  }
}

// Weight gradient of the OUTPUT conv (models/ddpm.py:277: 128 -> 3 / 6 channels, 3x3): grad_out is the fp32 NCHW image-space
// gradient, a handful of channels wide, so the generic 64 x 64 tiling above wastes 90% of its tile (1.7 ms per step).
// Here thread = input channel, one CTA per image: the image's grad_out (cout planes with a zero halo) sits in shared
// memory, every activation value is loaded once (coalesced, 2 B) and feeds 9 taps x cout FMAs against broadcast
// shared-memory reads; the CTA writes its [cout][kp] partial (slice = image) for wgrad_reduce_kernel.
//   dW[co][ci][r][s] = sum over input pixels (y, x) of a[y][x][ci] * g[co][y - r + 1][x - s + 1]
constexpr int kOutWgradRowSplit = 4;  // CTAs per image (row bands): 128 CTAs of four warps left the SMs at 6 % occupancy
template <int COUT>
__global__ void __launch_bounds__(256) conv_out_wgrad_kernel(const WgradParams p) {
  extern __shared__ float gs[];  // [COUT][h + 2][w + 4]: zero halo, rows padded to a multiple of 4 floats
  const int h = p.h_in, w = p.w_in, ws = w + 4, plane = (h + 2) * ws;
  const int n = blockIdx.x;
  const int band = blockIdx.y, bands = gridDim.y;  // input rows [y_lo, y_hi) of the image
  const int y_lo = band * h / bands, y_hi = (band + 1) * h / bands;
  const int ci = threadIdx.x;  // blockDim.x == c0
  for (int i = threadIdx.x; i < COUT * plane; i += blockDim.x) gs[i] = 0.f;
  __syncthreads();
  const float* g = static_cast<const float*>(p.g) + static_cast<long long>(n) * COUT * h * w;
  for (int i = threadIdx.x; i < COUT * h * w; i += blockDim.x) {
    const int co = i / (h * w), rem = i - co * h * w, y = rem / w, x = rem - y * w;
    if (y + 2 < y_lo || y > y_hi) continue;  // only the output rows this band's taps reach
    gs[co * plane + (y + 1) * ws + x + 1] = g[i];
  }
  __syncthreads();
  float acc[9][COUT];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[t][co] = 0.f;
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(p.src0) + static_cast<long long>(n) * h * w * p.c0 + ci;
  // the activation values of kAhead pixels are requested before the first is used: one load per pixel with its use right
  // behind it left a DRAM / L2 round trip per pixel exposed (714 us per launch)
  constexpr int kAhead = 16;
  const int p_lo = y_lo * w, p_hi = y_hi * w;
  for (int p0 = p_lo; p0 < p_hi; p0 += kAhead) {
    float avs[kAhead];
#pragma unroll
    for (int i = 0; i < kAhead; ++i)
      avs[i] = p0 + i < p_hi ? __bfloat162float(a[static_cast<long long>(p0 + i) * p.c0]) : 0.f;
    // four pixels of a row share one six-column window of grad_out per (tap row, channel): two vector reads feed twelve
    // FMAs (one broadcast read per FMA, the first version, was shared-memory bound); w % 4 == 0, so a group of four never
    // leaves its row, and every accumulator still meets the pixels in ascending order (same bits)
#pragma unroll
    for (int gq = 0; gq < kAhead / 4; ++gq) {
      const int pp = p0 + 4 * gq;
      if (pp >= p_hi) break;
      const int y = pp / w, x = pp - y * w;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        // output row y - r + 1 -> halo row y - r + 2; pixel x + i, tap s -> halo column x + i + 2 - s
        const float* row = gs + (y - r + 2) * ws + x;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float4 g03 = *reinterpret_cast<const float4*>(row + co * plane);
          const float2 g45 = *reinterpret_cast<const float2*>(row + co * plane + 4);
          const float gw[6] = {g03.x, g03.y, g03.z, g03.w, g45.x, g45.y};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float av = avs[4 * gq + i];
            acc[r * 3 + 0][co] = fmaf(av, gw[i + 2], acc[r * 3 + 0][co]);
            acc[r * 3 + 1][co] = fmaf(av, gw[i + 1], acc[r * 3 + 1][co]);
            acc[r * 3 + 2][co] = fmaf(av, gw[i], acc[r * 3 + 2][co]);
          }
        }
      }
    }
  }
  float* part = p.partial + (static_cast<long long>(n) * bands + band) * COUT * p.kp;  // slice = (image, band)
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < COUT; ++co) part[static_cast<long long>(co) * p.kp + t * p.c0 + ci] = acc[t][co];
  // bias column k = 9 c0: sum of grad_out over the image (first band; the others contribute zero)
  if (threadIdx.x < COUT) {
    float t = 0.f;
    if (band == 0) {
      const float* gp = g + static_cast<long long>(threadIdx.x) * h * w;
      for (int i = 0; i < h * w; ++i) t += gp[i];
    }
    part[static_cast<long long>(threadIdx.x) * p.kp + 9 * p.c0] = t;
  }
}

static bool conv_out_wgrad_supported(const dmme_conv_desc& d) {
  return d.kernel != DMME_CONV_GENERIC && d.act_dtype == DMME_BF16 && d.out_layout == DMME_OUT_NCHW_F32 &&
         d.in_layout == DMME_IN_NHWC && d.ksize == 3 && d.stride == 1 && !d.upsample && d.c1 == 0 && d.rc0 + d.rc1 == 0 &&
         (d.cout == 3 || d.cout == 6) && d.c0 >= 32 && d.c0 <= 256 && d.c0 % 32 == 0 && d.h_in <= 32 && d.w_in <= 32 &&
         d.w_in % 4 == 0;
}

// partial [slices][cout][kp] -> dW OIHW [cout][cin][taps], dWres [cout][rc], dbias [cout]
// img_sums (tensor-core path): per-image pixel sums of grad_out [n_img][cout]; the bias gradient is their column sum, taken
// here by the thread of the bias column (in image order, as colsum_kernel did in a launch of its own: 75 launches per step)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int slices, int cout, int kp, int cin, int taps, int rc,
                                    float* __restrict__ dw, float* __restrict__ dwres, float* __restrict__ dbias,
                                    const float* __restrict__ img_sums = nullptr, int n_img = 0) {
  const long long total = static_cast<long long>(cout) * kp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i / kp), k = static_cast<int>(i - static_cast<long long>(co) * kp);
    const int kconv = taps * cin;
    float s = 0.f;
    if (k >= kconv + rc && img_sums != nullptr) {
      if (dbias) {
        // 32 loads in flight per round: with 8 the cout threads doing this were the tail of every launch (23 us instead of 13)
        for (int r0 = 0; r0 < n_img; r0 += 32) {
          float t[32];
#pragma unroll
          for (int r = 0; r < 32; ++r) t[r] = r0 + r < n_img ? img_sums[static_cast<long long>(r0 + r) * cout + co] : 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) s += t[r];  // image order, as colsum_kernel summed
        }
        dbias[co] = s;
      }
      continue;
    }
#pragma unroll 4
    for (int z = 0; z < slices; ++z) s += partial[z * total + i];
    if (k < kconv) {
      const int tap = k / cin, ci = k - tap * cin;
      dw[(static_cast<long long>(co) * cin + ci) * taps + tap] = s;
    } else if (k < kconv + rc) {
      if (dwres) dwres[static_cast<long long>(co) * rc + (k - kconv)] = s;
    } else if (dbias) {
      dbias[co] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+ scale/shift) (+ SiLU) (+ mask) backward; one CTA per (image, group)
//   forward (groupnorm.cu): u = xhat * gamma + beta, z = (1 + scale) u + shift, y = mask * silu(z)
//   gz = gout * mask * silu'(z);  A_c = sum_px gz, B_c = sum_px gz * xhat  (per image and channel, kept in `sums`)
//   gx = rstd * (ge_c * gz - m1 - xhat * m2),  ge_c = (1 + scale) gamma_c,  m1 = mean_grp(ge gz), m2 = mean_grp(ge gz xhat)
// ------------------------------------------------------------------------------------------------
struct GnBwdParams {
  const void* gout;
  const void* src0; const void* src1; int c0, c1;
  int n, hw, groups; float eps;
  const float* gamma; const float* beta;
  const float* scale; const float* shift; int ss_rows, ss_ld;
  const float* mask; int silu;
  void* gin0; void* gin1;
  const void* add0; const void* add1;
  float* sums;  // [n][C][2]: A, B
};

template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_kernel(const GnBwdParams p) {
  extern __shared__ float sh[];  // A[cpg], B[cpg], then 32 floats of reduction scratch
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const int n = blockIdx.x / p.groups, g = blockIdx.x - n * p.groups;
  const int cb = g * cpg;
  const long long E = static_cast<long long>(p.hw) * cpg;
  float* sA = sh; float* sB = sh + cpg; float* red = sh + 2 * cpg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  auto load_x = [&](long long e, int& cl) -> float {
    const long long px = e / cpg;
    cl = static_cast<int>(e - px * cpg);
    const int c = cb + cl;
    const long long pix = static_cast<long long>(n) * p.hw + px;
    return c < p.c0 ? ld_act<T>(static_cast<const T*>(p.src0) + pix * p.c0 + c)
                    : ld_act<T>(static_cast<const T*>(p.src1) + pix * p.c1 + (c - p.c0));
  };
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    return t;
  };

  // statistics, two passes like torch.native_group_norm (biased variance)
  float s = 0.f;
  int cl;
  for (long long e = tid; e < E; e += blockDim.x) s += load_x(e, cl);
  const float mean = block_sum(s) / static_cast<float>(E);
  s = 0.f;
  for (long long e = tid; e < E; e += blockDim.x) { const float d = load_x(e, cl) - mean; s += d * d; }
  const float rstd = rsqrtf(block_sum(s) / static_cast<float>(E) + p.eps);

  for (int i = tid; i < 2 * cpg; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();

  const int ssr = p.scale ? (p.ss_rows == 1 ? 0 : n) : 0;
  auto gz_of = [&](long long e, float& xhat, int& c_out) -> float {
    int c_l;
    const float x = load_x(e, c_l);
    const int c = cb + c_l;
    c_out = c;
    xhat = (x - mean) * rstd;
    float z = xhat * p.gamma[c] + p.beta[c];
    if (p.scale) z = z * (1.0f + p.scale[static_cast<long long>(ssr) * p.ss_ld + c]) + p.shift[static_cast<long long>(ssr) * p.ss_ld + c];
    const long long px = e / cpg;
    float gv = ld_act<T>(static_cast<const T*>(p.gout) + (static_cast<long long>(n) * p.hw + px) * C + c);
    if (p.mask) gv *= p.mask[static_cast<long long>(n) * C + c];
    if (p.silu) gv *= dsilu_f(z);
    return gv;
  };

  for (long long e = tid; e < E; e += blockDim.x) {
    float xhat; int c;
    const float gz = gz_of(e, xhat, c);
    atomicAdd(&sA[c - cb], gz);
    atomicAdd(&sB[c - cb], gz * xhat);
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < cpg; i += blockDim.x) {
    const int c = cb + i;
    float ge = p.gamma[c];
    if (p.scale) ge *= 1.0f + p.scale[static_cast<long long>(ssr) * p.ss_ld + c];
    s1 += ge * sA[i];
    s2 += ge * sB[i];
    p.sums[(static_cast<long long>(n) * C + c) * 2 + 0] = sA[i];
    p.sums[(static_cast<long long>(n) * C + c) * 2 + 1] = sB[i];
  }
  const float m1 = block_sum(s1) / static_cast<float>(E);
  const float m2 = block_sum(s2) / static_cast<float>(E);

  for (long long e = tid; e < E; e += blockDim.x) {
    float xhat; int c;
    const float gz = gz_of(e, xhat, c);
    float ge = p.gamma[c];
    if (p.scale) ge *= 1.0f + p.scale[static_cast<long long>(ssr) * p.ss_ld + c];
    float gx = rstd * (ge * gz - m1 - xhat * m2);
    const long long pix = static_cast<long long>(n) * p.hw + e / cpg;
    if (c < p.c0) {
      const long long o = pix * p.c0 + c;
      if (p.add0) gx += ld_act<T>(static_cast<const T*>(p.add0) + o);
      if (p.gin0) st_act<T>(static_cast<T*>(p.gin0) + o, gx);
    } else {
      const long long o = pix * p.c1 + (c - p.c0);
      if (p.add1) gx += ld_act<T>(static_cast<const T*>(p.add1) + o);
      if (p.gin1) st_act<T>(static_cast<T*>(p.gin1) + o, gx);
    }
  }
}

// Fast path of the same backward (bf16, C % 32 == 0, channels-per-group in {1..32} dividing 32, HW <= 1024): one CTA owns
// one (image, 32-channel slab).  x and grad_out of the slab are read ONCE into registers (16-byte vectors, 64 contiguous
// bytes per pixel), the statistics are recomputed from the registers exactly like the forward slab kernel, the
// per-channel sums A, B and the group means m1, m2 are reduced with shuffles + shared memory (no atomics), and the input
// gradient is written once: 2 + 2 bytes read, 2 bytes written per element.
template <int MAXV>
__global__ void __launch_bounds__(MAXV == 8 ? 512 : 256, MAXV > 4 ? 1 : 2) gn_bwd_slab_kernel(const GnBwdParams p) {
  __shared__ float red[16][64];
  __shared__ float chan[64];      // per-channel scratch: [0,32) and [32,64)
  __shared__ float grp[64];       // per-channel broadcast of group quantities
  const int C = p.c0 + p.c1;
  const int slabs = C / 32;
  const int n = blockIdx.x / slabs;
  const int slab = blockIdx.x - n * slabs;
  const int cbase = slab * 32;
  const int cpg = C / p.groups;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int chunk = tid & 3;
  const int nvec = p.hw * 4;
  const bool first = cbase < p.c0;
  const int csrc = first ? p.c0 : p.c1, coff = first ? cbase : cbase - p.c0;
  const __nv_bfloat16* xsrc = static_cast<const __nv_bfloat16*>(first ? p.src0 : p.src1) +
                              static_cast<long long>(n) * p.hw * csrc + coff + chunk * 8;
  const __nv_bfloat16* gsrc = static_cast<const __nv_bfloat16*>(p.gout) + static_cast<long long>(n) * p.hw * C + cbase + chunk * 8;

  uint4 xv[MAXV], gv[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) {
      xv[i] = __ldg(reinterpret_cast<const uint4*>(xsrc + static_cast<long long>(vi >> 2) * csrc));
      gv[i] = __ldg(reinterpret_cast<const uint4*>(gsrc + static_cast<long long>(vi >> 2) * C));
    } else {
      xv[i] = make_uint4(0, 0, 0, 0);
      gv[i] = make_uint4(0, 0, 0, 0);
    }
  }
  auto unpack8 = [](const uint4& v, float (&f)[8]) {
    unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
  };
  // block reduction of 8 per-thread values per chunk -> chan[off + chunk*8 + j] (sum over all pixels)
  auto reduce8 = [&](float (&s)[8], int off) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 4);
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
    }
    __syncthreads();
    if (lane < 4) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][off + lane * 8 + j] = s[j];
    }
    __syncthreads();
    if (tid < 32) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += red[w][off + tid];
      chan[off + tid] = t;
    }
    __syncthreads();
  };
  const float inv_cnt = 1.0f / (static_cast<float>(p.hw) * cpg);

  // ---- statistics (two register passes, biased variance) ----
  float s[8], f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    unpack8(xv[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
  }
  reduce8(s, 0);
  if (tid < 32) {
    const int g0 = (tid / cpg) * cpg;
    float t = 0.f;
    for (int j = 0; j < cpg; ++j) t += chan[g0 + j];
    grp[tid] = t * inv_cnt;
  }
  __syncthreads();
  float mean[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mean[j] = grp[chunk * 8 + j];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) {
      unpack8(xv[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[j] - mean[j]; s[j] = fmaf(d, d, s[j]); }
    }
  }
  reduce8(s, 0);
  if (tid < 32) {
    const int g0 = (tid / cpg) * cpg;
    float t = 0.f;
    for (int j = 0; j < cpg; ++j) t += chan[g0 + j];
    grp[32 + tid] = rsqrtf(t * inv_cnt + p.eps);
  }
  __syncthreads();

  // ---- per-channel coefficients: xhat = x ha + hb, z = x za + zb, ge = (1 + scale) gamma ----
  float ha[8], hb[8], za[8], zb[8], ge[8], mk[8];
  const int ssr = p.scale ? (p.ss_rows == 1 ? 0 : n) : 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cbase + chunk * 8 + j;
    const float rs = grp[32 + chunk * 8 + j];
    float gam = p.gamma[c], bet = p.beta[c];
    if (p.scale) {
      const float sc = 1.0f + p.scale[static_cast<long long>(ssr) * p.ss_ld + c];
      gam *= sc;
      bet = bet * sc + p.shift[static_cast<long long>(ssr) * p.ss_ld + c];
    }
    ha[j] = rs; hb[j] = -mean[j] * rs;
    za[j] = rs * gam; zb[j] = bet - mean[j] * rs * gam;
    ge[j] = gam;
    mk[j] = p.mask ? p.mask[static_cast<long long>(n) * C + c] : 1.f;
  }
  auto gz8 = [&](const uint4& xq, const uint4& gq, float (&xh)[8], float (&gz)[8]) {
    float xf[8], gf[8];
    unpack8(xq, xf);
    unpack8(gq, gf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = fmaf(xf[j], ha[j], hb[j]);
      float gval = gf[j] * mk[j];
      if (p.silu) gval *= dsilu_fast(fmaf(xf[j], za[j], zb[j]));
      gz[j] = gval;
    }
  };

  // ---- A_c = sum gz, B_c = sum gz xhat ----
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) {
      float xh[8], gz[8];
      gz8(xv[i], gv[i], xh, gz);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sa[j] += gz[j]; sb[j] = fmaf(gz[j], xh[j], sb[j]); }
    }
  }
  reduce8(sa, 0);
  reduce8(sb, 32);
  if (tid < 32) {
    const int c = cbase + tid;
    p.sums[(static_cast<long long>(n) * C + c) * 2 + 0] = chan[tid];
    p.sums[(static_cast<long long>(n) * C + c) * 2 + 1] = chan[32 + tid];
  }
  // group means of ge gz and ge gz xhat
  if (tid < 32) {
    const int g0 = (tid / cpg) * cpg;
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const int c = cbase + g0 + j;
      float gam = p.gamma[c];
      if (p.scale) gam *= 1.0f + p.scale[static_cast<long long>(ssr) * p.ss_ld + c];
      t1 += gam * chan[g0 + j];
      t2 += gam * chan[32 + g0 + j];
    }
    grp[tid] = t1 * inv_cnt;
    red[0][tid] = t2 * inv_cnt;
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { m1[j] = grp[chunk * 8 + j]; m2[j] = red[0][chunk * 8 + j]; }

  // ---- gx = rstd (ge gz - m1 - xhat m2) (+ addend) ----
  __nv_bfloat16* gin = static_cast<__nv_bfloat16*>(first ? p.gin0 : p.gin1);
  const __nv_bfloat16* add = static_cast<const __nv_bfloat16*>(first ? p.add0 : p.add1);
  if (gin == nullptr) return;
  const long long obase = static_cast<long long>(n) * p.hw * csrc + coff + chunk * 8;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = tid + i * blockDim.x;
    if (vi < nvec) {
      float xh[8], gz[8], o[8];
      gz8(xv[i], gv[i], xh, gz);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ha[j] * (ge[j] * gz[j] - m1[j] - xh[j] * m2[j]);
      const long long off = obase + static_cast<long long>(vi >> 2) * csrc;
      if (add) {
        float af[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(add + off)), af);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += af[j];
      }
      uint4 ov;
      ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]);
      ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(gin + off) = ov;
    }
  }
}

// parameter gradients from the per-(image, channel) sums; one thread per channel
//   dgamma_c = sum_n (1+scale_nc) B_nc, dbeta_c = sum_n (1+scale_nc) A_nc
//   dshift_nc = A_nc, dscale_nc = gamma_c B_nc + beta_c A_nc   (needs per-image scale rows)
// Block = 32 channels x 8 image lanes (warp w takes images w, w + 8, ...; four images in flight per thread); the eight
// partial sums of a channel are added in warp order (deterministic).  One thread per channel walking all n images with a
// load-use round trip each took 24 us per launch, 56 launches per training step.
constexpr int kGnFinWarps = 8;
__global__ void __launch_bounds__(kGnFinWarps * 32) gn_bwd_finalize_kernel(
    const float* __restrict__ sums, int n, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ scale, int ss_rows, int ss_ld, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ dscale, float* __restrict__ dshift, int dss_ld) {
  __shared__ float pg[kGnFinWarps][32], pb[kGnFinWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float dg = 0.f, db = 0.f;
  if (c < C) {
    const float ga = dscale ? gamma[c] : 0.f, be = dscale ? beta[c] : 0.f;
    for (int i0 = warp; i0 < n; i0 += 4 * kGnFinWarps) {
      float2 ab[4];
      float f[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kGnFinWarps;
        ab[u] = i < n ? *reinterpret_cast<const float2*>(sums + (static_cast<long long>(i) * C + c) * 2) : make_float2(0.f, 0.f);
        f[u] = (scale && i < n) ? 1.0f + scale[static_cast<long long>(ss_rows == 1 ? 0 : i) * ss_ld + c] : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kGnFinWarps;
        if (i >= n) continue;
        dg += f[u] * ab[u].y;
        db += f[u] * ab[u].x;
        if (dscale) {
          dscale[static_cast<long long>(i) * dss_ld + c] = ga * ab[u].y + be * ab[u].x;
          dshift[static_cast<long long>(i) * dss_ld + c] = ab[u].x;
        }
      }
    }
  }
  pg[warp][lane] = dg;
  pb[warp][lane] = db;
  __syncthreads();
  if (warp == 0 && c < C) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < kGnFinWarps; ++w) { tg += pg[w][lane]; tb += pb[w][lane]; }
    dgamma[c] = tg;
    dbeta[c] = tb;
  }
}

// ------------------------------------------------------------------------------------------------
// attention backward pieces (row kernels over the fp32 score workspace)
// ------------------------------------------------------------------------------------------------
// in place: row <- softmax(row); one warp per row
__global__ void softmax_rows_kernel(float* __restrict__ s, long long rows, int L) {
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* r = s + row * L;
  float m = -INFINITY;
  for (int j = lane; j < L; j += 32) m = fmaxf(m, r[j]);
  m = warp_max(m);
  float z = 0.f;
  for (int j = lane; j < L; j += 32) { const float e = expf(r[j] - m); r[j] = e; z += e; }
  z = warp_sum(z);
  const float inv = 1.0f / z;
  for (int j = lane; j < L; j += 32) r[j] *= inv;
}
// in place on dp: dS = P o (dP - sum_j P o dP) * scale
__global__ void attn_ds_kernel(const float* __restrict__ pmat, float* __restrict__ dp, long long rows, int L, float scale) {
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* pr = pmat + row * L;
  float* dr = dp + row * L;
  float d = 0.f;
  for (int j = lane; j < L; j += 32) d += pr[j] * dr[j];
  d = warp_sum(d);
  for (int j = lane; j < L; j += 32) dr[j] = pr[j] * (dr[j] - d) * scale;
}
// dense fp32 copy of the output gradient in (batch, head) order:  dst[p][l][c] = dout[b'][l][h' dh + c]
// with (b', h') = (p % n, p / n) when the forward regrouped "(b head)" as "(head b)" (models/iddpm.py:44-46), else (b, h)
template <typename T>
__global__ void attn_gather_dout_kernel(const T* __restrict__ dout, float* __restrict__ dst, int n, int heads, int L, int dh,
                                        int swap) {
  const long long total = static_cast<long long>(n) * heads * L * dh;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % dh);
    long long r = i / dh;
    const int l = static_cast<int>(r % L);
    const int pidx = static_cast<int>(r / L);
    int bb, hh;
    if (swap) { hh = pidx / n; bb = pidx - hh * n; }
    else { bb = pidx / heads; hh = pidx - bb * heads; }
    dst[i] = ld_act<T>(dout + (static_cast<long long>(bb) * L + l) * (heads * dh) + hh * dh + c);
  }
}

// inverse of attn_gather_dout_kernel: out[b'][l][h' dh + c] = src[p][l][c]
template <typename T>
__global__ void attn_scatter_out_kernel(const float* __restrict__ src, T* __restrict__ out, int n, int heads, int L, int dh,
                                        int swap) {
  const long long total = static_cast<long long>(n) * heads * L * dh;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % dh);
    long long r = i / dh;
    const int l = static_cast<int>(r % L);
    const int pidx = static_cast<int>(r / L);
    int bb, hh;
    if (swap) { hh = pidx / n; bb = pidx - hh * n; }
    else { bb = pidx / heads; hh = pidx - bb * heads; }
    st_act<T>(out + (static_cast<long long>(bb) * L + l) * (heads * dh) + hh * dh + c, src[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// L_simple: loss = mean((noise - eps)^2), d_eps = 2 (eps - noise) / numel  (equations/ddpm/losses.py:5-13)
// noise is recovered as (x_t - mean) / std like the reference does (diffusion_models/ddpm.py:79).
// Deterministic: per-CTA partial sums, then one CTA adds them in order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ eps, const float* __restrict__ noise,
                                                          long long numel, float inv_numel, float gscale,
                                                          float* __restrict__ d_eps, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = eps[i] - noise[i];
    s += d * d;
    if (d_eps) d_eps[i] = 2.0f * d * inv_numel * gscale;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sum_partials_kernel(const float* __restrict__ partial, int count, float scale, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < count; ++i) t += partial[i];
    out[0] = t * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// IDDPM hybrid / VLB loss, forward value and gradient w.r.t. the network output in one elementwise pass
// (IDDPM.training_step diffusion_models/iddpm.py:62-116, forward_model :150-164, loss_vlb / discrete_nll_loss /
// true_reverse_process / interpolate_variance equations/iddpm/losses.py:8-90, simple_loss equations/ddpm/losses.py).
// The reference's masked gathers + cat + mean equal where(t == 1, nll, kl).mean(); the mean of
// p uses eps.detach(), so L_vlb sends gradient to the variance channels only.
//   partial[2 * block + {0, 1}] = block sums of (noise - eps)^2 and of the per-element vlb term
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float normal_cdf_f(float z) { return 0.5f * (1.0f + erff(z * 0.70710678118654752440f)); }
__device__ __forceinline__ float normal_pdf_f(float z) { return 0.39894228040143267794f * expf(-0.5f * z * z); }

__global__ void __launch_bounds__(256) iddpm_loss_kernel(const float* __restrict__ model_out, const float* __restrict__ x_t,
                                                         const float* __restrict__ x_0, const int64_t* __restrict__ t,
                                                         const float* __restrict__ beta, const float* __restrict__ alpha,
                                                         const float* __restrict__ alpha_bar, int n, int c, int hw,
                                                         float g_simple, float g_vlb, float* __restrict__ d_out,
                                                         float* __restrict__ partial) {
  __shared__ float red[2][8];
  const long long chw = static_cast<long long>(c) * hw;
  const long long total = static_cast<long long>(n) * chw;
  float s_simple = 0.f, s_vlb = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long ni = i / chw, rem = i - ni * chw;
    const long long oe = ni * 2 * chw + rem, ov = oe + chw;
    const int64_t tt = t[ni];
    const float b = beta[tt], a = alpha[tt], ab = alpha_bar[tt], abp = alpha_bar[tt - 1];
    const float eps = model_out[oe], v = model_out[ov], xt = x_t[i], x0 = x_0[i];
    // learned variance (forward_model)
    const float beta_tilde = (1.0f - abp) / (1.0f - ab) * b;
    const float lb = logf(b), lbt = logf(fmaxf(beta_tilde, 1e-12f));
    const float var = expf(v * lb + (1.0f - v) * lbt);
    // L_simple on the recovered noise
    const float qm = sqrtf(ab) * x0, qs = sqrtf(1.0f - ab);
    const float noise = (xt - qm) / qs;
    const float d = eps - noise;
    s_simple += d * d;
    // L_vlb term
    const float p_mean = 1.0f / sqrtf(a) * (xt - b / sqrtf(1.0f - ab) * eps);
    const float p_std = sqrtf(var);
    float term, dterm_dvar;
    if (tt == 1) {
      const float zp = (x0 + 1.0f / 255.0f - p_mean) / p_std, zm = (x0 - 1.0f / 255.0f - p_mean) / p_std;
      const bool hi = x0 < 1.0f, lo = x0 > -1.0f;
      const float up = hi ? normal_cdf_f(zp) : 1.0f, dn = lo ? normal_cdf_f(zm) : 0.0f;
      const float prob = up - dn;
      term = -logf(fmaxf(prob, 1e-12f));
      const float dprob = (hi ? -normal_pdf_f(zp) * zp / p_std : 0.f) - (lo ? -normal_pdf_f(zm) * zm / p_std : 0.f);
      const float dsigma = prob >= 1e-12f ? -dprob / prob : 0.f;
      dterm_dvar = dsigma / (2.0f * p_std);
    } else {
      const float q_mean = sqrtf(abp) * b / (1.0f - ab) * x0 + sqrtf(a) * (1.0f - abp) / (1.0f - ab) * xt;
      const float q_std = sqrtf(beta_tilde);
      const float r = q_std / p_std, vr = r * r;
      const float u = (q_mean - p_mean) / p_std, t1 = u * u;
      term = 0.5f * (vr + t1 - 1.0f - logf(vr));
      dterm_dvar = (1.0f - vr - t1) / (2.0f * var);
    }
    s_vlb += term;
    if (d_out) {
      d_out[oe] = g_simple * 2.0f * d;
      d_out[ov] = g_vlb * dterm_dvar * var * (lb - lbt);
    }
  }
  s_simple = warp_sum(s_simple);
  s_vlb = warp_sum(s_vlb);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_simple; red[1][threadIdx.x >> 5] = s_vlb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    partial[2 * blockIdx.x] = t0;
    partial[2 * blockIdx.x + 1] = t1;
  }
}
// out[0] = w_simple * sum(simple) + w_vlb * sum(vlb), out[1] = mean simple, out[2] = mean vlb
__global__ void iddpm_loss_finish_kernel(const float* __restrict__ partial, int count, float inv_numel, float w_simple,
                                         float w_vlb, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < count; ++i) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a *= inv_numel; b *= inv_numel;
    out[0] = w_simple * a + w_vlb * b;
    out[1] = a;
    out[2] = b;
  }
}

bool conv_wgrad_tc_supported(const dmme_conv_desc& d);
void conv_wgrad_tc_geometry(const dmme_conv_desc& d, int& items, int& chunks_total, int& chunks_per_slice, int& slices);
int conv_wgrad_tc_partials(const dmme_conv_desc& d, const void* grad_out, float* partial, int& slices, cudaStream_t stream);

int launch_linear(const float* in, const int64_t* t, const float* freq, int rows, int in_dim, const float* w,
                  const float* b, int out_dim, int act, float* out, cudaStream_t st, const char* what);

}  // namespace dmme

using namespace dmme;

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int dmme_gemm_strided(const void* a, int a_dtype, long long a_bo, long long a_h, long long a_r, long long a_c,
                                 const void* b, int b_dtype, long long b_bo, long long b_h, long long b_r, long long b_c,
                                 void* c, int c_dtype, long long c_bo, long long c_h, long long c_r, long long c_c, int M,
                                 int N, int K, int outer, int heads, float alpha, int accumulate, void* stream) {
  DMME_REQUIRE(a && b && c, DMME_E_BADARG, "gemm_strided: null pointer");
  DMME_REQUIRE(M > 0 && N > 0 && K > 0 && outer > 0 && heads > 0, DMME_E_BADARG, "gemm_strided: bad sizes");
  DMME_REQUIRE(static_cast<long long>(outer) * heads <= 65535, DMME_E_SHAPE, "gemm_strided: more than 65535 batches");
  GemmParams p;
  p.a = {a, a_dtype, a_bo, a_h, a_r, a_c};
  p.b = {b, b_dtype, b_bo, b_h, b_r, b_c};
  p.c = c; p.c_dtype = c_dtype; p.c_bo = c_bo; p.c_h = c_h; p.c_r = c_r; p.c_c = c_c;
  p.M = M; p.N = N; p.K = K; p.heads = heads; p.alpha = alpha; p.accumulate = accumulate;
  p.bf16_mma = 0;
  return launch_gemm(p, outer * heads, static_cast<cudaStream_t>(stream));
}

extern "C" int dmme_add(void* dst, const void* a, const void* b, long long numel, int act_dtype, void* stream) {
  DMME_REQUIRE(dst && a && b && numel > 0, DMME_E_BADARG, "add: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act_dtype == DMME_BF16)
    add_kernel<__nv_bfloat16><<<grid_1d(numel, 256), 256, 0, st>>>(static_cast<__nv_bfloat16*>(dst), static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), numel);
  else
    add_kernel<float><<<grid_1d(numel, 256), 256, 0, st>>>(static_cast<float*>(dst), static_cast<const float*>(a), static_cast<const float*>(b), numel);
  return check_launch("add_kernel");
}

extern "C" int dmme_pixel_sum(const void* g, int n, int hw, int c, float* out, long long out_ld, int act_dtype, void* stream) {
  DMME_REQUIRE(g && out && n > 0 && hw > 0 && c > 0, DMME_E_BADARG, "pixel_sum: bad arguments");
  dim3 grid(ceil_div(c, 32), n), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act_dtype == DMME_BF16) pixel_sum_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(g), hw, c, out, out_ld);
  else pixel_sum_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(g), hw, c, out, out_ld);
  return check_launch("pixel_sum_kernel");
}

extern "C" int dmme_pool2x_sum_nhwc(const void* g, void* out, int n, int h, int w, int c, int act_dtype, void* stream) {
  DMME_REQUIRE(g && out && n > 0 && h > 0 && w > 0 && c > 0, DMME_E_BADARG, "pool2x_sum: bad arguments");
  const long long total = static_cast<long long>(n) * h * w * c;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act_dtype == DMME_BF16)
    pool2x_sum_kernel<__nv_bfloat16><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(out), n, h, w, c);
  else
    pool2x_sum_kernel<float><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const float*>(g), static_cast<float*>(out), n, h, w, c);
  return check_launch("pool2x_sum_kernel");
}

extern "C" int dmme_dilate2x_nhwc(const void* src, void* dst, int n, int h, int w, int c, int act_dtype, void* stream) {
  DMME_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0, DMME_E_BADARG, "dilate2x: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act_dtype == DMME_BF16 && c % 8 == 0) {
    const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
    dilate2x_kernel<__nv_bfloat16, 8><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), n, h, w, c);
  } else if (act_dtype == DMME_BF16) {
    const long long total = static_cast<long long>(n) * 4 * h * w * c;
    dilate2x_kernel<__nv_bfloat16, 1><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), n, h, w, c);
  } else {
    const long long total = static_cast<long long>(n) * 4 * h * w * c;
    dilate2x_kernel<float, 1><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const float*>(src), static_cast<float*>(dst), n, h, w, c);
  }
  return check_launch("dilate2x_kernel");
}

// ---- conv wgrad ----------------------------------------------------------------------------------
static void wgrad_geometry(const dmme_conv_desc& d, int& ho, int& wo, int& kp, int& slices, long long& pps) {
  const int hin_eff = d.upsample ? 2 * d.h_in : d.h_in, win_eff = d.upsample ? 2 * d.w_in : d.w_in;
  const int pad = d.ksize / 2;
  ho = (hin_eff + 2 * pad - d.ksize) / d.stride + 1;
  wo = (win_eff + 2 * pad - d.ksize) / d.stride + 1;
  kp = d.ksize * d.ksize * (d.c0 + d.c1) + d.rc0 + d.rc1 + 1;
  const long long mtot = static_cast<long long>(d.n) * ho * wo;
  const long long tiles = static_cast<long long>(ceil_div(kp, SG_N)) * ceil_div(d.cout, SG_M);
  long long want = ceil_div_ll(148 * 6, tiles);
  const long long max_slices = ceil_div_ll(mtot, 128);
  if (want > max_slices) want = max_slices;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  pps = ceil_div_ll(ceil_div_ll(mtot, want), SG_K) * SG_K;
  slices = static_cast<int>(ceil_div_ll(mtot, pps));
}

static bool wgrad_use_tc(const dmme_conv_desc& d) { return d.kernel != DMME_CONV_GENERIC && conv_wgrad_tc_supported(d); }

extern "C" int dmme_conv2d_wgrad_uses_tc(const dmme_conv_desc* d) { return d && wgrad_use_tc(*d) ? 1 : 0; }

extern "C" long long dmme_conv2d_wgrad_workspace(const dmme_conv_desc* d) {
  if (!d || d->n <= 0 || d->cout <= 0) return 0;
  int ho, wo, kp, slices; long long pps;
  wgrad_geometry(*d, ho, wo, kp, slices, pps);
  if (wgrad_use_tc(*d)) {
    int items, chunks_total, cps, tslices;
    conv_wgrad_tc_geometry(*d, items, chunks_total, cps, tslices);
    // partials + the per-image pixel sums of grad_out for the bias gradient
    return (static_cast<long long>(tslices) * d->cout * kp + static_cast<long long>(d->n) * d->cout) * sizeof(float);
  }
  if (conv_out_wgrad_supported(*d))  // one slice per (image, row band)
    slices = slices > d->n * kOutWgradRowSplit ? slices : d->n * kOutWgradRowSplit;
  return static_cast<long long>(slices) * d->cout * kp * sizeof(float);
}

extern "C" int dmme_conv2d_wgrad(const dmme_conv_desc* d, const void* grad_out, float* dweight, float* dweight_res,
                                 float* dbias, void* workspace, long long workspace_bytes, void* stream) {
  DMME_REQUIRE(d && grad_out && dweight && workspace, DMME_E_BADARG, "conv2d_wgrad: null pointer");
  DMME_REQUIRE(d->src0 && d->n > 0 && d->c0 > 0 && d->cout > 0, DMME_E_BADARG, "conv2d_wgrad: bad descriptor");
  DMME_REQUIRE(d->ksize == 1 || d->ksize == 3, DMME_E_SHAPE, "conv2d_wgrad: ksize must be 1 or 3");
  DMME_REQUIRE(d->out_layout != DMME_OUT_QKV, DMME_E_UNSUPPORTED, "conv2d_wgrad: QKV-split outputs have no backward; use the NHWC layout in training");
  DMME_REQUIRE((d->rc0 + d->rc1 == 0) || dweight_res, DMME_E_BADARG, "conv2d_wgrad: fused residual needs dweight_res");
  int ho, wo, kp, slices; long long pps;
  wgrad_geometry(*d, ho, wo, kp, slices, pps);
  DMME_REQUIRE(workspace_bytes >= dmme_conv2d_wgrad_workspace(d), DMME_E_BADARG,
               "conv2d_wgrad: workspace too small (%lld bytes)", workspace_bytes);
  if (wgrad_use_tc(*d)) {
    // tensor-core path: MN-major tcgen05 products into per-slice partials, the same deterministic reduction, and the
    // bias gradient as pixel sums per image followed by a column sum over the images
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partial = static_cast<float*>(workspace);
    int tslices = 0;
    int rc = conv_wgrad_tc_partials(*d, grad_out, partial, tslices, st);
    if (rc) return rc;
    const long long total = static_cast<long long>(d->cout) * kp;
    float* img_sums = partial + static_cast<long long>(tslices) * d->cout * kp;
    if (dbias) {
      dim3 grid(ceil_div(d->cout, 32), d->n), block(32, 8);
      pixel_sum_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(grad_out), ho * wo, d->cout, img_sums, d->cout);
      if ((rc = check_launch("pixel_sum_kernel"))) return rc;
    }
    wgrad_reduce_kernel<<<grid_1d(total, 256), 256, 0, st>>>(partial, tslices, d->cout, kp, d->c0 + d->c1, d->ksize * d->ksize,
                                                            d->rc0 + d->rc1, dweight, dweight_res, dbias, img_sums, d->n);
    if ((rc = check_launch("wgrad_reduce_kernel"))) return rc;
    return 0;
  }
  WgradParams p;
  p.g = grad_out; p.g_nchw = d->out_layout == DMME_OUT_NCHW_F32;
  p.src0 = d->src0; p.src1 = d->src1; p.c0 = d->c0; p.c1 = d->c1;
  p.res0 = d->res0; p.res1 = d->res1; p.rc0 = d->rc0; p.rc1 = d->rc1;
  p.n = d->n; p.h_in = d->h_in; p.w_in = d->w_in; p.ho = ho; p.wo = wo;
  p.ksize = d->ksize; p.stride = d->stride; p.upsample = d->upsample; p.in_nchw = d->in_layout == DMME_IN_NCHW_F32;
  p.cout = d->cout; p.kp = kp; p.partial = static_cast<float*>(workspace); p.pix_per_slice = pps;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (conv_out_wgrad_supported(*d)) {
    const int smem = d->cout * (d->h_in + 2) * (d->w_in + 4) * static_cast<int>(sizeof(float));
    const int bands = d->h_in >= kOutWgradRowSplit ? kOutWgradRowSplit : 1;
    const dim3 ogrid(d->n, bands);
    if (d->cout == 3) conv_out_wgrad_kernel<3><<<ogrid, d->c0, smem, st>>>(p);
    else conv_out_wgrad_kernel<6><<<ogrid, d->c0, smem, st>>>(p);
    int rc = check_launch("conv_out_wgrad_kernel");
    if (rc) return rc;
    const long long total = static_cast<long long>(d->cout) * kp;
    wgrad_reduce_kernel<<<grid_1d(total, 256), 256, 0, st>>>(p.partial, d->n * bands, d->cout, kp, d->c0, 9, 0, dweight, nullptr, dbias);
    return check_launch("wgrad_reduce_kernel");
  }
  dim3 grid(ceil_div(kp, SG_N), ceil_div(d->cout, SG_M), slices);
  if (d->act_dtype == DMME_BF16) conv_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else conv_wgrad_kernel<float><<<grid, 256, 0, st>>>(p);
  int rc = check_launch("conv_wgrad_kernel");
  if (rc) return rc;
  const long long total = static_cast<long long>(d->cout) * kp;
  wgrad_reduce_kernel<<<grid_1d(total, 256), 256, 0, st>>>(p.partial, slices, d->cout, kp, d->c0 + d->c1, d->ksize * d->ksize,
                                                          d->rc0 + d->rc1, dweight, dweight_res, dbias);
  return check_launch("wgrad_reduce_kernel");
}

// ---- GroupNorm backward ----------------------------------------------------------------------------
extern "C" int dmme_groupnorm_bwd(const void* grad_out, const void* src0, const void* src1, int c0, int c1, int n, int hw,
                                  int groups, float eps, const float* gamma, const float* beta, const float* scale,
                                  const float* shift, int ss_rows, int ss_ld, const float* chan_mask, int apply_silu,
                                  void* gin0, void* gin1, const void* add0, const void* add1, float* dgamma, float* dbeta,
                                  float* dscale, float* dshift, int dss_ld, float* sums, int act_dtype, void* stream) {
  DMME_REQUIRE(grad_out && src0 && gamma && beta && sums && dgamma && dbeta, DMME_E_BADARG, "groupnorm_bwd: null pointer");
  DMME_REQUIRE(n > 0 && hw > 0 && c0 > 0 && c1 >= 0 && groups > 0, DMME_E_BADARG, "groupnorm_bwd: bad sizes");
  DMME_REQUIRE(c1 == 0 || src1, DMME_E_BADARG, "groupnorm_bwd: c1 > 0 but src1 is null");
  const int C = c0 + c1;
  DMME_REQUIRE(C % groups == 0, DMME_E_SHAPE, "groupnorm_bwd: C=%d not divisible by groups=%d", C, groups);
  DMME_REQUIRE((scale == nullptr) == (shift == nullptr), DMME_E_BADARG, "groupnorm_bwd: scale and shift come together");
  DMME_REQUIRE((dscale == nullptr) == (dshift == nullptr), DMME_E_BADARG, "groupnorm_bwd: dscale and dshift come together");
  DMME_REQUIRE(!dscale || scale, DMME_E_BADARG, "groupnorm_bwd: dscale without scale/shift");
  GnBwdParams p;
  p.gout = grad_out; p.src0 = src0; p.src1 = src1; p.c0 = c0; p.c1 = c1; p.n = n; p.hw = hw; p.groups = groups; p.eps = eps;
  p.gamma = gamma; p.beta = beta; p.scale = scale; p.shift = shift; p.ss_rows = ss_rows; p.ss_ld = ss_ld;
  p.mask = chan_mask; p.silu = apply_silu; p.gin0 = gin0; p.gin1 = gin1; p.add0 = add0; p.add1 = add1; p.sums = sums;
  const int cpg = C / groups;
  const size_t smem = sizeof(float) * (2 * cpg + 32);
  DMME_REQUIRE(smem <= 48 * 1024, DMME_E_SHAPE, "groupnorm_bwd: %d channels per group is too many", cpg);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  const bool fast = act_dtype == DMME_BF16 && C % 32 == 0 && c0 % 32 == 0 && (32 % cpg == 0) && hw <= 1024 && (hw * 4) % 32 == 0;
  if (fast) {
    const int nvec = hw * 4;
    const int threads = nvec < 256 ? nvec : 256;
    const int maxv = ceil_div(nvec, threads);
    const int blocks = n * (C / 32);
    if (maxv <= 1) gn_bwd_slab_kernel<1><<<blocks, threads, 0, st>>>(p);
    else if (maxv <= 4) gn_bwd_slab_kernel<4><<<blocks, threads, 0, st>>>(p);
    // 32x32 maps: 512 threads with 8 vectors each instead of 256 x 16 (255 registers, spills, eight warps per SM: 120 us
    // per launch for 100 MB of traffic)
    else if (nvec >= 2048 && nvec <= 4096) gn_bwd_slab_kernel<8><<<blocks, 512, 0, st>>>(p);
    else gn_bwd_slab_kernel<16><<<blocks, threads, 0, st>>>(p);
    rc = check_launch("gn_bwd_slab_kernel");
  } else {
    if (act_dtype == DMME_BF16) gn_bwd_kernel<__nv_bfloat16><<<n * groups, 256, smem, st>>>(p);
    else gn_bwd_kernel<float><<<n * groups, 256, smem, st>>>(p);
    rc = check_launch("gn_bwd_kernel");
  }
  if (rc) return rc;
  gn_bwd_finalize_kernel<<<ceil_div(C, 32), kGnFinWarps * 32, 0, st>>>(sums, n, C, gamma, beta, scale, ss_rows, ss_ld, dgamma, dbeta,
                                                         dscale, dshift, dss_ld);
  return check_launch("gn_bwd_finalize_kernel");
}

// ---- attention backward ------------------------------------------------------------------------------
extern "C" long long dmme_attention_bwd_workspace(int n, int heads, int L, int dh) {
  return (2LL * n * heads * L * L + static_cast<long long>(n) * heads * L * dh) * sizeof(float);
}

// Training-mode forward: the same strided products as the backward pass, with the softmax matrix P kept for it
// (fp32 [n*heads][L][L]).  o_tmp: fp32 [n*heads][L][dh] scratch.
extern "C" int dmme_attention_fwd_train(const void* q, const void* k, const void* v, long long batch_stride, int row_stride,
                                        int head_stride, int n, int heads, int L, int dh, float scale, int head_batch_swap,
                                        void* out, int act_dtype, float* p_out, float* o_tmp, void* stream) {
  DMME_REQUIRE(q && k && v && out && p_out && o_tmp, DMME_E_BADARG, "attention_fwd_train: null pointer");
  DMME_REQUIRE(n > 0 && heads > 0 && L > 0 && dh > 0, DMME_E_BADARG, "attention_fwd_train: bad sizes");
  DMME_REQUIRE(static_cast<long long>(n) * heads <= 65535, DMME_E_SHAPE, "attention_fwd_train: more than 65535 (image, head) pairs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // bf16 mode: the fused kernels also write the normalised P the backward pass needs -- tcgen05 at 256 tokens and at 64
  // tokens with 64-channel heads (attention_tc.cu), mma.sync elsewhere (attention_mma.cu)
  if (attn_tensor_core_multi_head_enabled() &&
      attn_tc_mh_supported(act_dtype, heads, L, dh, row_stride, head_stride, batch_stride, 0, q, k, v, out))
    return attn_tc_mh_forward(q, n, heads, L, dh, scale, head_batch_swap, out, p_out, st);
  if (attn_mma_supported(act_dtype, heads, L, dh, row_stride, head_stride, batch_stride, 0, q, k, v, out))
    return attn_mma_forward(q, k, v, batch_stride, row_stride, head_stride, n, heads, L, dh, scale, head_batch_swap, out,
                            p_out, st);
  const long long bh = static_cast<long long>(n) * heads;
  const long long LL = static_cast<long long>(L) * L, Ld = static_cast<long long>(L) * dh;
  int rc;
  GemmParams g;
  g.heads = heads; g.accumulate = 0;
  g.bf16_mma = act_dtype == DMME_BF16 ? 1 : 0;
  g.a = {q, act_dtype, batch_stride, head_stride, row_stride, 1};
  g.b = {k, act_dtype, batch_stride, head_stride, 1, row_stride};
  g.c = p_out; g.c_dtype = DMME_F32; g.c_bo = heads * LL; g.c_h = LL; g.c_r = L; g.c_c = 1;
  g.M = L; g.N = L; g.K = dh; g.alpha = scale;
  if ((rc = launch_gemm(g, static_cast<int>(bh), st))) return rc;
  softmax_rows_kernel<<<static_cast<unsigned>(ceil_div_ll(bh * L, 8)), 256, 0, st>>>(p_out, bh * L, L);
  if ((rc = check_launch("softmax_rows_kernel"))) return rc;
  g.a = {p_out, DMME_F32, heads * LL, LL, L, 1};
  g.b = {v, act_dtype, batch_stride, head_stride, row_stride, 1};
  g.c = o_tmp; g.c_bo = heads * Ld; g.c_h = Ld; g.c_r = dh; g.c_c = 1;
  g.M = L; g.N = dh; g.K = L; g.alpha = 1.0f;
  if ((rc = launch_gemm(g, static_cast<int>(bh), st))) return rc;
  const long long total = bh * Ld;
  if (act_dtype == DMME_BF16)
    attn_scatter_out_kernel<__nv_bfloat16><<<grid_1d(total, 256), 256, 0, st>>>(o_tmp, static_cast<__nv_bfloat16*>(out), n, heads, L, dh, head_batch_swap);
  else
    attn_scatter_out_kernel<float><<<grid_1d(total, 256), 256, 0, st>>>(o_tmp, static_cast<float*>(out), n, heads, L, dh, head_batch_swap);
  return check_launch("attn_scatter_out_kernel");
}

extern "C" int dmme_attention_bwd(const void* q, const void* k, const void* v, long long batch_stride, int row_stride,
                                  int head_stride, int n, int heads, int L, int dh, float scale, int head_batch_swap,
                                  const void* dout, void* dq, void* dk, void* dv, int act_dtype, const float* p_saved,
                                  void* workspace, long long workspace_bytes, void* stream) {
  DMME_REQUIRE(q && k && v && dout && dq && dk && dv && workspace, DMME_E_BADARG, "attention_bwd: null pointer");
  DMME_REQUIRE(n > 0 && heads > 0 && L > 0 && dh > 0, DMME_E_BADARG, "attention_bwd: bad sizes");
  DMME_REQUIRE(workspace_bytes >= dmme_attention_bwd_workspace(n, heads, L, dh), DMME_E_BADARG, "attention_bwd: workspace too small");
  DMME_REQUIRE(static_cast<long long>(n) * heads <= 65535, DMME_E_SHAPE, "attention_bwd: more than 65535 (image, head) pairs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long bh = static_cast<long long>(n) * heads;
  float* Pw = static_cast<float*>(workspace);         // [bh][L][L] (unused when the forward's P is passed in)
  float* dP = Pw + bh * L * L;                        // [bh][L][L]
  float* dO = dP + bh * L * L;                        // [bh][L][dh]
  const long long LL = static_cast<long long>(L) * L, Ld = static_cast<long long>(L) * dh;
  int rc;
  GemmParams g;
  g.heads = heads; g.accumulate = 0;
  g.bf16_mma = act_dtype == DMME_BF16 ? 1 : 0;
  const int rows_per_block = 8;
  const float* P = p_saved;
  if (P == nullptr) {
    // S = scale * Q K^T, P = softmax(S) (recomputed)
    g.a = {q, act_dtype, batch_stride, head_stride, row_stride, 1};
    g.b = {k, act_dtype, batch_stride, head_stride, 1, row_stride};
    g.c = Pw; g.c_dtype = DMME_F32; g.c_bo = heads * LL; g.c_h = LL; g.c_r = L; g.c_c = 1;
    g.M = L; g.N = L; g.K = dh; g.alpha = scale;
    if ((rc = launch_gemm(g, static_cast<int>(bh), st))) return rc;
    softmax_rows_kernel<<<static_cast<unsigned>(ceil_div_ll(bh * L, rows_per_block)), rows_per_block * 32, 0, st>>>(Pw, bh * L, L);
    if ((rc = check_launch("softmax_rows_kernel"))) return rc;
    P = Pw;
  }
  // dO in (batch, head) order
  {
    const long long total = bh * Ld;
    if (act_dtype == DMME_BF16)
      attn_gather_dout_kernel<__nv_bfloat16><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout), dO, n, heads, L, dh, head_batch_swap);
    else
      attn_gather_dout_kernel<float><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const float*>(dout), dO, n, heads, L, dh, head_batch_swap);
    if ((rc = check_launch("attn_gather_dout_kernel"))) return rc;
  }
  // dP = dO V^T
  g.a = {dO, DMME_F32, heads * Ld, Ld, dh, 1};
  g.b = {v, act_dtype, batch_stride, head_stride, 1, row_stride};
  g.c = dP; g.c_dtype = DMME_F32; g.c_bo = heads * LL; g.c_h = LL; g.c_r = L; g.c_c = 1;
  g.alpha = 1.0f; g.M = L; g.N = L; g.K = dh;
  if ((rc = launch_gemm(g, static_cast<int>(bh), st))) return rc;
  attn_ds_kernel<<<static_cast<unsigned>(ceil_div_ll(bh * L, rows_per_block)), rows_per_block * 32, 0, st>>>(P, dP, bh * L, L, scale);
  if ((rc = check_launch("attn_ds_kernel"))) return rc;
  // dQ = dS K
  g.a = {dP, DMME_F32, heads * LL, LL, L, 1};
  g.b = {k, act_dtype, batch_stride, head_stride, row_stride, 1};
  g.c = dq; g.c_dtype = act_dtype; g.c_bo = batch_stride; g.c_h = head_stride; g.c_r = row_stride; g.c_c = 1;
  g.M = L; g.N = dh; g.K = L;
  if ((rc = launch_gemm(g, static_cast<int>(bh), st))) return rc;
  // dK = dS^T Q
  g.a = {dP, DMME_F32, heads * LL, LL, 1, L};
  g.b = {q, act_dtype, batch_stride, head_stride, row_stride, 1};
  g.c = dk;
  if ((rc = launch_gemm(g, static_cast<int>(bh), st))) return rc;
  // dV = P^T dO
  g.a = {P, DMME_F32, heads * LL, LL, 1, L};
  g.b = {dO, DMME_F32, heads * Ld, Ld, dh, 1};
  g.c = dv;
  return launch_gemm(g, static_cast<int>(bh), st);
}

// ---- timestep MLP backward -----------------------------------------------------------------------------
// forward (temb.cu): s = [sin(t f), cos(t f)], h = silu(W1 s + b1), emb = silu(W2 h + b2), all = Wcat emb + bcat
// workspace: fp32 [rows][2 half] + 3 [rows][emb_dim]
extern "C" long long dmme_temb_bwd_workspace(int rows, int half, int emb_dim) {
  return (static_cast<long long>(rows) * 2 * half + 3LL * rows * emb_dim) * sizeof(float);
}

extern "C" int dmme_temb_bwd(const int64_t* t, int rows, const float* freq, int half, const float* w1, const float* b1,
                             const float* w2, const float* b2, int emb_dim, const float* hidden, const float* emb,
                             const float* wcat, int total, const float* d_all, float* dw1, float* db1, float* dw2, float* db2,
                             float* dwcat, float* dbcat, void* workspace, long long workspace_bytes, int bf16_mma,
                             void* stream) {
  DMME_REQUIRE(t && freq && w1 && b1 && w2 && b2 && hidden && emb && wcat && d_all && dw1 && db1 && dw2 && db2 && dwcat && dbcat && workspace,
               DMME_E_BADARG, "temb_bwd: null pointer");
  DMME_REQUIRE(rows > 0 && half > 0 && emb_dim > 0 && total > 0, DMME_E_BADARG, "temb_bwd: bad sizes");
  DMME_REQUIRE(workspace_bytes >= dmme_temb_bwd_workspace(rows, half, emb_dim), DMME_E_BADARG, "temb_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pos = 2 * half;
  float* S = static_cast<float*>(workspace);            // [rows][pos]
  float* Z = S + static_cast<long long>(rows) * pos;    // [rows][emb]   pre-activations (recomputed)
  float* G2 = Z + static_cast<long long>(rows) * emb_dim;  // [rows][emb] d emb -> d z2
  float* G1 = G2 + static_cast<long long>(rows) * emb_dim; // [rows][emb] d h -> d z1
  int rc;
  // dWcat = d_all^T emb ; dbcat = colsum(d_all) ; d_emb = d_all Wcat
  // the two products over the batched projection ([rows][total], total ~ 14k columns: 1.9 GFLOP each, 0.4 ms on the FFMA
  // kernel) take the tensor-core kernel in bf16 training mode (operands rounded to bf16, fp32 accumulation)
  if ((rc = gemm_f32(d_all, 1, total, emb, emb_dim, 1, dwcat, emb_dim, total, emb_dim, rows, 1.f, 0, st, bf16_mma))) return rc;
  colsum_kernel<<<ceil_div(total, 128), 128, 0, st>>>(d_all, rows, total, total, dbcat, 0);
  if ((rc = check_launch("colsum_kernel"))) return rc;
  if ((rc = gemm_f32(d_all, total, 1, wcat, emb_dim, 1, G2, emb_dim, rows, emb_dim, total, 1.f, 0, st, bf16_mma))) return rc;
  // z2 = W2 h + b2 (recomputed) ; d z2 = d emb * silu'(z2)
  const long long ne = static_cast<long long>(rows) * emb_dim;
  if ((rc = launch_linear(hidden, nullptr, nullptr, rows, emb_dim, w2, b2, emb_dim, 0, Z, st, "temb_bwd(z2)"))) return rc;
  silu_bwd_kernel<<<grid_1d(ne, 256), 256, 0, st>>>(G2, Z, ne);
  if ((rc = check_launch("silu_bwd_kernel"))) return rc;
  // dW2 = dz2^T h ; db2 = colsum(dz2) ; d h = dz2 W2
  if ((rc = gemm_f32(G2, 1, emb_dim, hidden, emb_dim, 1, dw2, emb_dim, emb_dim, emb_dim, rows, 1.f, 0, st))) return rc;
  colsum_kernel<<<ceil_div(emb_dim, 128), 128, 0, st>>>(G2, rows, emb_dim, emb_dim, db2, 0);
  if ((rc = check_launch("colsum_kernel"))) return rc;
  if ((rc = gemm_f32(G2, emb_dim, 1, w2, emb_dim, 1, G1, emb_dim, rows, emb_dim, emb_dim, 1.f, 0, st))) return rc;
  // z1 = W1 s + b1 (recomputed from t) ; d z1 = d h * silu'(z1)
  if ((rc = launch_linear(nullptr, t, freq, rows, pos, w1, b1, emb_dim, 0, Z, st, "temb_bwd(z1)"))) return rc;
  silu_bwd_kernel<<<grid_1d(ne, 256), 256, 0, st>>>(G1, Z, ne);
  if ((rc = check_launch("silu_bwd_kernel"))) return rc;
  sin_embed_kernel<<<ceil_div(rows * pos, 256), 256, 0, st>>>(t, freq, rows, half, S);
  if ((rc = check_launch("sin_embed_kernel"))) return rc;
  // dW1 = dz1^T s ; db1 = colsum(dz1)
  if ((rc = gemm_f32(G1, 1, emb_dim, S, pos, 1, dw1, pos, emb_dim, pos, rows, 1.f, 0, st))) return rc;
  colsum_kernel<<<ceil_div(emb_dim, 128), 128, 0, st>>>(G1, rows, emb_dim, emb_dim, db1, 0);
  return check_launch("colsum_kernel");
}

extern "C" int dmme_colsum_f32(const float* in, int rows, int cols, long long ld, float* out, int accumulate, void* stream) {
  DMME_REQUIRE(in && out && rows > 0 && cols > 0, DMME_E_BADARG, "colsum: bad arguments");
  colsum_kernel<<<ceil_div(cols, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(in, rows, cols, ld, out, accumulate);
  return check_launch("colsum_kernel");
}

// ---- L_simple ----------------------------------------------------------------------------------------
// loss_out[0] = mean((eps - noise)^2); d_eps (optional) = grad_scale * 2 (eps - noise) / numel; partial: >= 1024 floats
extern "C" int dmme_mse_loss(const float* eps, const float* noise, long long numel, float grad_scale, float* d_eps,
                             float* loss_out, float* partial, void* stream) {
  DMME_REQUIRE(eps && noise && loss_out && partial && numel > 0, DMME_E_BADARG, "mse_loss: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blocks = grid_1d(numel, 256);
  if (blocks > 1024) blocks = 1024;
  const float inv = 1.0f / static_cast<float>(numel);
  mse_partial_kernel<<<blocks, 256, 0, st>>>(eps, noise, numel, inv, grad_scale, d_eps, partial);
  int rc = check_launch("mse_partial_kernel");
  if (rc) return rc;
  sum_partials_kernel<<<1, 32, 0, st>>>(partial, blocks, inv, loss_out);
  return check_launch("sum_partials_kernel");
}

// ---- IDDPM hybrid / VLB loss ----------------------------------------------------------------------------
// loss_out[0] = w_simple * L_simple + w_vlb * L_vlb, loss_out[1] = L_simple, loss_out[2] = L_vlb.
// hybrid (diffusion_models/iddpm.py:109-116): w_simple = 1, w_vlb = gamma; "vlb": w_simple = 0, w_vlb = 1.
// d_out (optional, [n][2c][hw]) = grad_scale * d loss / d model_out.  partial: fp32 workspace of >= 2048 floats.
extern "C" int dmme_iddpm_loss(const float* model_out, const float* x_t, const float* x_0, const int64_t* t,
                               const float* beta, const float* alpha, const float* alpha_bar, int n, int c, int hw,
                               float w_simple, float w_vlb, float grad_scale, float* d_out, float* loss_out, float* partial,
                               void* stream) {
  DMME_REQUIRE(model_out && x_t && x_0 && t && beta && alpha && alpha_bar && loss_out && partial, DMME_E_BADARG,
               "iddpm_loss: null pointer");
  DMME_REQUIRE(n > 0 && c > 0 && hw > 0, DMME_E_BADARG, "iddpm_loss: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n) * c * hw;
  int blocks = grid_1d(total, 256);
  if (blocks > 1024) blocks = 1024;
  const float inv = 1.0f / static_cast<float>(total);
  iddpm_loss_kernel<<<blocks, 256, 0, st>>>(model_out, x_t, x_0, t, beta, alpha, alpha_bar, n, c, hw,
                                            grad_scale * w_simple * inv, grad_scale * w_vlb * inv, d_out, partial);
  int rc = check_launch("iddpm_loss_kernel");
  if (rc) return rc;
  iddpm_loss_finish_kernel<<<1, 32, 0, st>>>(partial, blocks, inv, w_simple, w_vlb, loss_out);
  return check_launch("iddpm_loss_finish_kernel");
}
