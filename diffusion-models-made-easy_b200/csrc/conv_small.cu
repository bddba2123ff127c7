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
};

// ---- input conv: one thread = one pixel x 8 output channels --------------------------------------
__global__ void __launch_bounds__(256) conv_in_kernel(const ConvSmallParams p) {
  extern __shared__ float wsm[];  // [9*cin][cout] then bias [cout]
  const int K = 9 * p.cin;
  for (int i = threadIdx.x; i < K * p.cout; i += blockDim.x) wsm[i] = p.weight[i];
  float* bsm = wsm + K * p.cout;
  for (int i = threadIdx.x; i < p.cout; i += blockDim.x) bsm[i] = p.bias ? p.bias[i] : 0.f;
  __syncthreads();
  const int cg = p.cout / 8;
  const long long total = static_cast<long long>(p.n) * p.h * p.w * cg;
  const float* x = static_cast<const float*>(p.src);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
  const long long plane = static_cast<long long>(p.h) * p.w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % cg);
    const long long pix = i / cg;
    const int xx = static_cast<int>(pix % p.w);
    const int yy = static_cast<int>((pix / p.w) % p.h);
    const long long ni = pix / plane;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bsm[g * 8 + j];
    for (int r = 0; r < 3; ++r) {
      const int iy = yy + r - 1;
      if (iy < 0 || iy >= p.h) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = xx + s - 1;
        if (ix < 0 || ix >= p.w) continue;
        for (int ci = 0; ci < p.cin; ++ci) {
          const float v = __ldg(x + (ni * p.cin + ci) * plane + static_cast<long long>(iy) * p.w + ix);
          const float4* wr = reinterpret_cast<const float4*>(wsm + ((r * 3 + s) * p.cin + ci) * p.cout + g * 8);
          const float4 w0 = wr[0], w1 = wr[1];
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + pix * p.cout + g * 8) = o;
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
    for (int r = 0; r < 3; ++r) {
      const int iy = yy + r - 1;
      if (iy < 0 || iy >= p.h) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = xx + s - 1;
        if (ix < 0 || ix >= p.w) continue;
        const __nv_bfloat16* row = src + ((ni * p.h + iy) * p.w + ix) * p.cin;
        const int kbase = (r * 3 + s) * p.cin;
        for (int u = 0; u < units; ++u) {
          const int ch = (u * 8 + lane8) * 8;
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + ch));
          float f[8];
          unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
          unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            const float4* wr = reinterpret_cast<const float4*>(wsm + c * K + kbase + ch);
            const float4 w0 = wr[0], w1 = wr[1];
            acc[c] = fmaf(f[0], w0.x, acc[c]); acc[c] = fmaf(f[1], w0.y, acc[c]);
            acc[c] = fmaf(f[2], w0.z, acc[c]); acc[c] = fmaf(f[3], w0.w, acc[c]);
            acc[c] = fmaf(f[4], w1.x, acc[c]); acc[c] = fmaf(f[5], w1.y, acc[c]);
            acc[c] = fmaf(f[6], w1.z, acc[c]); acc[c] = fmaf(f[7], w1.w, acc[c]);
          }
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
      float v = acc[0];
#pragma unroll
      for (int c = 1; c < COUT; ++c)
        if (lane8 == c) v = acc[c];
      v += p.bias ? p.bias[lane8] : 0.f;
      out[(ni * COUT + lane8) * plane + static_cast<long long>(yy) * p.w + xx] = v;
    }
  }
}

static bool plain(const dmme_conv_desc& d) {
  return d.ksize == 3 && d.stride == 1 && !d.upsample && d.c1 == 0 && d.rc0 == 0 && d.rc1 == 0 && !d.temb &&
         !d.addend && d.act_dtype == DMME_BF16;
}
bool conv_in_supported(const dmme_conv_desc& d) {
  return plain(d) && d.in_layout == DMME_IN_NCHW_F32 && d.out_layout == DMME_OUT_NHWC && d.c0 <= 4 &&
         d.cout % 8 == 0 && d.cout <= 512;
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
  if (conv_in_supported(d)) {
    const size_t smem = sizeof(float) * (static_cast<size_t>(9) * d.c0 * d.cout + d.cout);
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e != cudaSuccess) { set_error("conv_in: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
      configured = true;
    }
    const long long total = static_cast<long long>(d.n) * d.h_in * d.w_in * (d.cout / 8);
    const long long blocks = ceil_div_ll(total, 256);
    const int grid = static_cast<int>(blocks < 148 * 8 ? blocks : 148 * 8);
    conv_in_kernel<<<grid, 256, smem, stream>>>(p);
    return check_launch("conv_in_kernel");
  }
  DMME_REQUIRE(conv_out_supported(d), DMME_E_SHAPE, "conv_small: unsupported shape");
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
