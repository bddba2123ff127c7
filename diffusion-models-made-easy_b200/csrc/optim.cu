// Fused optimizer tail of one training step: global gradient-norm clip (Lightning `gradient_clip_val`,
// configs/ddpm/cifar10.yaml:24 -> torch.nn.utils.clip_grad_norm_), Adam (lit_modules/ddpm.py:130, torch.optim.Adam
// defaults), the WarmupLR scale (lr_scheduler/warmup.py:10-19, passed in as the step's learning rate) and the EMA update
// `ema = d * ema + (1 - d) * w` (callbacks/ema.py:169-176) as TWO multi-tensor launches over all parameters:
//
//   pass 1  grad_sumsq    : per-block partial sums of g^2 (fixed chunk -> block mapping, fixed-order tree: deterministic)
//   pass 2  adam_ema_step : every block re-reduces the partials in the same order (so all blocks see the same norm),
//                           then streams its chunks: g, m, v, w, ema read once, m, v, w, ema written once
//
// HBM-bound: 4 B (g) + 2 x 16 B (m, v, w, ema read + write) = 36 B per parameter in pass 2, 4 B in pass 1; the torch path
// (clip_grad_norm_ + foreach Adam + two foreach EMA ops) moves about 100 B per parameter in ~20 launches.
// Tensors are addressed through a device table of (pointer, size) built by the host side (dmme_b200/optim.py).
#include <math.h>

#include "common.cuh"

namespace dmme {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = 4096;  // elements per work item: a block walks items round-robin

struct OptTensor {
  float* w;
  const float* g;
  float* m;
  float* v;
  float* ema;  // may be null
  long long numel;
  long long first_item;  // index of this tensor's first work item
};

__device__ __forceinline__ int find_tensor(const OptTensor* __restrict__ t, int count, long long item) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid].first_item <= item) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (kOptThreads >> 5) ? red[lane] : 0.f;
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  return r;
}

__global__ void __launch_bounds__(kOptThreads) grad_sumsq_kernel(const OptTensor* __restrict__ t, int count,
                                                                 long long items, float* __restrict__ partial) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int ti = find_tensor(t, count, it);
    const long long off = (it - t[ti].first_item) * kOptChunk;
    const long long n = t[ti].numel - off < kOptChunk ? t[ti].numel - off : kOptChunk;
    if (t[ti].g == nullptr) continue;
    const float* g = t[ti].g + off;
    for (long long i = threadIdx.x; i < n; i += kOptThreads) {
      const float x = g[i];
      acc = fmaf(x, x, acc);
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct AdamParams {
  float step_size;          // lr / (1 - beta1^t), computed in double like torch's Python scalar
  float beta2, eps;
  float omb1, omb2;         // 1 - beta1, 1 - beta2 (double arithmetic, rounded once)
  float bias2_sqrt;         // sqrt(1 - beta2^t)
  float max_norm;           // <= 0: no clipping
  float ema_decay, om_decay;  // used when a tensor has an ema pointer
  int n_partial;
};

__global__ void __launch_bounds__(kOptThreads) adam_ema_step_kernel(const OptTensor* __restrict__ t, int count,
                                                                    long long items, const float* __restrict__ partial,
                                                                    AdamParams a, float* __restrict__ norm_out) {
  __shared__ float red[kOptThreads / 32];
  float clip = 1.f;
  if (a.max_norm > 0.f) {
    // same order in every block: identical total norm everywhere
    float acc = 0.f;
    for (int i = threadIdx.x; i < a.n_partial; i += kOptThreads) acc += partial[i];
    const float total = sqrtf(block_sum(acc, red));
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = total;
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    clip = fminf(a.max_norm / (total + 1e-6f), 1.f);
  }
  const float step_size = a.step_size, omb1 = a.omb1, omb2 = a.omb2, omd = a.om_decay;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int ti = find_tensor(t, count, it);
    const OptTensor T = t[ti];
    if (T.g == nullptr) continue;
    const long long off = (it - T.first_item) * kOptChunk;
    const long long n = T.numel - off < kOptChunk ? T.numel - off : kOptChunk;
    for (long long i = threadIdx.x; i < n; i += kOptThreads) {
      const long long j = off + i;
      const float g = T.g[j] * clip;
      // torch.optim.Adam (single tensor form): exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      float m = T.m[j];
      m = fmaf(omb1, g - m, m);
      float v = T.v[j] * a.beta2;
      v = fmaf(omb2 * g, g, v);
      const float denom = sqrtf(v) / a.bias2_sqrt + a.eps;
      float w = T.w[j];
      w = w - step_size * (m / denom);
      T.m[j] = m;
      T.v[j] = v;
      T.w[j] = w;
      if (T.ema) {
        // callbacks/ema.py:169-176: ema.mul_(decay); ema.add_(w, alpha = 1 - decay)
        float e = T.ema[j] * a.ema_decay;
        T.ema[j] = fmaf(omd, w, e);
      }
    }
  }
}

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_optim_table_entry_bytes(void) { return static_cast<int>(sizeof(OptTensor)); }
extern "C" int dmme_optim_chunk(void) { return kOptChunk; }

// table: device array of `count` entries {w, g, m, v, ema, numel, first_item} (7 x 8 bytes); partial: >= grid floats.
extern "C" int dmme_adam_ema_step(const void* table, int count, long long items, double lr, double beta1, double beta2,
                                  double eps, int step, double max_norm, double ema_decay, float* partial, int grid,
                                  float* norm_out, void* stream) {
  DMME_REQUIRE(table && count > 0 && items > 0 && partial && grid > 0 && step >= 1, DMME_E_BADARG,
               "adam_ema_step: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const OptTensor* t = static_cast<const OptTensor*>(table);
  if (max_norm > 0.0) {
    grad_sumsq_kernel<<<grid, kOptThreads, 0, st>>>(t, count, items, partial);
    int rc = check_launch("grad_sumsq_kernel");
    if (rc) return rc;
  }
  AdamParams a;
  // scalars in double, rounded once (torch keeps them as Python floats)
  a.step_size = static_cast<float>(lr / (1.0 - pow(beta1, step)));
  a.bias2_sqrt = static_cast<float>(sqrt(1.0 - pow(beta2, step)));
  a.beta2 = static_cast<float>(beta2);
  a.eps = static_cast<float>(eps);
  a.omb1 = static_cast<float>(1.0 - beta1);
  a.omb2 = static_cast<float>(1.0 - beta2);
  a.max_norm = static_cast<float>(max_norm);
  a.ema_decay = static_cast<float>(ema_decay);
  a.om_decay = static_cast<float>(1.0 - ema_decay);
  a.n_partial = grid;
  adam_ema_step_kernel<<<grid, kOptThreads, 0, st>>>(t, count, items, partial, a, norm_out);
  return check_launch("adam_ema_step_kernel");
}
