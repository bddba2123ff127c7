// C-ABI glue: error reporting, launch accounting and the convolution dispatcher.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace dmme {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool conv_tc_supported(const dmme_conv_desc& d);
int conv_tc_forward(const dmme_conv_desc& d, cudaStream_t stream);
int conv_generic_forward(const dmme_conv_desc& d, cudaStream_t stream);
bool conv_in_supported(const dmme_conv_desc& d);
bool conv_out_supported(const dmme_conv_desc& d);
int conv_small_forward(const dmme_conv_desc& d, cudaStream_t stream);
bool conv_halo_supported(const dmme_conv_desc& d);
bool conv_halo_preferred(const dmme_conv_desc& d);
bool conv_tct_epilogue_norm(const dmme_conv_desc& d);
int conv_halo_forward(const dmme_conv_desc& d, cudaStream_t stream);
bool conv_in_tc_supported(const dmme_conv_desc& d);
int conv_in_tc_forward(const dmme_conv_desc& d, cudaStream_t stream);
bool conv_out_tc_supported(const dmme_conv_desc& d);
bool conv_out_dx_supported(const dmme_conv_desc& d);
long long conv_splitk_workspace(const dmme_conv_desc& d);
int conv_out_tc_forward(const dmme_conv_desc& d, cudaStream_t stream);

}  // namespace dmme

using namespace dmme;

extern "C" int dmme_abi_version(void) { return DMME_ABI_VERSION; }
// 1 when the library was built with -DDMME_EXPERIMENTAL (the cta_group::2 conv and the weight-multicast halo variants:
// parity-green negative results kept out of the shipped build)
extern "C" int dmme_has_experimental(void) {
#ifdef DMME_EXPERIMENTAL
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* dmme_last_error(void) { return g_err; }
extern "C" long long dmme_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void dmme_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" int dmme_conv2d_uses_tc(const dmme_conv_desc* d) {
  if (!d) return 0;
  if (d->kernel == DMME_CONV_GENERIC) return 0;
  if (d->kernel == DMME_CONV_HALO) return conv_halo_supported(*d) ? 1 : 0;
  if (d->kernel == DMME_CONV_AUTO && conv_out_tc_supported(*d)) return 1;
  return conv_tc_supported(*d) ? 1 : 0;
}

extern "C" int dmme_conv2d_writes_stats(const dmme_conv_desc* d) {
  if (!d || d->kernel == DMME_CONV_GENERIC || d->out_layout != DMME_OUT_NHWC || d->cout % 4) return 0;
  if (d->kernel == DMME_CONV_HALO) return conv_halo_supported(*d) ? 1 : 0;
  if (conv_tc_supported(*d)) return 1;
  if (d->kernel == DMME_CONV_AUTO && conv_in_supported(*d) && (static_cast<long long>(d->h_in) * d->w_in) % 64 == 0) return 1;
  return 0;
}

// fused GroupNorm of the input: the halo kernel's row-tile mode (8x8 / 16x16 / 32x32 maps) has the transform stage
static bool runs_halo_rows(const dmme_conv_desc& d) {
  const bool halo = d.kernel == DMME_CONV_HALO ? conv_halo_supported(d)
                                               : (d.kernel == DMME_CONV_AUTO && conv_tc_supported(d) && conv_halo_preferred(d));
  return halo && d.w_in >= 8;
}

// ... and so has the output conv's row-tile kernel (32x32 maps)
static bool runs_out_dx(const dmme_conv_desc& d) {
  return d.kernel == DMME_CONV_AUTO && !conv_tc_supported(d) && conv_out_dx_supported(d) && d.c1 == 0;
}

extern "C" int dmme_conv2d_fuses_gn(const dmme_conv_desc* d) { return d && (runs_halo_rows(*d) || runs_out_dx(*d)) ? 1 : 0; }

// split-K is a choice of the AUTO / TC dispatch below 16x16-and-up halo territory: the same conditions as the dispatch
static bool takes_conv_tc(const dmme_conv_desc& d) {
  if (d.kernel == DMME_CONV_TC) return conv_tc_supported(d);
  return d.kernel == DMME_CONV_AUTO && conv_tc_supported(d) && !conv_halo_preferred(d);
}

extern "C" int dmme_conv2d_fuses_sampler(const dmme_conv_desc* d) {
  return d && d->kernel == DMME_CONV_AUTO && !conv_tc_supported(*d) && conv_out_tc_supported(*d) ? 1 : 0;
}

extern "C" int dmme_conv2d_epilogue_norm(const dmme_conv_desc* d) {
  return d && takes_conv_tc(*d) && conv_tct_epilogue_norm(*d) ? 1 : 0;
}

extern "C" long long dmme_conv2d_splitk_workspace(const dmme_conv_desc* d) {
  if (!d || !takes_conv_tc(*d)) return 0;
  return conv_splitk_workspace(*d);
}

extern "C" int dmme_conv2d_fwd(const dmme_conv_desc* d, void* stream) {
  DMME_REQUIRE(d != nullptr, DMME_E_BADARG, "conv2d_fwd: null descriptor");
  DMME_REQUIRE(d->gn_ab == nullptr || runs_halo_rows(*d) || runs_out_dx(*d), DMME_E_UNSUPPORTED,
               "conv2d_fwd: a fused GroupNorm (gn_ab) needs the halo kernel or the 32x32 output conv; ask dmme_conv2d_fuses_gn");
  DMME_REQUIRE((d->out_norm[0].out == nullptr && d->out_norm[1].out == nullptr) ||
                   (takes_conv_tc(*d) && (d->splitk_ws != nullptr || conv_tct_epilogue_norm(*d))),
               DMME_E_UNSUPPORTED,
               "conv2d_fwd: out_norm needs the split-K path (ask dmme_conv2d_splitk_workspace) or the transposed kernel on an "
               "8x8 map (ask dmme_conv2d_epilogue_norm)");
  DMME_REQUIRE(d->sampler == nullptr || d->sampler->kind == DMME_SAMPLER_NONE || dmme_conv2d_fuses_sampler(d), DMME_E_UNSUPPORTED,
               "conv2d_fwd: a fused sampler update needs the tcgen05 output-conv kernel; ask dmme_conv2d_fuses_sampler");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d->kernel) {
    case DMME_CONV_TC:
      return conv_tc_forward(*d, st);
    case DMME_CONV_GENERIC:
      return conv_generic_forward(*d, st);
    case DMME_CONV_HALO:
      return conv_halo_forward(*d, st);
    case DMME_CONV_AUTO:
      if (conv_tc_supported(*d)) return conv_halo_preferred(*d) ? conv_halo_forward(*d, st) : conv_tc_forward(*d, st);
      if (conv_out_tc_supported(*d)) return conv_out_tc_forward(*d, st);
      if (conv_in_tc_supported(*d)) return conv_in_tc_forward(*d, st);
      if (conv_in_supported(*d) || conv_out_supported(*d)) return conv_small_forward(*d, st);
      return conv_generic_forward(*d, st);
    default:
      set_error("conv2d_fwd: unknown kernel selector %d", d->kernel);
      return DMME_E_BADARG;
  }
}
