/*
 * dmme_b200.h -- C ABI of the B200-native hot path of diffusion-models-made-easy (dmme 0.5.2).
 *
 * The reference has no FFI layer: its hot path is a tree of torch.nn.Module calls.  Each entry
 * point below replaces the ATen op sequence of one reference call site (cited per function,
 * paths relative to the reference checkout).  All pointers are raw device pointers owned by
 * the caller (torch owns the memory, kernels borrow it for the duration of the launch), all
 * launches go to the given cudaStream_t (passed as void*), nothing synchronises, nothing
 * allocates: every entry point is CUDA-graph capturable.
 *
 * Return value: 0 on success, a negative DMME_E_* code for an argument/shape error, or a
 * positive cudaError_t.  dmme_last_error() returns a thread-local message for the last failure.
 *
 * Activation layout between kernels is NHWC ("pixel rows of channels"); activation storage
 * type is selected per call by `act_dtype` (DMME_BF16 for the tensor-core path, DMME_F32 for
 * the fp32 parity mode).  Image-space tensors (x_t, eps, noise) stay NCHW fp32 exactly as the
 * reference holds them.
 */
#ifndef DMME_B200_H
#define DMME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMME_ABI_VERSION 8
#define DMME_STATS_FRAC_BITS 20

enum { DMME_BF16 = 0, DMME_F32 = 1 };

enum {
  DMME_OK = 0,
  DMME_E_BADARG = -1,      /* null pointer, non-positive size */
  DMME_E_SHAPE = -2,       /* shape not supported by the requested kernel */
  DMME_E_UNSUPPORTED = -3, /* feature combination not implemented */
  DMME_E_DRIVER = -4       /* CUDA driver entry point unavailable (no tensor-map encoder) */
};

/* input / output layouts of a convolution call */
enum { DMME_IN_NHWC = 0, DMME_IN_NCHW_F32 = 1 };
enum {
  DMME_OUT_NHWC = 0,     /* out[n][y][x][cout], act_dtype */
  DMME_OUT_NCHW_F32 = 1, /* out[n][cout][y][x], fp32 (eps in image space) */
  DMME_OUT_QKV = 2       /* cout = 3*C: q -> out[n][L][C], k -> out2[n][L][C], v -> out3[n][C][L] (V^T) */
};

/* which convolution kernel to run */
enum {
  DMME_CONV_AUTO = 0,    /* tcgen05 when the shape allows, generic otherwise */
  DMME_CONV_GENERIC = 1, /* FFMA implicit GEMM, any shape, fp32 math */
  DMME_CONV_TC = 2,      /* tcgen05/TMEM + TMA implicit GEMM, one A tile per filter tap (any supported shape) */
  DMME_CONV_HALO = 3     /* tcgen05 3x3 stride-1 kernel that keeps the activation halo tile in shared memory for all
                            nine taps */
};

/*
 * One fused convolution:   out = conv_{k x k, stride, pad k/2}( cat(src0, src1) [nearest x2] )
 *                                + conv_{1x1}( cat(res0, res1) )        (optional fused residual conv)
 *                                + bias + temb[n or 0][:] + addend      (optional epilogue terms)
 * Replaces: nn.Conv2d call sites models/ddpm.py:30,51,52,109,147,162,219,277 together with the
 * in-place adds of ResBlock.forward models/ddpm.py:129,131, the skip torch.cat models/ddpm.py:310,
 * nn.Upsample models/ddpm.py:161 and the attention residual models/ddpm.py:75.
 */
/*
 * Fused GroupNorm(+SiLU) of the conv OUTPUT for one consumer (split-K path only, see dmme_conv_desc.splitk_ws): besides the
 * raw tensor `out`, the finishing pass writes  y = [silu]( [(1 + scale)] * GN(out) * gamma + beta [+ shift] )  -- the
 * operand the consumer's conv reads (norm_act_drop_conv models/ddpm.py:25-35, Attention.norm models/ddpm.py:73, the
 * scale-shift norm models/iddpm.py:119).  When the consumer normalises a channel concat (models/ddpm.py:310) this tensor
 * is one part of it: gamma / beta / scale / shift then point at this part's first channel and cpg is the CONSUMER's group
 * width (groups never straddle the two parts).
 */
typedef struct dmme_out_norm {
  void* out;                          /* NHWC act_dtype tensor of the output shape; NULL: unused */
  const float* gamma; const float* beta;
  const float* scale; const float* shift; /* optional [ss_rows][ss_ld] fp32, rows = 1 (broadcast) or n */
  int ss_rows, ss_ld;
  int cpg;                            /* channels per group of the consumer's GroupNorm (1..32, divides 32) */
  int silu;
  float eps;
} dmme_out_norm;

/*
 * Sampler update fused into the OUTPUT conv's epilogue (models/ddpm.py:277 followed by the tail of
 * DDPM.sampling_step diffusion_models/ddpm.py:94-110, DDIM.sampling_step diffusion_models/ddim.py:55-77 or
 * IDDPM.sampling_step diffusion_models/iddpm.py:118-164): eps (and v) never leave the registers, x_t is updated in
 * place, the noise is drawn in the epilogue (Philox4x32-10 keyed by (seed, t, element index), the same draws as
 * dmme_ddpm_step) or read from `noise`.  Same arithmetic, same bits as the stand-alone dmme_*_step kernels.
 * Output-conv tcgen05 kernel only: ask dmme_conv2d_fuses_sampler.
 */
enum { DMME_SAMPLER_NONE = 0, DMME_SAMPLER_DDPM = 1, DMME_SAMPLER_DDIM = 2, DMME_SAMPLER_IDDPM = 3 };
typedef struct dmme_sampler_epilogue {
  int kind;                           /* DMME_SAMPLER_* */
  float* x;                           /* x_t, NCHW fp32 [n][C][h][w], updated in place (C = cout, IDDPM: cout / 2) */
  const float* noise;                 /* optional injected standard normals shaped like x (NULL: in-kernel Philox) */
  const float* beta; const float* alpha; const float* alpha_bar; /* schedule tables of length table_len */
  const int64_t* t_ptr;               /* device scalar: t (DDPM / IDDPM) or the sub-sequence index i (DDIM) */
  const int64_t* tau;                 /* DDIM: int64 [tau_len] */
  int table_len, tau_len;
  unsigned long long seed, noise_offset;
} dmme_sampler_epilogue;

typedef struct dmme_conv_desc {
  const void* src0; const void* src1; /* NHWC activations (or NCHW fp32 image when in_layout says so) */
  int c0, c1;                         /* channels of src0 / src1 (c1 = 0: no concat) */
  const void* res0; const void* res1; /* operands of the fused 1x1 residual conv, output resolution */
  int rc0, rc1;                       /* channels of res0 / res1 (0: none) */
  int n, h_in, w_in;                  /* batch and spatial size of src0/src1 */
  int ksize;                          /* 1 or 3 */
  int stride;                         /* 1 or 2 */
  int upsample;                       /* 1: nearest-neighbour x2 of the source before the conv;
                                         2: zero-dilated x2 source (data gradient of a stride-2 conv; generic kernel);
                                         3: nearest x2 + 3x3 conv as four 2x2 phase convs of the low-resolution source
                                            (tcgen05 path; weight = [4 phases][cout][4 taps x cin] bf16, out is x2) */
  int cout;
  const void* weight;                 /* packed by dmme_pack_conv_weight for the chosen kernel */
  const float* bias;                  /* [cout] fp32, may be NULL */
  const float* temb;                  /* optional [temb_rows][temb_ld] fp32, rows = 1 (broadcast) or n */
  int temb_rows, temb_ld;
  const void* addend;                 /* optional NHWC tensor of the output shape, act_dtype */
  void* out; void* out2; void* out3;
  long long* stats;                   /* optional [n][cout/4][2] int64 fixed-point (2^-DMME_STATS_FRAC_BITS) sums of
                                         the stored output and of its square per 4-channel micro-group, accumulated
                                         with integer atomics (order-independent, hence deterministic); the caller
                                         zeroes it.  Consumed by dmme_groupnorm_fwd.  tcgen05 path, NHWC output only. */
  int in_layout, out_layout;
  int act_dtype;
  int kernel;                         /* DMME_CONV_* */
  const float* gn_ab;                 /* optional fused GroupNorm(+SiLU) of the conv INPUT (norm_act_drop_conv,
                                         models/ddpm.py:25-35): interleaved (a, b) fp32 pairs [n][c0 + c1][2] from
                                         dmme_groupnorm_coeff; the kernel convolves [silu](a * x + b) with zero padding
                                         applied after the activation, as the reference does.  Halo kernel (16x16 and
                                         32x32 ResBlock convs) and the 32x32 output conv only: ask dmme_conv2d_fuses_gn
                                         first */
  int gn_silu;                        /* 1: SiLU after the fused norm */
  void* splitk_ws;                    /* optional fp32 workspace of dmme_conv2d_splitk_workspace(desc) bytes.  With it, a
                                         3x3 conv whose grid would leave most SMs idle (the 4x4 / 8x8 levels, small
                                         batches) runs split-K: every (pixel tile, channel tile, K slice) is a work unit
                                         that stores its fp32 partial tile, and a finishing pass sums the slices, adds
                                         bias / temb / addend, writes `out` and `stats` and applies out_norm[] */
  long long splitk_ws_bytes;
  dmme_out_norm out_norm[2];          /* optional fused GroupNorm(+SiLU) of the output for up to two consumers */
  const dmme_sampler_epilogue* sampler; /* optional fused sampler update (output conv); `out` may then be NULL: eps is
                                         not written at all */
} dmme_conv_desc;

/* library / device ------------------------------------------------------------------------- */
int dmme_abi_version(void);
/* 1 when built with -DDMME_EXPERIMENTAL: the cta_group::2 transposed conv (dmme_set_conv_pair_mode) and the
 * weight-multicast halo conv (dmme_set_conv_halo_multicast) exist; 0 (the shipped build): those switches do nothing */
int dmme_has_experimental(void);
const char* dmme_last_error(void);
/* number of kernel launches issued through this library since the last reset (process-wide) */
long long dmme_launch_count(void);
void dmme_reset_launch_count(void);

/* A/B switch for measurements: 0 = AUTO never picks the halo kernel, 1 = default; further bits (with bit 0 set): 16 = the
 * chunks of a fused 1x1 residual after the conv chunks instead of between them, 32 = equal row tiles only (no tail tiles),
 * 64 = the round-1 rule that keeps 16x16 convs with a wide fused residual on the transposed kernel, 128 = AUTO also takes the
 * 8x8 convs of batches >= 160 (row tiles of two whole images, GroupNorm inside: measured slower, off by default) */
void dmme_set_conv_halo_mode(int mode);
int dmme_get_conv_halo_mode(void);
/* halo kernel: clusters of two CTAs share the weight stream through TMA multicast.  0 = off (default; measured slower),
 * 1 = launches with several work items per CTA, 2 = every launch */
void dmme_set_conv_halo_multicast(int mode);
/* A/B switch: 0 = AUTO/TC never use the transposed tcgen05 kernel (thread = channel epilogue), 1 = default
 * (where it has a unit for most SMs), 2 = wherever it is supported */
void dmme_set_conv_tct_mode(int mode);
int dmme_get_conv_tct_mode(void);
/* A/B switch for the cta_group::2 (two-SM MMA) variant of the transposed kernel: 0 = never (default: it measured no
 * faster), 1 = 256-channel 3x3 convs with enough work units, 2 = wherever it is supported */
void dmme_set_conv_pair_mode(int mode);
/* A/B switch: 0 = multi-head attention (models/iddpm.py:16-59) stays on the CUDA-core kernel, 1 = mma.sync kernel */
void dmme_set_attn_mma_mode(int mode);
/* A/B switch: 0 = the output conv (models/ddpm.py:277) stays on the FFMA kernel, 1 = tcgen05 (default; on 32x32 maps the
 * row-tile kernel that also applies desc->gn_ab), 2 = tcgen05 per-tap kernel on every size */
void dmme_set_conv_out_tc_mode(int mode);
/* A/B switch: 0 = the input conv (models/ddpm.py:219) stays on the FFMA kernel, 1 = tcgen05 on 32x32 images from 64
 * images up (default), 2 = at every batch size */
void dmme_set_conv_in_tc_mode(int mode);

/* weights ---------------------------------------------------------------------------------- */
/*
 * Pack OIHW fp32 conv weights (state_dict layout, models/ddpm.py:30) and an optional fused
 * 1x1 residual weight [cout][rc][1][1] (models/ddpm.py:109) into the GEMM-B layout:
 *   kernel == DMME_CONV_TC      -> bf16 [cout][K]   (K-major rows, K = k*k*cin + rc, tap-major)
 *   kernel == DMME_CONV_GENERIC -> fp32 [K][cout]
 */
int dmme_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, const float* w_res, int rc,
                          void* packed, int kernel, void* stream);

/* layout helpers --------------------------------------------------------------------------- */
int dmme_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int act_dtype, void* stream);
int dmme_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int act_dtype, void* stream);
int dmme_upsample2x_nhwc(const void* src, void* dst, int n, int h, int w, int c, int act_dtype, void* stream);

/* convolution ------------------------------------------------------------------------------ */
int dmme_conv2d_fwd(const dmme_conv_desc* desc, void* stream);
/* 1 when dmme_conv2d_fwd would take the tcgen05 path for this descriptor */
int dmme_conv2d_uses_tc(const dmme_conv_desc* desc);
/* 1 when the kernel dmme_conv2d_fwd would run for this descriptor fills desc->stats */
int dmme_conv2d_writes_stats(const dmme_conv_desc* desc);
/* 1 when the kernel dmme_conv2d_fwd would run for this descriptor can apply desc->gn_ab (fused GroupNorm of the input) */
int dmme_conv2d_fuses_gn(const dmme_conv_desc* desc);
/* 1 when the kernel dmme_conv2d_fwd would run for this descriptor can apply desc->sampler in its epilogue */
int dmme_conv2d_fuses_sampler(const dmme_conv_desc* desc);
/* bytes of fp32 workspace (desc->splitk_ws) with which dmme_conv2d_fwd runs this descriptor split-K, 0 when it would not
 * (enough work units without splitting, or a shape / layout the split-K kernel does not take).  desc->out_norm[] is only
 * honoured on the split-K path. */
long long dmme_conv2d_splitk_workspace(const dmme_conv_desc* desc);
/* 1 when the kernel that runs `desc` WITHOUT a split-K workspace honours out_norm[] too: 3x3 convs on 8x8 maps with 128 /
 * 256 output channels on the transposed tcgen05 kernel, whose epilogue warps hold whole images (an image = two 32-pixel
 * chunks of one warp, a GroupNorm group = neighbouring lanes) and finish the consumers' GroupNorm(+SiLU) themselves */
int dmme_conv2d_epilogue_norm(const dmme_conv_desc* desc);
/* A/B switch: 0 = never split K, 1 = default (by the cost model), 2 = wherever the split-K kernel supports the shape */
void dmme_set_conv_splitk_mode(int mode);
/* A/B switch: 0 = the finishing pass of a split-K conv on 4x4 maps takes the block-per-slab kernel of the larger maps,
 * 1 = its warp-per-slab kernel where there are at least 2048 (image, 32-channel slab) units (default; same bits),
 * 2 = the warp-per-slab kernel always */
void dmme_set_splitk_finish_small(int mode);
/* A/B and test switch: 1 (default) = split-K convs on 4x4 maps reduce their K slices inside a thread-block cluster through
 * distributed shared memory and finish in the same launch (same bits as GEMM + finishing pass); 0 = partial tiles in the
 * workspace + finishing pass */
void dmme_set_conv_splitk_cluster(int mode);

/* chain of 3x3 convolutions (low-resolution ResBlocks) in one launch ------------------------ */
/*
 * One conv of a chain:  out = conv3x3( cat(src0, src1) ) + conv1x1( cat(res0, res1) ) + bias + temb[n or 0][:] + addend,
 * 256 output channels, stride 1, 4x4 or 8x8 maps, bf16 NHWC.  src0 / src1 are ALREADY NORMALISED operands
 * (norm_act_drop_conv models/ddpm.py:25-35 applies GroupNorm + SiLU before the conv): tensors in global memory, or -- src0 ==
 * NULL -- the operand the previous op of the chain left in shared memory (its out_norm[keep]).  res0 / res1 / addend are raw
 * tensors (ResBlock.residual models/ddpm.py:109,131) and may be `out` of an earlier op of the same chain.  The epilogue
 * writes the raw output (`out`, optional), its GroupNorm micro-group statistics (`stats`, optional, the format of
 * dmme_conv_desc.stats) and the GroupNorm(+SiLU) of up to two consumers (dmme_out_norm; `.out` may be NULL for the kept one).
 */
typedef struct dmme_chain_op {
  const void* src0; const void* src1; int c0, c1;
  const void* res0; const void* res1; int rc0, rc1;
  const void* weight;                 /* packed by dmme_pack_conv_weight(DMME_CONV_TC): bf16 [256][9 (c0 + c1) + rc0 + rc1] */
  const float* bias;
  const float* temb; int temb_rows, temb_ld;
  const void* addend;
  void* out;
  long long* stats;
  dmme_out_norm out_norm[2];
  int keep;                           /* index of the out_norm kept in shared memory as the next op's src0, -1: none */
} dmme_chain_op;
/*
 * Runs ops[0..nops) back to back in ONE persistent kernel (csrc/conv_chain.cu): every CTA owns whole images for the length
 * of the chain, so the GroupNorm between two convs is finished in the epilogue and its result stays in shared memory as
 * the next conv's tensor-core operand; only the weights stream.  Replaces the per-conv launches (split-K GEMM + finishing
 * pass, or conv + GroupNorm) of consecutive ResBlocks at the 8x8 / 4x4 levels (UNet.forward models/ddpm.py:295-313).
 * nops <= 16; n images of h x w (4x4 or 8x8).
 */
int dmme_conv_chain_fwd(const dmme_chain_op* ops, int nops, int n, int h, int w, void* stream);
/* 1 when dmme_conv_chain_fwd takes maps of this size with this many output channels */
int dmme_conv_chain_supported(int n, int h, int w, int cout);
/* A/B switch: images per CTA of the chain kernel (0 = cost model) */
void dmme_set_conv_chain_ipc(int ipc);
/* debugging: int64[6 * 1024] device buffer receiving CTA 0's per-role clock64 timestamps (tools/trace_chain.py) */
void dmme_debug_set_chain_trace(long long* buf);
/* NOTE: the chain kernel is a measured negative result (slower than the per-conv launches at every batch): it
 * is compiled only with -DDMME_EXPERIMENTAL (`make EXPERIMENTAL=1`); in the shipped build dmme_conv_chain_supported
 * answers 0 and dmme_conv_chain_fwd returns DMME_E_UNSUPPORTED. */

/* GroupNorm (+ scale/shift) (+ SiLU) (+ channel dropout mask) -------------------------------- */
/*
 * y = [mask[n][c] *] [silu]( [ (1 + scale[n or 0][c]) * ] GN_{groups,eps}(cat(src0,src1)) * gamma + beta [+ shift] )
 * Replaces: nn.GroupNorm -> nn.SiLU -> nn.Dropout2d of norm_act_drop_conv models/ddpm.py:25-35,
 * Attention.norm models/ddpm.py:73 and the scale-shift norm of models/iddpm.py:119.
 */
int dmme_groupnorm_fwd(const void* src0, const void* src1, int c0, int c1, int n, int hw, int groups, float eps,
                       const float* gamma, const float* beta, const float* scale, const float* shift,
                       int ss_rows, int ss_ld, const float* chan_mask, int apply_silu, void* out,
                       int act_dtype, const long long* stats0, const long long* stats1, void* stream);
/* (a, b) of y = a * x + b per (image, channel) from the producers' statistics, for dmme_conv_desc.gn_ab */
int dmme_groupnorm_coeff(const long long* stats0, const long long* stats1, int c0, int c1, int n, int hw, int groups,
                         float eps, const float* gamma, const float* beta, const float* scale, const float* shift,
                         int ss_rows, int ss_ld, float* ab_out, void* stream);
/* stats0 / stats1: optional micro-group sums of src0 / src1 written by the producing convolution
 * (dmme_conv_desc.stats).  When every source has them, GroupNorm is a single streaming pass
 * (2 B read + 2 B written per element); otherwise the kernel reduces the statistics itself. */

/* self-attention core ---------------------------------------------------------------------- */
/*
 * out[b'][l][h'*dh + c] = softmax_j( scale * <q[b][l][h], k[b][j][h]> ) v[b][j][h][c]
 * q/k/v element (b, l, h, c) lives at ptr[b*batch_stride + l*row_stride + h*head_stride + c]
 * (v_transposed = 1: v element at v[b*batch_stride_v + (h*dh + c)*L + l]).
 * head_batch_swap = 1 reproduces the "(b head)" -> "(head b)" regrouping of
 * MultiHeadAttention.forward_attention models/iddpm.py:38-46; 0 is Attention models/ddpm.py:54-63.
 * kernel: DMME_CONV_AUTO / DMME_CONV_GENERIC / DMME_CONV_TC (same selector values as the convolution).
 */
int dmme_attention_fwd(const void* q, const void* k, const void* v, long long batch_stride, int row_stride,
                       int head_stride, int v_transposed, long long v_batch_stride, int n, int heads, int L,
                       int dh, float scale, int head_batch_swap, void* out, int act_dtype, int kernel,
                       void* stream);
/* 1 when DMME_CONV_AUTO would run the fused tcgen05 kernel (bf16, single head, L = 256, dh in {64,128,192,256},
 * dense [n][L][dh] q/k and transposed v) instead of the generic CUDA-core kernel */
int dmme_attention_uses_tc(long long batch_stride, int row_stride, int v_transposed, long long v_batch_stride,
                           int heads, int L, int dh, int head_batch_swap, int act_dtype);

/*
 * The whole attention block of Attention.forward (models/ddpm.py:54-75) in one launch:
 *   out = x + proj( softmax(scale * q k^T) v ),  [q | k | v] = qkv_proj(GroupNorm(x))
 * x / out: NHWC act_dtype [n][L][c]; gn_ab: the (a, b) pairs of the block's GroupNorm from dmme_groupnorm_coeff
 * ([n][c][2], no SiLU) -- or NULL with stats_in, see below; wqkv [3c][c] / wproj [c][c]: the 1x1 conv weights packed by dmme_pack_conv_weight (bf16, K
 * contiguous); bias_qkv [3c], bias_proj [c] fp32; stats: optional micro-group sums of `out` (as dmme_conv_desc.stats).
 * Nothing between x and out is written to global memory.  Supported (ask first): bf16, heads = 1, and L = 256 with c = 256
 * or 128 (the five 16x16 sites of the default DDPM UNet, configs/ddpm/cifar10.yaml: a cluster of two CTAs per image) or
 * L = 16 with c = 256 (its 4x4 middle block: eight images per CTA); other shapes take the qkv conv + dmme_attention_fwd +
 * proj conv path.
 */
int dmme_attention_block_supported(int heads, int L, int c, int act_dtype);
int dmme_attention_block_fwd(const void* x, const float* gn_ab, const long long* stats_in, const float* gamma,
                             const float* beta, int groups, float eps, const void* wqkv, const float* bias_qkv,
                             const void* wproj, const float* bias_proj, int n, int heads, int L, int c, float scale,
                             void* out, long long* stats, int act_dtype, void* stream);
/* The block's GroupNorm: either gn_ab (coefficient pairs from dmme_groupnorm_coeff) or, with gn_ab = NULL, stats_in (the
 * micro-group sums its producer wrote, dmme_conv_desc.stats) + gamma / beta / groups / eps: the kernel then forms the
 * coefficients itself (same arithmetic, same bits) and the coefficient launch disappears. */

/* timestep embedding ----------------------------------------------------------------------- */
/*
 * emb = SiLU(W2 SiLU(W1 [sin(t f), cos(t f)] + b1) + b2)   (UNet.condition models/ddpm.py:211-217,338-349)
 * t: int64 device pointer [rows]; freq: the persistent `condition.0.embeddings` buffer [half];
 * scratch: fp32 [rows][emb_dim] workspace for the hidden layer.
 */
int dmme_temb_mlp_fwd(const int64_t* t, int rows, const float* freq, int half, const float* w1, const float* b1,
                      const float* w2, const float* b2, int emb_dim, float* scratch, float* emb_out, void* stream);
/* out[rows][total] = emb[rows][emb_dim] Wcat[total][emb_dim]^T + bcat: all ResBlock.condition Linears
 * (models/ddpm.py:101-104, models/iddpm.py:89-92) batched into one launch. */
int dmme_temb_proj_fwd(const float* emb, int rows, int emb_dim, const float* wcat, const float* bcat, int total,
                       float* out, void* stream);

/* sampler updates (image space, NCHW fp32, elementwise) ------------------------------------- */
/*
 * tables: fp32 device arrays of length table_len = T+1 (beta, alpha, alpha_bar as registered by
 * DDPM.__init__ diffusion_models/ddpm.py:41-51); t_ptr: int64 device scalar holding the current t.  Table indices
 * (t, t-1, tau[i], tau[i-1]) wrap when negative like the reference's torch indexing and are clamped to the table
 * otherwise (the reference raises IndexError for t > T; the Python wrappers do that on the host in eager mode).
 * noise: fp32 tensor of standard normals (NULL: draw Philox4x32-10 normals from (seed, *t_ptr, element index)).
 * noise_offset: index of x[0] inside the whole (unsharded) sample batch, a multiple of 4 -- a rank that owns images
 * [i0, i1) passes i0 * C*H*W and draws exactly the noise the single-GPU run draws for those images.
 * ddpm:  x <- where(t==1, mean, mean + sqrt(beta_t) z), mean = 1/sqrt(alpha_t) (x - beta_t/sqrt(1-abar_t) eps)
 *        (diffusion_models/ddpm.py:83-111, equations/ddpm/ddpm.py:44-72)
 */
int dmme_ddpm_step(float* x, const float* eps, const float* noise, const float* beta, const float* alpha,
                   const float* alpha_bar, const int64_t* t_ptr, int table_len, long long numel, unsigned long long seed,
                   unsigned long long noise_offset, void* stream);
/* ddim (as written in equations/ddim/ddim.py:52-57): x0 = (x - sqrt(1-abar_i) eps)/sqrt(abar_prev); x <- sqrt(abar_prev) x0
 * i_ptr: int64 device scalar with the sub-sequence index i; tau: int64 [S+1]. */
int dmme_ddim_step(float* x, const float* eps, const float* alpha_bar, const int64_t* tau, const int64_t* i_ptr,
                   int table_len, int tau_len, long long numel, void* stream);
/* iddpm learned variance (diffusion_models/iddpm.py:118-164, equations/iddpm/losses.py:34-37):
 * model_out NCHW fp32 [n][2*c][hw]: first c channels eps, last c channels v. */
int dmme_iddpm_step(float* x, const float* model_out, const float* noise, const float* beta, const float* alpha,
                    const float* alpha_bar, const int64_t* t_ptr, int table_len, int n, int c, int hw,
                    unsigned long long seed, unsigned long long noise_offset, void* stream);
/* writes tau[*i_ptr] into *t_out (DDIM: the model is evaluated at tau_i) */
int dmme_gather_i64(const int64_t* table, const int64_t* idx_ptr, int64_t* out, void* stream);
/* *value += delta: advances the device-resident step counter between graph replays
 * (the host loop `for t in range(T, 0, -1)` of DDPM.generate diffusion_models/ddpm.py:130) */
int dmme_add_i64(int64_t* value, int64_t delta, void* stream);
/* standard normals from Philox4x32-10 keyed by (seed, stream_id): used for x_T and per-step z */
int dmme_philox_normal(float* out, long long numel, unsigned long long seed, unsigned long long stream_id,
                       unsigned long long noise_offset, void* stream);

/* ============================================================================================
 * Backward pass.  The reference obtains it from autograd through ATen after
 * DDPM.training_step (diffusion_models/ddpm.py:53-81) / IDDPM.training_step
 * (diffusion_models/iddpm.py:62-116); here each backward op is an explicit kernel.
 * ============================================================================================ */

/* Data-gradient weights of a convolution: the gradient w.r.t. input channels [ci_off, ci_off + ci_cnt) of
 * conv(w_oihw) is dmme_conv2d_fwd(grad_out) with these weights (cout' = ci_cnt, cin' = cout; spatially flipped).
 * Stride-2 convs (models/ddpm.py:147) additionally set desc.upsample = 2 on the grad_out source. */
int dmme_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin, int ksize, int ci_off, int ci_cnt,
                                void* packed, int kernel, void* stream);

/*
 * All tensor-core (bf16, [rows][K]) weight packs of a model in one launch: items_dev is a DEVICE array of n_items entries,
 * each the arguments of one dmme_pack_conv_weight (dgrad = 0) or dmme_pack_conv_weight_dgrad (dgrad = 1) call with
 * kernel = DMME_CONV_TC; first_block = the sum of ceil(elements / dmme_pack_block_elems()) over the preceding entries,
 * total_blocks = that sum over all entries.  Same values as the per-weight calls.  Used by the graph-captured training step,
 * which re-packs every weight after each optimizer step.
 */
typedef struct dmme_pack_item {
  const float* w;        /* OIHW fp32 */
  const float* wres;     /* optional fused 1x1 residual weight [cout][rc] (dgrad = 0) */
  void* packed;          /* bf16 destination */
  int cout, cin, ksize, rc;
  int dgrad, ci_off, ci_cnt, reserved;
  long long first_block;
} dmme_pack_item;
int dmme_pack_block_elems(void);
int dmme_pack_conv_weights_batch(const dmme_pack_item* items_dev, int n_items, long long total_blocks, void* stream);

/* Weight, fused-residual-weight and bias gradients of the forward call described by `fwd` (sources and geometry;
 * its weight/out fields are ignored).  grad_out: NHWC act dtype [n][ho][wo][cout] (NCHW fp32 when fwd->out_layout
 * is DMME_OUT_NCHW_F32).  dweight: OIHW fp32 [cout][c0+c1][k][k]; dweight_res: [cout][rc0+rc1] or NULL; dbias:
 * [cout] or NULL.  Results are written (not accumulated).  Deterministic two-stage reduction through `workspace`
 * (dmme_conv2d_wgrad_workspace bytes). */
long long dmme_conv2d_wgrad_workspace(const dmme_conv_desc* fwd);
/* 1 when dmme_conv2d_wgrad takes the tcgen05 path (bf16 NHWC, stride 1, cout % 128 == 0, channels % 64 == 0, fwd->kernel
 * != DMME_CONV_GENERIC); the CUDA-core kernel handles everything else */
int dmme_conv2d_wgrad_uses_tc(const dmme_conv_desc* fwd);
int dmme_conv2d_wgrad(const dmme_conv_desc* fwd, const void* grad_out, float* dweight, float* dweight_res,
                      float* dbias, void* workspace, long long workspace_bytes, void* stream);

/* Backward of dmme_groupnorm_fwd.  gin0/gin1: gradients w.r.t. src0/src1 (act dtype; NULL: not needed);
 * add0/add1: optional tensors added into gin0/gin1 (gradient accumulation of skip / residual branches);
 * dgamma/dbeta [C] written; dscale/dshift [n][dss_ld] written when given (IDDPM scale-shift; always one row per image,
 * also when the forward broadcast a single scale/shift row -- the caller then sums the rows);
 * sums: fp32 workspace [n][C][2]. */
int dmme_groupnorm_bwd(const void* grad_out, const void* src0, const void* src1, int c0, int c1, int n, int hw,
                       int groups, float eps, const float* gamma, const float* beta, const float* scale,
                       const float* shift, int ss_rows, int ss_ld, const float* chan_mask, int apply_silu,
                       void* gin0, void* gin1, const void* add0, const void* add1, float* dgamma, float* dbeta,
                       float* dscale, float* dshift, int dss_ld, float* sums, int act_dtype, void* stream);

/* Backward of dmme_attention_fwd for q/k/v stored as strided views of one tensor (v not transposed):
 * dq/dk/dv use the same strides as q/k/v.  dout has the layout of the forward output [n][L][heads*dh].
 * p_saved: the softmax matrix kept by dmme_attention_fwd_train, or NULL to recompute it. */
long long dmme_attention_bwd_workspace(int n, int heads, int L, int dh);
int dmme_attention_bwd(const void* q, const void* k, const void* v, long long batch_stride, int row_stride,
                       int head_stride, int n, int heads, int L, int dh, float scale, int head_batch_swap,
                       const void* dout, void* dq, void* dk, void* dv, int act_dtype, const float* p_saved,
                       void* workspace, long long workspace_bytes, void* stream);
/* Training-mode forward of the same attention core that keeps the softmax matrix for the backward pass:
 * p_out fp32 [n*heads][L][L] (pass it to dmme_attention_bwd as p_saved), o_tmp fp32 [n*heads][L][dh] scratch. */
int dmme_attention_fwd_train(const void* q, const void* k, const void* v, long long batch_stride, int row_stride,
                             int head_stride, int n, int heads, int L, int dh, float scale, int head_batch_swap,
                             void* out, int act_dtype, float* p_out, float* o_tmp, void* stream);

/*
 * Fused tcgen05 backward of the multi-head attention core (MultiHeadAttention.forward_attention models/iddpm.py:36-59) for
 * the packed qkv layout [n][L][heads][q | k | v][dh] with dh = 64 or 32 and L = 256 or 64, bf16: one CTA per (image, head) -- two
 * images per CTA at L = 64 -- recomputes the softmax from Q and K, so neither the forward's softmax matrix nor any other
 * L x L matrix touches global memory.  out / dout: the forward output and its gradient, [n][L][heads * dh] at the
 * "(b head) -> (head b)" position when head_batch_swap; dqkv: gradient of the packed tensor (every element written).
 */
/* A/B switch: waves of CTAs the tcgen05 weight-gradient kernel slices the pixel axis for (0 = exactly one wave, rounded down: the default; n > 0: n waves) */
void dmme_set_wgrad_waves(int waves);
int dmme_attention_bwd_fused_supported(int heads, int L, int dh, int act_dtype);
int dmme_attention_bwd_fused(const void* qkv, const void* out, const void* dout, void* dqkv, int n, int heads, int L, int dh,
                             float scale, int head_batch_swap, int act_dtype, void* stream);

/* Backward of dmme_temb_mlp_fwd + dmme_temb_proj_fwd.  hidden/emb: the forward's scratch / emb_out;
 * d_all [rows][total]: gradient of the batched projection output.  All parameter gradients are written. */
long long dmme_temb_bwd_workspace(int rows, int half, int emb_dim);
int dmme_temb_bwd(const int64_t* t, int rows, const float* freq, int half, const float* w1, const float* b1,
                  const float* w2, const float* b2, int emb_dim, const float* hidden, const float* emb,
                  const float* wcat, int total, const float* d_all, float* dw1, float* db1, float* dw2, float* db2,
                  float* dwcat, float* dbcat, void* workspace, long long workspace_bytes, int bf16_mma, void* stream);
/* bf16_mma = 1 (bf16 training mode): the two products over the batched projection run on the tensor cores with operands
 * rounded to bf16 (fp32 accumulation); 0: fp32 FFMA throughout */

/* C[b](i,j) = alpha * sum_k A[b](i,k) B[b](k,j) (+ C when accumulate); b = bo*heads + h; element strides per operand
 * (outer batch, head, row, column); dtypes DMME_BF16 / DMME_F32.  CUDA-core product used by the attention and
 * timestep-MLP backward passes. */
int dmme_gemm_strided(const void* a, int a_dtype, long long a_bo, long long a_h, long long a_r, long long a_c,
                      const void* b, int b_dtype, long long b_bo, long long b_h, long long b_r, long long b_c,
                      void* c, int c_dtype, long long c_bo, long long c_h, long long c_r, long long c_c, int M,
                      int N, int K, int outer, int heads, float alpha, int accumulate, void* stream);

/* glue */
/* dst = a + b (act dtype; dst may alias a or b): gradient accumulation where no kernel can fuse it */
int dmme_add(void* dst, const void* a, const void* b, long long numel, int act_dtype, void* stream);
/* out[n][c] = sum over pixels of g[n][px][c]: gradient of the broadcast timestep-embedding add (models/ddpm.py:129) */
int dmme_pixel_sum(const void* g, int n, int hw, int c, float* out, long long out_ld, int act_dtype, void* stream);
/* backward of nearest x2 upsampling (models/ddpm.py:161): out[n][y][x][c] = sum of the 2x2 block of g [n][2h][2w][c] */
int dmme_pool2x_sum_nhwc(const void* g, void* out, int n, int h, int w, int c, int act_dtype, void* stream);
/* dst[n][2y][2x] = src[n][y][x], zeros elsewhere: turns the gradients of a stride-2 conv (models/ddpm.py:147) into the
 * stride-1 data / weight gradients of the dilated grad_out (tensor-core kernels) */
int dmme_dilate2x_nhwc(const void* src, void* dst, int n, int h, int w, int c, int act_dtype, void* stream);
int dmme_colsum_f32(const float* in, int rows, int cols, long long ld, float* out, int accumulate, void* stream);

/* L_simple (equations/ddpm/losses.py:5-13): loss_out[0] = mean((eps - noise)^2);
 * d_eps (optional) = grad_scale * 2 (eps - noise) / numel.  partial: fp32 workspace of >= 1024 floats. */
int dmme_mse_loss(const float* eps, const float* noise, long long numel, float grad_scale, float* d_eps,
                  float* loss_out, float* partial, void* stream);
/* IDDPM hybrid / VLB loss and its gradient w.r.t. the network output in one pass (IDDPM.training_step
 * diffusion_models/iddpm.py:62-116, forward_model :150-164, equations/iddpm/losses.py:8-90).
 * model_out NCHW fp32 [n][2c][hw]; x_t, x_0 [n][c][hw]; t int64 [n]; tables of length T+1.
 * loss_out[0] = w_simple L_simple + w_vlb L_vlb, [1] = L_simple, [2] = L_vlb; d_out optional [n][2c][hw];
 * partial: fp32 workspace of >= 2048 floats. */
int dmme_iddpm_loss(const float* model_out, const float* x_t, const float* x_0, const int64_t* t, const float* beta,
                    const float* alpha, const float* alpha_bar, int n, int c, int hw, float w_simple, float w_vlb,
                    float grad_scale, float* d_out, float* loss_out, float* partial, void* stream);

/* image-space tail ------------------------------------------------------------------------- */
/*
 * y = clip((x + 1) / 2, 0, 1) (common/norm.py:9-11 `denorm`), written as fp32 (out_f32) and / or as uint8 round(255 y)
 * (out_u8); either output may be NULL.  Replaces the denorm of every snapshot in GenerateImage.generate_img
 * (callbacks/generate.py:64-90) and of LitDDPM.test_step (lit_modules/ddpm.py:97-100).  fp32 output is bit-exact.
 */
int dmme_denorm(const float* x, float* out_f32, uint8_t* out_u8, long long numel, void* stream);

/* fused optimizer tail ---------------------------------------------------------------------- */
/*
 * One training step's parameter update for ALL tensors in two launches: global gradient-norm clip
 * (Lightning gradient_clip_val, configs/ddpm/cifar10.yaml:24 = torch.nn.utils.clip_grad_norm_), Adam with torch.optim.Adam's
 * default arithmetic (lit_modules/ddpm.py:130), the caller's WarmupLR-scaled learning rate (lr_scheduler/warmup.py:10-19)
 * and the EMA update ema = d*ema + (1-d)*w (callbacks/ema.py:169-176).
 * table: device array of `count` entries { float* w; const float* g; float* m; float* v; float* ema (or NULL);
 * long long numel; long long first_item; } where first_item = running sum of ceil(numel / dmme_optim_chunk()).
 * items = total work items; partial: >= grid floats of scratch; norm_out (optional): the pre-clip global L2 norm.
 * max_norm <= 0 disables clipping; step is the 1-based optimizer step (bias correction).
 */
int dmme_optim_table_entry_bytes(void);
int dmme_optim_chunk(void);
int dmme_adam_ema_step(const void* table, int count, long long items, double lr, double beta1, double beta2,
                       double eps, int step, double max_norm, double ema_decay, float* partial, int grid,
                       float* norm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMME_B200_H */
