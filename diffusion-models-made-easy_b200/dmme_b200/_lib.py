"""ctypes binding of the C-ABI library ``libdmme_b200.so`` (declared in ``include/dmme_b200.h``).

The library is the product: there is no PyTorch or CPU fallback behind these calls.  Loading fails
loudly when the shared object is missing, and every wrapper raises ``RuntimeError`` carrying
``dmme_last_error()`` when a launch is refused, mirroring how the reference surfaces errors as
Python exceptions (e.g. ``NotImplementedError`` in diffusion_models/ddim.py:50-51).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMME_LIB_PATH") or os.path.join(_HERE, "libdmme_b200.so")  # override: A/B builds only

BF16, F32 = 0, 1
IN_NHWC, IN_NCHW_F32 = 0, 1
OUT_NHWC, OUT_NCHW_F32, OUT_QKV = 0, 1, 2
CONV_AUTO, CONV_GENERIC, CONV_TC, CONV_HALO = 0, 1, 2, 3

# every symbol include/dmme_b200.h declares (checked by tests/test_abi.py without a GPU)
EXPORTS = (
    "dmme_abi_version", "dmme_has_experimental", "dmme_last_error", "dmme_launch_count", "dmme_reset_launch_count",
    "dmme_pack_conv_weight", "dmme_nchw_to_nhwc", "dmme_nhwc_to_nchw", "dmme_upsample2x_nhwc",
    "dmme_conv2d_fwd", "dmme_conv2d_uses_tc", "dmme_conv2d_writes_stats", "dmme_conv2d_fuses_gn", "dmme_conv2d_splitk_workspace", "dmme_conv2d_epilogue_norm", "dmme_conv2d_fuses_sampler", "dmme_set_conv_splitk_mode", "dmme_set_splitk_finish_small", "dmme_set_conv_splitk_cluster", "dmme_conv_chain_fwd", "dmme_conv_chain_supported", "dmme_set_conv_chain_ipc", "dmme_debug_set_chain_trace", "dmme_groupnorm_coeff", "dmme_groupnorm_fwd", "dmme_attention_fwd", "dmme_attention_uses_tc", "dmme_attention_block_fwd", "dmme_attention_block_supported",
    "dmme_temb_mlp_fwd", "dmme_temb_proj_fwd", "dmme_ddpm_step", "dmme_ddim_step", "dmme_iddpm_step",
    "dmme_gather_i64", "dmme_add_i64", "dmme_philox_normal", "dmme_set_conv_halo_mode", "dmme_get_conv_halo_mode", "dmme_set_conv_halo_multicast",
    "dmme_set_conv_tct_mode", "dmme_get_conv_tct_mode", "dmme_set_conv_out_tc_mode", "dmme_set_conv_in_tc_mode", "dmme_set_conv_pair_mode", "dmme_set_attn_mma_mode",
    "dmme_denorm", "dmme_optim_table_entry_bytes", "dmme_optim_chunk", "dmme_adam_ema_step",
    "dmme_pack_conv_weight_dgrad", "dmme_pack_block_elems", "dmme_pack_conv_weights_batch", "dmme_conv2d_wgrad_workspace", "dmme_conv2d_wgrad", "dmme_conv2d_wgrad_uses_tc", "dmme_groupnorm_bwd",
    "dmme_attention_bwd_workspace", "dmme_attention_bwd", "dmme_attention_bwd_fused", "dmme_attention_bwd_fused_supported", "dmme_set_wgrad_waves", "dmme_attention_fwd_train", "dmme_temb_bwd_workspace", "dmme_temb_bwd",
    "dmme_gemm_strided", "dmme_add", "dmme_pixel_sum", "dmme_pool2x_sum_nhwc", "dmme_dilate2x_nhwc", "dmme_colsum_f32", "dmme_mse_loss", "dmme_iddpm_loss",
)


class PackItem(C.Structure):
    """Mirror of ``struct dmme_pack_item``."""

    _fields_ = [
        ("w", C.c_void_p), ("wres", C.c_void_p), ("packed", C.c_void_p),
        ("cout", C.c_int), ("cin", C.c_int), ("ksize", C.c_int), ("rc", C.c_int),
        ("dgrad", C.c_int), ("ci_off", C.c_int), ("ci_cnt", C.c_int), ("reserved", C.c_int),
        ("first_block", C.c_longlong),
    ]


class OutNorm(C.Structure):
    """Mirror of ``struct dmme_out_norm``."""

    _fields_ = [
        ("out", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("ss_rows", C.c_int), ("ss_ld", C.c_int), ("cpg", C.c_int), ("silu", C.c_int), ("eps", C.c_float),
    ]


SAMPLER_NONE, SAMPLER_DDPM, SAMPLER_DDIM, SAMPLER_IDDPM = 0, 1, 2, 3


class SamplerEpilogue(C.Structure):
    """Mirror of ``struct dmme_sampler_epilogue``."""

    _fields_ = [
        ("kind", C.c_int), ("x", C.c_void_p), ("noise", C.c_void_p),
        ("beta", C.c_void_p), ("alpha", C.c_void_p), ("alpha_bar", C.c_void_p),
        ("t_ptr", C.c_void_p), ("tau", C.c_void_p), ("table_len", C.c_int), ("tau_len", C.c_int),
        ("seed", C.c_ulonglong), ("noise_offset", C.c_ulonglong),
    ]


class ConvDesc(C.Structure):
    """Mirror of ``struct dmme_conv_desc``."""

    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
        ("res0", C.c_void_p), ("res1", C.c_void_p), ("rc0", C.c_int), ("rc1", C.c_int),
        ("n", C.c_int), ("h_in", C.c_int), ("w_in", C.c_int),
        ("ksize", C.c_int), ("stride", C.c_int), ("upsample", C.c_int), ("cout", C.c_int),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("temb", C.c_void_p),
        ("temb_rows", C.c_int), ("temb_ld", C.c_int),
        ("addend", C.c_void_p), ("out", C.c_void_p), ("out2", C.c_void_p), ("out3", C.c_void_p),
        ("stats", C.c_void_p),
        ("in_layout", C.c_int), ("out_layout", C.c_int), ("act_dtype", C.c_int), ("kernel", C.c_int),
        ("gn_ab", C.c_void_p), ("gn_silu", C.c_int),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_longlong), ("out_norm", OutNorm * 2),
        ("sampler", C.POINTER(SamplerEpilogue)),
    ]


class ChainOp(C.Structure):
    """Mirror of ``struct dmme_chain_op``."""

    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
        ("res0", C.c_void_p), ("res1", C.c_void_p), ("rc0", C.c_int), ("rc1", C.c_int),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("temb", C.c_void_p),
        ("temb_rows", C.c_int), ("temb_ld", C.c_int),
        ("addend", C.c_void_p), ("out", C.c_void_p), ("stats", C.c_void_p),
        ("out_norm", OutNorm * 2), ("keep", C.c_int),
    ]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dmme_b200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C diffusion-models-made-easy_b200/csrc`). There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    lib.dmme_last_error.restype = C.c_char_p
    lib.dmme_launch_count.restype = C.c_longlong
    lib.dmme_abi_version.restype = C.c_int
    vp, i, ll, ull, f = C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_float
    lib.dmme_pack_conv_weight.argtypes = [vp, i, i, i, vp, i, vp, i, vp]
    lib.dmme_nchw_to_nhwc.argtypes = [vp, vp, i, i, i, i, i, vp]
    lib.dmme_nhwc_to_nchw.argtypes = [vp, vp, i, i, i, i, i, vp]
    lib.dmme_upsample2x_nhwc.argtypes = [vp, vp, i, i, i, i, i, vp]
    lib.dmme_conv2d_fwd.argtypes = [C.POINTER(ConvDesc), vp]
    lib.dmme_conv2d_uses_tc.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_conv2d_writes_stats.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_groupnorm_fwd.argtypes = [vp, vp, i, i, i, i, i, f, vp, vp, vp, vp, i, i, vp, i, vp, i, vp, vp, vp]
    lib.dmme_groupnorm_coeff.argtypes = [vp, vp, i, i, i, i, i, f, vp, vp, vp, vp, i, i, vp, vp]
    lib.dmme_conv2d_fuses_gn.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_conv2d_fuses_sampler.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_conv2d_splitk_workspace.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_conv2d_epilogue_norm.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_pack_block_elems.argtypes = []
    lib.dmme_pack_conv_weights_batch.argtypes = [vp, i, ll, vp]
    lib.dmme_conv2d_splitk_workspace.restype = ll
    lib.dmme_set_conv_splitk_mode.argtypes = [i]
    lib.dmme_set_conv_splitk_mode.restype = None
    lib.dmme_set_splitk_finish_small.argtypes = [i]
    lib.dmme_set_splitk_finish_small.restype = None
    lib.dmme_set_conv_splitk_cluster.argtypes = [i]
    lib.dmme_set_conv_splitk_cluster.restype = None
    lib.dmme_conv_chain_fwd.argtypes = [C.POINTER(ChainOp), i, i, i, i, vp]
    lib.dmme_conv_chain_supported.argtypes = [i, i, i, i]
    lib.dmme_set_conv_chain_ipc.argtypes = [i]
    lib.dmme_set_conv_chain_ipc.restype = None
    lib.dmme_attention_fwd.argtypes = [vp, vp, vp, ll, i, i, i, ll, i, i, i, i, f, i, vp, i, i, vp]
    lib.dmme_attention_uses_tc.argtypes = [ll, i, i, ll, i, i, i, i, i]
    lib.dmme_attention_block_supported.argtypes = [i, i, i, i]
    lib.dmme_attention_block_fwd.argtypes = [vp, vp, vp, vp, vp, i, f, vp, vp, vp, vp, i, i, i, i, f, vp, vp, i, vp]
    lib.dmme_temb_mlp_fwd.argtypes = [vp, i, vp, i, vp, vp, vp, vp, i, vp, vp, vp]
    lib.dmme_temb_proj_fwd.argtypes = [vp, i, i, vp, vp, i, vp, vp]
    lib.dmme_ddpm_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, ll, ull, ull, vp]
    lib.dmme_ddim_step.argtypes = [vp, vp, vp, vp, vp, i, i, ll, vp]
    lib.dmme_iddpm_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, ull, ull, vp]
    lib.dmme_gather_i64.argtypes = [vp, vp, vp, vp]
    lib.dmme_add_i64.argtypes = [vp, C.c_int64, vp]
    lib.dmme_philox_normal.argtypes = [vp, ll, ull, ull, ull, vp]
    lib.dmme_set_conv_halo_mode.argtypes = [i]
    lib.dmme_set_conv_halo_multicast.argtypes = [i]
    lib.dmme_set_conv_halo_multicast.restype = None
    lib.dmme_set_conv_tct_mode.argtypes = [i]
    lib.dmme_set_conv_out_tc_mode.argtypes = [i]
    lib.dmme_set_conv_in_tc_mode.argtypes = [i]
    lib.dmme_set_conv_pair_mode.argtypes = [i]
    lib.dmme_set_attn_mma_mode.argtypes = [i]
    lib.dmme_denorm.argtypes = [vp, vp, vp, ll, vp]
    lib.dmme_adam_ema_step.argtypes = [vp, i, ll, C.c_double, C.c_double, C.c_double, C.c_double, i, C.c_double, C.c_double,
                                       vp, i, vp, vp]
    lib.dmme_pack_conv_weight_dgrad.argtypes = [vp, i, i, i, i, i, vp, i, vp]
    lib.dmme_conv2d_wgrad_workspace.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_conv2d_wgrad_workspace.restype = ll
    lib.dmme_conv2d_wgrad_uses_tc.argtypes = [C.POINTER(ConvDesc)]
    lib.dmme_conv2d_wgrad.argtypes = [C.POINTER(ConvDesc), vp, vp, vp, vp, vp, ll, vp]
    lib.dmme_groupnorm_bwd.argtypes = [vp, vp, vp, i, i, i, i, i, f, vp, vp, vp, vp, i, i, vp, i,
                                       vp, vp, vp, vp, vp, vp, vp, vp, i, vp, i, vp]
    lib.dmme_attention_bwd_workspace.argtypes = [i, i, i, i]
    lib.dmme_attention_bwd_workspace.restype = ll
    lib.dmme_attention_bwd.argtypes = [vp, vp, vp, ll, i, i, i, i, i, i, f, i, vp, vp, vp, vp, i, vp, vp, ll, vp]
    lib.dmme_attention_bwd_fused_supported.argtypes = [i, i, i, i]
    lib.dmme_attention_bwd_fused.argtypes = [vp, vp, vp, vp, i, i, i, i, f, i, i, vp]
    lib.dmme_attention_fwd_train.argtypes = [vp, vp, vp, ll, i, i, i, i, i, i, f, i, vp, i, vp, vp, vp]
    lib.dmme_temb_bwd_workspace.argtypes = [i, i, i]
    lib.dmme_temb_bwd_workspace.restype = ll
    lib.dmme_temb_bwd.argtypes = [vp, i, vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, ll, i, vp]
    lib.dmme_gemm_strided.argtypes = [vp, i, ll, ll, ll, ll, vp, i, ll, ll, ll, ll, vp, i, ll, ll, ll, ll,
                                      i, i, i, i, i, f, i, vp]
    lib.dmme_add.argtypes = [vp, vp, vp, ll, i, vp]
    lib.dmme_pixel_sum.argtypes = [vp, i, i, i, vp, ll, i, vp]
    lib.dmme_pool2x_sum_nhwc.argtypes = [vp, vp, i, i, i, i, i, vp]
    lib.dmme_dilate2x_nhwc.argtypes = [vp, vp, i, i, i, i, i, vp]
    lib.dmme_colsum_f32.argtypes = [vp, i, i, ll, vp, i, vp]
    lib.dmme_mse_loss.argtypes = [vp, vp, ll, f, vp, vp, vp, vp]
    lib.dmme_iddpm_loss.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, f, f, f, vp, vp, vp, vp]
    lib.dmme_set_conv_halo_mode.restype = None
    lib.dmme_set_conv_tct_mode.restype = None
    lib.dmme_set_conv_out_tc_mode.restype = None
    lib.dmme_set_conv_in_tc_mode.restype = None
    lib.dmme_set_conv_pair_mode.restype = None
    lib.dmme_set_attn_mma_mode.restype = None
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("dmme_abi_version",):
            pass
    mode = os.environ.get("DMME_TCT_MODE")  # A/B measurements only: 0 = never use the transposed tcgen05 conv kernel
    if mode:
        lib.dmme_set_conv_tct_mode(int(mode))
    mode = os.environ.get("DMME_SPLITK_MODE")  # A/B measurements only: 0 = never split K
    if mode:
        lib.dmme_set_conv_splitk_mode(int(mode))
    mode = os.environ.get("DMME_WGRAD_WAVES")  # A/B measurements only
    if mode:
        lib.dmme_set_wgrad_waves.argtypes = [i]
        lib.dmme_set_wgrad_waves.restype = None
        lib.dmme_set_wgrad_waves(int(mode))
    mode = os.environ.get("DMME_SPLITK_CLUSTER")  # A/B measurements only: 0 = GEMM + finishing pass at 4x4, 2 = bend the plan
    if mode:
        lib.dmme_set_conv_splitk_cluster(int(mode))
    mode = os.environ.get("DMME_SPLITK_FORCE_4X4")  # A/B measurements only: "np,split" for split-K on 4x4 maps
    if mode:
        lib.dmme_debug_force_splitk_4x4.argtypes = [i, i]
        lib.dmme_debug_force_splitk_4x4.restype = None
        lib.dmme_debug_force_splitk_4x4(*[int(v) for v in mode.split(",")])
    mode = os.environ.get("DMME_FINISH_SMALL")  # A/B measurements only: 0 = block-per-slab split-K finishing kernel at 4x4
    if mode:
        lib.dmme_set_splitk_finish_small(int(mode))
    mode = os.environ.get("DMME_CHAIN_IPC")  # A/B measurements only: images per CTA of the chain kernel
    if mode:
        lib.dmme_set_conv_chain_ipc(int(mode))
    mode = os.environ.get("DMME_PAIR_MODE")  # A/B measurements only: 0 = no cta_group::2 kernels
    if mode:
        lib.dmme_set_conv_pair_mode(int(mode))
    mode = os.environ.get("DMME_ATTN_MMA_MODE")  # A/B measurements only: 0 = multi-head attention on CUDA cores
    if mode:
        lib.dmme_set_attn_mma_mode(int(mode))
    mode = os.environ.get("DMME_HALO_MODE")  # A/B measurements only: dmme_set_conv_halo_mode value
    if mode:
        lib.dmme_set_conv_halo_mode(int(mode))
    mode = os.environ.get("DMME_HALO_MC")  # A/B measurements only: 0 = no weight multicast in the halo conv kernel
    if mode:
        lib.dmme_set_conv_halo_multicast(int(mode))
    mode = os.environ.get("DMME_OUT_TC_MODE")  # A/B measurements only: 0 = output conv on the FFMA kernel
    if mode:
        lib.dmme_set_conv_out_tc_mode(int(mode))
    mode = os.environ.get("DMME_IN_TC_MODE")  # A/B measurements only: 0 = input conv on the FFMA kernel
    if mode:
        lib.dmme_set_conv_in_tc_mode(int(mode))
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dmme_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"dmme_b200.{what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Raw device pointer of a tensor (None stays a NULL pointer)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    """The CUDA stream every launch goes to: torch's current stream of the CURRENT device (see callbacks/ema.py:273-296).
    The module-level entry points (UNet.forward, DDPM.sampling_step / generate / training_step, FusedAdamEMA.step) make the
    tensors' device current first (``on_device``), so a model on cuda:1 launches on cuda:1's stream whatever device the
    caller had selected."""
    return torch.cuda.current_stream().cuda_stream


def on_device(t: torch.Tensor):
    """Context manager: make ``t``'s CUDA device the current one (kernel launches, streams and the C side's per-device
    kernel attributes all follow the current device)."""
    if not t.is_cuda:
        raise RuntimeError("dmme_b200 kernels need CUDA tensors; there is no CPU path")
    return torch.cuda.device(t.device)


def act_code(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float32:
        return F32
    raise ValueError(f"activation dtype must be bfloat16 or float32, got {dtype}")


def require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dmme_b200 kernels need CUDA tensors; there is no CPU path")
        if t is not None and not t.is_contiguous():
            raise RuntimeError("dmme_b200 kernels need contiguous tensors")
        if t is not None and t.device.index != torch.cuda.current_device():
            raise RuntimeError(f"dmme_b200: tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}; "
                               "call through the module entry points or wrap the call in torch.cuda.device(tensor.device)")
