"""Shared helpers for the parity tests (tests may use torch freely; the product path may not)."""
import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def to_nhwc(x: torch.Tensor, dtype) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).contiguous().float()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()
