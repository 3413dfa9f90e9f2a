"""Spherical-harmonics colour evaluation (the first "next" row after the graded path).

Same interface as the reference's `spherical_harmonics` (gsplat/cuda/_wrapper.py:151-181, autograd node :1799-1831):
real SH up to degree 4 of the NORMALISED direction, coefficients [..., K, 3].  Forward and backward are CUDA kernels
behind the C ABI (`rs_sh_fwd` / `rs_sh_bwd`, csrc/sh.cu) reached through `_C.spherical_harmonics_fwd/bwd`, like every
other operator; CPU tensors are refused (RuntimeError), there is no CPU path.
Masked-out entries return 0 (the reference leaves them uninitialised, SphericalHarmonics.cpp:29).
"""
from typing import Optional

import torch
from torch import Tensor

from . import _C


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh_degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor]) -> Tensor:
        colors = _C.spherical_harmonics_fwd(sh_degree, dirs, coeffs, masks)
        ctx.save_for_backward(dirs, coeffs, masks)
        ctx.sh_degree = sh_degree
        ctx.num_bases = coeffs.shape[-2]
        return colors

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        dirs, coeffs, masks = ctx.saved_tensors
        compute_v_dirs = ctx.needs_input_grad[1]
        v_coeffs, v_dirs = _C.spherical_harmonics_bwd(ctx.num_bases, ctx.sh_degree, dirs, coeffs, masks,
                                                      v_colors.contiguous(), compute_v_dirs)
        return None, (v_dirs if compute_v_dirs else None), v_coeffs, None


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    assert 0 <= degrees_to_use <= 4, degrees_to_use
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    batch_dims = dirs.shape[:-1]
    assert dirs.shape == batch_dims + (3,), dirs.shape
    assert coeffs.dim() == len(batch_dims) + 2 and coeffs.shape[:-2] == batch_dims and coeffs.shape[-1] == 3, coeffs.shape
    if masks is not None:
        assert masks.shape == batch_dims, masks.shape
        masks = masks.contiguous()
    if not (dirs.is_cuda and coeffs.is_cuda):
        raise RuntimeError("spherical_harmonics: dirs and coeffs must be CUDA tensors (there is no CPU path)")
    return _SphericalHarmonics.apply(degrees_to_use, dirs.contiguous(), coeffs.contiguous(), masks)
