"""Operator wrappers + autograd nodes: the host-side mirror of the hot-path part of gsplat/cuda/_wrapper.py.

Same function names, argument meaning, defaults and error behaviour as the reference:
  fully_fused_projection   _wrapper.py:288-439   (+ optional `rigid=RigidPoses(...)`, fusing main.py:183-228)
  isect_tiles              _wrapper.py:442-517
  isect_offset_encode      _wrapper.py:520-540
  rasterize_to_pixels      _wrapper.py:543-675   (no channel padding needed: the kernels take any channel count)
  _FullyFusedProjection    _wrapper.py:1030-1160
  _RasterizeToPixels       _wrapper.py:1251-1378
They call `_C` (this package's stand-in for the reference's pybind module) with positional arguments, exactly as the
reference's `_make_lazy_cuda_func(name)(*args)` does (_wrapper.py:12-19).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from . import _C
from ._C import RigidPoses

_CAMERA_MODELS = {"pinhole": _C.PINHOLE, "ortho": _C.ORTHO, "fisheye": _C.FISHEYE, "ftheta": _C.FTHETA}


def _shared_rows(t: Tensor, n_trailing: int) -> Optional[Tensor]:
    """If `t` is a broadcast view whose leading (image) dims all have stride 0 (what torch.broadcast_to produces in
    rendering.py:446-448 / 481-485), return the underlying per-Gaussian tensor; else None."""
    n_lead = t.dim() - n_trailing
    if n_lead <= 0:
        return None
    if all(t.stride(d) == 0 or t.shape[d] == 1 for d in range(n_lead)) and any(
        t.shape[d] > 1 for d in range(n_lead)
    ):
        base = t[(0,) * n_lead]
        return base if base.is_contiguous() else None
    return None


def fully_fused_projection(
    means: Tensor,  # [..., N, 3]
    covars: Optional[Tensor],  # [..., N, 6] or None
    quats: Optional[Tensor],  # [..., N, 4] or None
    scales: Optional[Tensor],  # [..., N, 3] or None
    viewmats: Tensor,  # [..., C, 4, 4]
    Ks: Tensor,  # [..., C, 3, 3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    packed: bool = False,
    sparse_grad: bool = False,
    calc_compensations: bool = False,
    camera_model: Literal["pinhole", "ortho", "fisheye", "ftheta"] = "pinhole",
    opacities: Optional[Tensor] = None,  # [..., N] or None
    rigid: Optional[RigidPoses] = None,
):
    """Projects Gaussians to 2D (reference: _wrapper.py:288-439).

    packed=False returns (radii, means2d, depths, conics, compensations) of shape [..., C, N, ...];
    packed=True returns (batch_ids, camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations) over the
    nnz visible (camera, Gaussian) pairs, in the reference's row-major (batch, camera, gaussian) order.
    """
    batch_dims = means.shape[:-2]
    N = means.shape[-2]
    C = viewmats.shape[-3]
    assert means.shape == batch_dims + (N, 3), means.shape
    assert viewmats.shape == batch_dims + (C, 4, 4), viewmats.shape
    assert Ks.shape == batch_dims + (C, 3, 3), Ks.shape
    means = means.contiguous()
    if covars is not None:
        assert covars.shape == batch_dims + (N, 6), covars.shape
        covars = covars.contiguous()
    else:
        assert quats is not None, "covars or quats is required"
        assert scales is not None, "covars or scales is required"
        assert quats.shape == batch_dims + (N, 4), quats.shape
        assert scales.shape == batch_dims + (N, 3), scales.shape
        quats = quats.contiguous()
        scales = scales.contiguous()
    if sparse_grad:
        assert packed, "sparse_grad is only supported when packed is True"
        assert batch_dims == (), "sparse_grad does not support batch dimensions"
    if opacities is not None:
        assert opacities.shape == batch_dims + (N,), opacities.shape
        opacities = opacities.contiguous()
    assert (
        camera_model != "ftheta"
    ), "ftheta camera is only supported via UT, please set with_ut=True in the rasterization()"
    viewmats = viewmats.contiguous()
    Ks = Ks.contiguous()

    if packed:
        # COO rows of the visible (batch, camera, gaussian) triples, in row-major order: one projection pass that places its
        # rows by a decoupled look-back (reference: two passes + cumsum, csrc/ProjectionEWA3DGSPacked.cu:17-375)
        return _FullyFusedProjectionPacked.apply(
            means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
            sparse_grad, calc_compensations, camera_model, opacities, rigid,
        )
    return _FullyFusedProjection.apply(
        means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        calc_compensations, camera_model, opacities, rigid,
    )


@torch.no_grad()
def isect_tiles(
    means2d: Tensor,  # [..., N, 2] or [nnz, 2]
    radii: Tensor,  # [..., N, 2] or [nnz, 2]
    depths: Tensor,  # [..., N] or [nnz]
    tile_size: int,
    tile_width: int,
    tile_height: int,
    sort: bool = True,
    segmented: bool = False,
    packed: bool = False,
    n_images: Optional[int] = None,
    image_ids: Optional[Tensor] = None,
    gaussian_ids: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """Maps projected Gaussians to intersecting tiles (reference: _wrapper.py:442-517).

    Returns (tiles_per_gauss int32, isect_ids int64 [n_isects] = image | tile | depth bits, flatten_ids int32)."""
    if packed:
        nnz = means2d.size(0)
        assert means2d.shape == (nnz, 2), means2d.shape
        assert radii.shape == (nnz, 2), radii.shape
        assert depths.shape == (nnz,), depths.shape
        assert image_ids is not None, "image_ids is required if packed is True"
        assert gaussian_ids is not None, "gaussian_ids is required if packed is True"
        assert n_images is not None, "n_images is required if packed is True"
        image_ids = image_ids.contiguous()
        gaussian_ids = gaussian_ids.contiguous()
        I = n_images
    else:
        image_dims = means2d.shape[:-2]
        I = math.prod(image_dims)
        N = means2d.shape[-2]
        assert means2d.shape == image_dims + (N, 2), means2d.shape
        assert radii.shape == image_dims + (N, 2), radii.shape
        assert depths.shape == image_dims + (N,), depths.shape
    return _C.intersect_tile(
        means2d.contiguous(), radii.contiguous(), depths.contiguous(), image_ids, gaussian_ids, I, tile_size,
        tile_width, tile_height, sort, segmented,
    )


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, n_images: int, tile_width: int, tile_height: int) -> Tensor:
    """Encodes intersection ids to offsets [I, tile_height, tile_width] (reference: _wrapper.py:520-540)."""
    return _C.intersect_offset(isect_ids.contiguous(), n_images, tile_width, tile_height)


def rasterize_to_pixels(
    means2d: Tensor,  # [..., N, 2] or [nnz, 2]
    conics: Tensor,  # [..., N, 3] or [nnz, 3]
    colors: Tensor,  # [..., N, channels] or [nnz, channels]
    opacities: Tensor,  # [..., N] or [nnz]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [..., tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
    backgrounds: Optional[Tensor] = None,  # [..., channels]
    masks: Optional[Tensor] = None,  # [..., tile_height, tile_width]
    packed: bool = False,
    absgrad: bool = False,
) -> Tuple[Tensor, Tensor]:
    """Rasterizes Gaussians to pixels (reference: _wrapper.py:543-675).

    Returns (render_colors [..., H, W, channels], render_alphas [..., H, W, 1])."""
    image_dims = means2d.shape[:-2]
    channels = colors.shape[-1]
    if packed:
        # the reference takes image_dims from means2d here too, i.e. () for packed rows, and then rejects every
        # [..., C, channels] background (_wrapper.py:582, 598-599); the image dims of packed rows are those of the offsets
        image_dims = isect_offsets.shape[:-2]
        nnz = means2d.size(0)
        assert means2d.shape == (nnz, 2), means2d.shape
        assert conics.shape == (nnz, 3), conics.shape
        assert colors.shape[0] == nnz, colors.shape
        assert opacities.shape == (nnz,), opacities.shape
    else:
        N = means2d.size(-2)
        assert means2d.shape == image_dims + (N, 2), means2d.shape
        assert conics.shape == image_dims + (N, 3), conics.shape
        assert colors.shape == image_dims + (N, channels), colors.shape
        assert opacities.shape == image_dims + (N,), opacities.shape
    if backgrounds is not None:
        assert backgrounds.shape == image_dims + (channels,), backgrounds.shape
        backgrounds = backgrounds.contiguous()
    if masks is not None:
        assert masks.shape == isect_offsets.shape, masks.shape
        masks = masks.contiguous()
    if channels > 513 or channels == 0:
        raise ValueError(f"Unsupported number of color channels: {channels}")

    tile_height, tile_width = isect_offsets.shape[-2:]
    assert (
        tile_height * tile_size >= image_height
    ), f"Assert Failed: {tile_height} * {tile_size} >= {image_height}"
    assert (
        tile_width * tile_size >= image_width
    ), f"Assert Failed: {tile_width} * {tile_size} >= {image_width}"

    # Broadcast views (one colour / opacity row per Gaussian shared by every image) are consumed as such: the kernels
    # index them modulo N instead of reading a materialised [..., C, N, D] copy.
    shared_colors = None if packed else _shared_rows(colors, 2)
    shared_opacities = None if packed else _shared_rows(opacities, 1)
    return _RasterizeToPixels.apply(
        means2d.contiguous(),
        conics.contiguous(),
        shared_colors if shared_colors is not None else colors.contiguous(),
        shared_opacities if shared_opacities is not None else opacities.contiguous(),
        backgrounds,
        masks,
        image_width,
        image_height,
        tile_size,
        isect_offsets.contiguous(),
        flatten_ids.contiguous(),
        absgrad,
        shared_colors is not None,
        shared_opacities is not None,
    )


class _FullyFusedProjection(torch.autograd.Function):
    """Projects Gaussians to 2D (reference: _wrapper.py:1030-1160), with the rigid transform fused in."""

    @staticmethod
    def forward(ctx, means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                radius_clip, calc_compensations, camera_model, opacities, rigid):
        assert (
            camera_model != "ftheta"
        ), "ftheta camera is only supported via UT, please set with_ut=True in the rasterization()"
        camera_model_type = _CAMERA_MODELS[camera_model]
        radii, means2d, depths, conics, compensations = _C.projection_ewa_3dgs_fused_fwd(
            means, covars, quats, scales, opacities, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
            radius_clip, calc_compensations, camera_model_type, rigid,
        )
        if not calc_compensations:
            compensations = None
        ctx.save_for_backward(means, covars, quats, scales, viewmats, Ks, radii, conics, compensations)
        ctx.width = width
        ctx.height = height
        ctx.eps2d = eps2d
        ctx.camera_model_type = camera_model_type
        ctx.rigid = rigid
        ctx.mark_non_differentiable(radii)
        return radii, means2d, depths, conics, compensations

    @staticmethod
    def backward(ctx, v_radii, v_means2d, v_depths, v_conics, v_compensations):
        means, covars, quats, scales, viewmats, Ks, radii, conics, compensations = ctx.saved_tensors
        if v_compensations is not None:
            v_compensations = v_compensations.contiguous()
        v_means, v_covars, v_quats, v_scales, v_viewmats = _C.projection_ewa_3dgs_fused_bwd(
            means, covars, quats, scales, viewmats, Ks, ctx.width, ctx.height, ctx.eps2d, ctx.camera_model_type,
            radii, conics, compensations, v_means2d.contiguous(), v_depths.contiguous(), v_conics.contiguous(),
            v_compensations, ctx.needs_input_grad[4], ctx.rigid,
        )
        if not ctx.needs_input_grad[0]:
            v_means = None
        if not ctx.needs_input_grad[1]:
            v_covars = None
        if not ctx.needs_input_grad[2]:
            v_quats = None
        if not ctx.needs_input_grad[3]:
            v_scales = None
        if not ctx.needs_input_grad[4]:
            v_viewmats = None
        return (v_means, v_covars, v_quats, v_scales, v_viewmats) + (None,) * 11


class _FullyFusedProjectionPacked(torch.autograd.Function):
    """Projects Gaussians to 2D, packed rows (reference: _wrapper.py:1579-1796), with the rigid transform fused in."""

    @staticmethod
    def forward(ctx, means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                radius_clip, sparse_grad, calc_compensations, camera_model, opacities, rigid):
        assert (
            camera_model != "ftheta"
        ), "ftheta camera is only supported via UT, please set with_ut=True in the rasterization()"
        camera_model_type = _CAMERA_MODELS[camera_model]
        (_indptr, batch_ids, camera_ids, gaussian_ids, radii, means2d, depths, conics,
         compensations) = _C.projection_ewa_3dgs_packed_fwd(
            means, covars, quats, scales, opacities, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
            radius_clip, calc_compensations, camera_model_type, rigid,
        )
        if not calc_compensations:
            compensations = None
        ctx.save_for_backward(batch_ids, camera_ids, gaussian_ids, means, covars, quats, scales, viewmats, Ks, conics,
                              compensations)
        ctx.width = width
        ctx.height = height
        ctx.eps2d = eps2d
        ctx.sparse_grad = sparse_grad
        ctx.camera_model_type = camera_model_type
        ctx.rigid = rigid
        ctx.mark_non_differentiable(batch_ids, camera_ids, gaussian_ids, radii)
        return batch_ids, camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations

    @staticmethod
    def backward(ctx, v_batch_ids, v_camera_ids, v_gaussian_ids, v_radii, v_means2d, v_depths, v_conics,
                 v_compensations):
        (batch_ids, camera_ids, gaussian_ids, means, covars, quats, scales, viewmats, Ks, conics,
         compensations) = ctx.saved_tensors
        sparse_grad = ctx.sparse_grad
        if v_compensations is not None:
            v_compensations = v_compensations.contiguous()
        v_means, v_covars, v_quats, v_scales, v_viewmats = _C.projection_ewa_3dgs_packed_bwd(
            means, covars, quats, scales, viewmats, Ks, ctx.width, ctx.height, ctx.eps2d, ctx.camera_model_type,
            batch_ids, camera_ids, gaussian_ids, conics, compensations, v_means2d.contiguous(), v_depths.contiguous(),
            v_conics.contiguous(), v_compensations, ctx.needs_input_grad[4], sparse_grad, ctx.rigid,
        )

        def finish(needed: bool, values: Optional[Tensor], like: Optional[Tensor]):
            if not needed or values is None:
                return None
            if sparse_grad:  # [nnz, D] rows -> sparse COO over the Gaussian axis (_wrapper.py:1726-1772)
                return torch.sparse_coo_tensor(indices=gaussian_ids[None], values=values, size=like.shape,
                                               is_coalesced=len(viewmats) == 1)
            return values

        return (
            finish(ctx.needs_input_grad[0], v_means, means),
            finish(ctx.needs_input_grad[1], v_covars, covars),
            finish(ctx.needs_input_grad[2], v_quats, quats),
            finish(ctx.needs_input_grad[3], v_scales, scales),
            v_viewmats if ctx.needs_input_grad[4] else None,
        ) + (None,) * 12


class _RasterizeToPixels(torch.autograd.Function):
    """Rasterize gaussians (reference: _wrapper.py:1251-1378)."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size, isect_offsets,
                flatten_ids, absgrad, shared_colors=False, shared_opacities=False):
        mod_c = colors.shape[-2] if shared_colors else 0
        mod_o = opacities.shape[-1] if shared_opacities else 0
        render_colors, render_alphas, last_ids = _C.rasterize_to_pixels_3dgs_fwd(
            means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size, isect_offsets,
            flatten_ids, mod_c, mod_o,
        )
        ctx.save_for_backward(means2d, conics, colors, opacities, backgrounds, masks, isect_offsets, flatten_ids,
                              render_alphas, last_ids)
        ctx.width = width
        ctx.height = height
        ctx.tile_size = tile_size
        ctx.absgrad = absgrad
        ctx.mods = (mod_c, mod_o)
        ctx.means2d_ref = means2d
        return render_colors, render_alphas

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas):
        (means2d, conics, colors, opacities, backgrounds, masks, isect_offsets, flatten_ids, render_alphas,
         last_ids) = ctx.saved_tensors
        v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities = _C.rasterize_to_pixels_3dgs_bwd(
            means2d, conics, colors, opacities, backgrounds, masks, ctx.width, ctx.height, ctx.tile_size,
            isect_offsets, flatten_ids, render_alphas, last_ids, v_render_colors.contiguous(),
            v_render_alphas.contiguous(), ctx.absgrad, ctx.mods[0], ctx.mods[1],
        )
        if ctx.absgrad:
            means2d.absgrad = v_means2d_abs
        if ctx.needs_input_grad[4]:
            v_backgrounds = (v_render_colors * (1.0 - render_alphas).float()).sum(dim=(-3, -2))
        else:
            v_backgrounds = None
        return (v_means2d, v_conics, v_colors, v_opacities, v_backgrounds) + (None,) * 9
