"""`rasterization()`: the API boundary of the hot path, signature-compatible with gsplat.rendering.rasterization()
(gsplat/rendering.py:33-770), so it drops into main.py:328-339 and examples/simple_trainer.py:601-624 unchanged.

Extra optional keyword arguments (defaults reproduce the reference exactly):
    cluster_ids  int32 [N]   body index per Gaussian (< 0 = static)
    body_quats   [K, 4]      wxyz rotation of each body (normalised internally, main.py:207)
    body_trans   [K, 3]      translation of each body
    body_centers [K, 3]      pivot of each body (apply_transform() uses the body's mean centre, main.py:210)
With them the per-body `apply_transform()` calls of the animation loop (main.py:366-400) collapse into the projection
kernel: no clones of the splat tensors, no extra pass over HBM.

Out of scope (raise NotImplementedError): with_ut / with_eval3d / lens distortion / rolling shutter (the 3DGUT path,
SURVEY.md section 2 rows 17) -- `with_ut=False` on every hot-path call of the reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.distributed
from torch import Tensor
from typing_extensions import Literal

from .rigid import make_rigid
from .sh import spherical_harmonics
from .wrapper import fully_fused_projection, isect_offset_encode, isect_tiles, rasterize_to_pixels


def rasterization(
    means: Tensor,  # [..., N, 3]
    quats: Tensor,  # [..., N, 4]
    scales: Tensor,  # [..., N, 3]
    opacities: Tensor,  # [..., N]
    colors: Tensor,  # [..., (C,) N, D] or [..., (C,) N, K, 3]
    viewmats: Tensor,  # [..., C, 4, 4]
    Ks: Tensor,  # [..., C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: Literal["RGB", "D", "ED", "RGB+D", "RGB+ED"] = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: Literal["classic", "antialiased"] = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: Literal["pinhole", "ortho", "fisheye", "ftheta"] = "pinhole",
    segmented: bool = False,
    covars: Optional[Tensor] = None,
    with_ut: bool = False,
    with_eval3d: bool = False,
    radial_coeffs: Optional[Tensor] = None,
    tangential_coeffs: Optional[Tensor] = None,
    thin_prism_coeffs: Optional[Tensor] = None,
    ftheta_coeffs=None,
    rolling_shutter=None,
    viewmats_rs: Optional[Tensor] = None,
    # --- rigid-body extension (this framework) ---
    cluster_ids: Optional[Tensor] = None,  # [N] int32
    body_quats: Optional[Tensor] = None,  # [K, 4]
    body_trans: Optional[Tensor] = None,  # [K, 3]
    body_centers: Optional[Tensor] = None,  # [K, 3]
) -> Tuple[Tensor, Tensor, Dict]:
    """Rasterize a set of 3D Gaussians (N) to a batch of image planes (C).

    Returns (render_colors [..., C, H, W, X], render_alphas [..., C, H, W, 1], meta) with the reference's `meta` keys
    (rendering.py:455-468, 651-665); `meta["means2d"]` is a grad-tracking non-leaf ([..., C, N, 2] or [nnz, 2]).
    """
    meta: Dict = {}

    batch_dims = means.shape[:-2]
    num_batch_dims = len(batch_dims)
    B = math.prod(batch_dims)
    N = means.shape[-2]
    C = viewmats.shape[-3]
    I = B * C
    device = means.device
    assert means.shape == batch_dims + (N, 3), means.shape
    if covars is None:
        assert quats.shape == batch_dims + (N, 4), quats.shape
        assert scales.shape == batch_dims + (N, 3), scales.shape
    else:
        assert covars.shape == batch_dims + (N, 3, 3), covars.shape
        quats, scales = None, None
        tri_indices = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
        covars = covars[..., tri_indices[0], tri_indices[1]]
    assert opacities.shape == batch_dims + (N,), opacities.shape
    assert viewmats.shape == batch_dims + (C, 4, 4), viewmats.shape
    assert Ks.shape == batch_dims + (C, 3, 3), Ks.shape
    assert render_mode in ["RGB", "D", "ED", "RGB+D", "RGB+ED"], render_mode
    assert tile_size == 16, "tile_size must be 16 (the only value the reference exercises, rendering.py:184-185)"

    if with_ut or with_eval3d:
        raise NotImplementedError("with_ut / with_eval3d (3DGUT) are outside this framework's hot path")
    if (radial_coeffs is not None or tangential_coeffs is not None or thin_prism_coeffs is not None
            or ftheta_coeffs is not None or viewmats_rs is not None
            or (rolling_shutter is not None and getattr(rolling_shutter, "name", str(rolling_shutter)) != "GLOBAL")):
        raise AssertionError("Distortion and rolling shutter are only supported with `with_ut=True`.")

    if sh_degree is None:
        assert (colors.dim() == num_batch_dims + 2 and colors.shape[:-1] == batch_dims + (N,)) or (
            colors.dim() == num_batch_dims + 3 and colors.shape[:-1] == batch_dims + (C, N)
        ), colors.shape
        if distributed:
            assert colors.dim() == num_batch_dims + 2, "Distributed mode only supports per-Gaussian colors."
    else:
        assert (
            colors.dim() == num_batch_dims + 3 and colors.shape[:-2] == batch_dims + (N,) and colors.shape[-1] == 3
        ) or (
            colors.dim() == num_batch_dims + 4 and colors.shape[:-2] == batch_dims + (C, N) and colors.shape[-1] == 3
        ), colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[-2], colors.shape
        if distributed:
            assert colors.dim() == num_batch_dims + 3, "Distributed mode only supports per-Gaussian colors."
    if absgrad:
        assert not distributed, "AbsGrad is not supported in distributed mode."

    rigid = make_rigid(cluster_ids, body_quats, body_trans, body_centers)
    if rigid is not None:
        assert cluster_ids.shape == (N,), cluster_ids.shape

    if distributed:
        from .distributed import all_gather_int32, all_gather_tensor_list

        assert batch_dims == (), "Distributed mode does not support batch dimensions"
        world_rank = torch.distributed.get_rank()
        world_size = torch.distributed.get_world_size()
        N_world = all_gather_int32(world_size, N, device=device)
        C_world = [C] * world_size
        viewmats, Ks = all_gather_tensor_list(world_size, [viewmats, Ks])
        C = len(viewmats)

    proj_results = fully_fused_projection(
        means, covars, quats, scales, viewmats, Ks, width, height,
        eps2d=eps2d, packed=packed, near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip,
        sparse_grad=sparse_grad, calc_compensations=(rasterize_mode == "antialiased"), camera_model=camera_model,
        opacities=opacities, rigid=rigid,
    )

    if packed:
        batch_ids, camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations = proj_results
        opacities = opacities.view(B, N)[batch_ids, gaussian_ids]  # [nnz]
        image_ids = batch_ids * C + camera_ids
    else:
        radii, means2d, depths, conics, compensations = proj_results
        opacities = torch.broadcast_to(opacities[..., None, :], batch_dims + (C, N))  # [..., C, N]
        batch_ids, camera_ids, gaussian_ids = None, None, None
        image_ids = None

    if compensations is not None:
        opacities = opacities * compensations

    meta.update(
        {
            "batch_ids": batch_ids,
            "camera_ids": camera_ids,
            "gaussian_ids": gaussian_ids,
            "radii": radii,
            "means2d": means2d,
            "depths": depths,
            "conics": conics,
            "opacities": opacities,
        }
    )

    if sh_degree is None:
        if packed:
            if colors.dim() == num_batch_dims + 2:
                colors = colors.view(B, N, -1)[batch_ids, gaussian_ids]
            else:
                colors = colors.view(B, C, N, -1)[batch_ids, camera_ids, gaussian_ids]
        else:
            if colors.dim() == num_batch_dims + 2:
                colors = torch.broadcast_to(colors[..., None, :, :], batch_dims + (C, N, colors.shape[-1]))
    else:
        # apply_transform() moves the means but leaves sh0/shN untouched (main.py:200-226), so the view directions of
        # rendering.py:491-525 follow the MOVED means while the SH coefficients are not rotated with the body.
        world_means = means if rigid is None else _transformed_means(means, rigid)
        campos = torch.inverse(viewmats)[..., :3, 3]  # [..., C, 3]
        if packed:
            dirs = world_means.view(B, N, 3)[batch_ids, gaussian_ids] - campos.view(B, C, 3)[batch_ids, camera_ids]
            masks = (radii > 0).all(dim=-1)
            if colors.dim() == num_batch_dims + 3:
                shs = colors.view(B, N, -1, 3)[batch_ids, gaussian_ids]
            else:
                shs = colors.view(B, C, N, -1, 3)[batch_ids, camera_ids, gaussian_ids]
            colors = spherical_harmonics(sh_degree, dirs, shs, masks=masks)
        else:
            dirs = world_means[..., None, :, :] - campos[..., None, :]  # [..., C, N, 3]
            masks = (radii > 0).all(dim=-1)
            if colors.dim() == num_batch_dims + 3:
                shs = torch.broadcast_to(colors[..., None, :, :, :], batch_dims + (C, N) + colors.shape[-2:])
            else:
                shs = colors
            colors = spherical_harmonics(sh_degree, dirs, shs, masks=masks)
        colors = torch.clamp_min(colors + 0.5, 0.0)

    if distributed:
        from .distributed import exchange_projected

        (radii, means2d, depths, conics, opacities, colors, camera_ids, gaussian_ids, C) = exchange_projected(
            world_rank, world_size, packed, N, N_world, C_world, radii, means2d, depths, conics, opacities, colors,
            camera_ids, gaussian_ids,
        )
        if packed:
            image_ids = camera_ids  # B == 1 in distributed mode
        I = C

    if render_mode in ["RGB+D", "RGB+ED"]:
        colors = torch.cat((colors, depths[..., None]), dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat(
                [backgrounds, torch.zeros(batch_dims + (C, 1), device=backgrounds.device)], dim=-1
            )
    elif render_mode in ["D", "ED"]:
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(batch_dims + (C, 1), device=backgrounds.device)

    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(
        means2d, radii, depths, tile_size, tile_width, tile_height,
        segmented=segmented, packed=packed, n_images=I, image_ids=image_ids, gaussian_ids=gaussian_ids,
    )
    isect_offsets = isect_offset_encode(isect_ids, I, tile_width, tile_height)
    isect_offsets = isect_offsets.reshape(batch_dims + (C, tile_height, tile_width))

    meta.update(
        {
            "tile_width": tile_width,
            "tile_height": tile_height,
            "tiles_per_gauss": tiles_per_gauss,
            "isect_ids": isect_ids,
            "flatten_ids": flatten_ids,
            "isect_offsets": isect_offsets,
            "width": width,
            "height": height,
            "tile_size": tile_size,
            "n_batches": B,
            "n_cameras": C,
        }
    )

    if colors.shape[-1] > channel_chunk:
        n_chunks = (colors.shape[-1] + channel_chunk - 1) // channel_chunk
        render_colors, render_alphas = [], []
        for i in range(n_chunks):
            colors_chunk = colors[..., i * channel_chunk : (i + 1) * channel_chunk]
            backgrounds_chunk = (
                backgrounds[..., i * channel_chunk : (i + 1) * channel_chunk] if backgrounds is not None else None
            )
            render_colors_, render_alphas_ = rasterize_to_pixels(
                means2d, conics, colors_chunk, opacities, width, height, tile_size, isect_offsets, flatten_ids,
                backgrounds=backgrounds_chunk, packed=packed, absgrad=absgrad,
            )
            render_colors.append(render_colors_)
            render_alphas.append(render_alphas_)
        render_colors = torch.cat(render_colors, dim=-1)
        render_alphas = render_alphas[0]  # discard the rest
    else:
        render_colors, render_alphas = rasterize_to_pixels(
            means2d, conics, colors, opacities, width, height, tile_size, isect_offsets, flatten_ids,
            backgrounds=backgrounds, packed=packed, absgrad=absgrad,
        )
    if render_mode in ["ED", "RGB+ED"]:
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1
        )
    return render_colors, render_alphas, meta


def _transformed_means(means: Tensor, rigid) -> Tensor:
    """Means after the rigid transform, in torch (only used for SH view directions)."""
    from .torch_ref import apply_rigid_torch

    return apply_rigid_torch(means, None, rigid)[0]
