"""`rasterization()`: the API boundary of the hot path.

Call-compatible with gsplat.rendering.rasterization() (gsplat/rendering.py:33-770: same positional / keyword arguments,
defaults, return triple and `meta` keys), so the reference's callers -- main.py:328-339, examples/simple_trainer.py:601-624,
examples/simple_viewer.py:60-73 -- run unchanged.  Four optional keyword arguments are new; left at None the behaviour is
the reference's:

    cluster_ids  int32 [N]   body index of every Gaussian (< 0: static background)
    body_quats   [K, 4]      per-body rotation, wxyz, normalised inside the kernel (main.py:207)
    body_trans   [K, 3]      per-body translation
    body_centers [K, 3]      per-body pivot; None = the mean of the body's Gaussian centres, as apply_transform() uses
                             (main.py:210) -- computed per call, so an animation loop should pass it precomputed

They replace the animation loop's per-body `apply_transform()` calls (main.py:366-400): the pose table is consumed inside
the projection kernel, so no splat tensor is cloned and no extra pass over HBM is made.

The body is organised as five stages -- project, shade, (exchange), bin, composite -- each a thin call into the operator
layer (`wrapper.py` -> `_C.py` -> C ABI).  Not supported (NotImplementedError): the 3DGUT options (`with_ut`,
`with_eval3d`, distortion coefficients, rolling shutter); no hot-path caller of the reference enables them.
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

_DEPTH_MODES = ("D", "ED", "RGB+D", "RGB+ED")
_MODES = ("RGB",) + _DEPTH_MODES


class _Dims:
    """Leading batch dims [...], B = prod(...), N Gaussians, C cameras of one call."""

    def __init__(self, means: Tensor, viewmats: Tensor):
        self.batch = tuple(means.shape[:-2])
        self.nb = len(self.batch)
        self.B = math.prod(self.batch)
        self.N = means.shape[-2]
        self.C = viewmats.shape[-3]

    def expect(self, t: Tensor, tail: Tuple[int, ...], what: str) -> None:
        assert tuple(t.shape) == self.batch + tail, f"{what}: {tuple(t.shape)}"


def _reject_3dgut(with_ut, with_eval3d, radial, tangential, prism, ftheta, shutter, viewmats_rs) -> None:
    if with_ut or with_eval3d:
        raise NotImplementedError("with_ut / with_eval3d (3DGUT) are outside this framework's hot path")
    global_shutter = shutter is None or getattr(shutter, "name", str(shutter)) == "GLOBAL"
    if any(x is not None for x in (radial, tangential, prism, ftheta, viewmats_rs)) or not global_shutter:
        raise AssertionError("Distortion and rolling shutter are only supported with `with_ut=True`.")


def _check_colors(d: _Dims, colors: Tensor, sh_degree: Optional[int], distributed: bool) -> bool:
    """Validates the colour tensor; returns True when it carries a camera dimension."""
    extra = 2 if sh_degree is None else 3  # trailing dims after the (C,) N part: [D] or [K, 3]
    per_camera = colors.dim() == d.nb + extra + 1
    lead = d.batch + ((d.C, d.N) if per_camera else (d.N,))
    assert colors.dim() in (d.nb + extra, d.nb + extra + 1) and tuple(colors.shape[: len(lead)]) == lead, colors.shape
    if sh_degree is not None:
        assert colors.shape[-1] == 3 and (sh_degree + 1) ** 2 <= colors.shape[-2], colors.shape
    if distributed:
        assert not per_camera, "Distributed mode only supports per-Gaussian colors."
    return per_camera


def _view_dependent_colors(d, colors, per_camera, sh_degree, means, rigid, viewmats, radii, packed, rows):
    """SH colours for the view directions of the (moved) Gaussians: rendering.py:491-525.  apply_transform() moves the
    means but leaves sh0 / shN untouched (main.py:200-226), so directions follow the moved means and the coefficients are
    NOT rotated with the body -- reproduced here on purpose."""
    if rigid is not None:
        from .torch_ref import apply_rigid_torch

        means = apply_rigid_torch(means, None, rigid)[0]
    cam_origin = torch.linalg.inv(viewmats)[..., :3, 3]  # [..., C, 3]
    visible = (radii > 0).all(dim=-1)
    if packed:
        b, c, g = rows
        dirs = means.reshape(d.B, d.N, 3)[b, g] - cam_origin.reshape(d.B, d.C, 3)[b, c]
        coeffs = (colors.reshape(d.B, d.C, d.N, -1, 3)[b, c, g] if per_camera
                  else colors.reshape(d.B, d.N, -1, 3)[b, g])
    else:
        dirs = means.unsqueeze(-3) - cam_origin.unsqueeze(-2)  # [..., C, N, 3]
        coeffs = colors if per_camera else colors.unsqueeze(-4).expand(d.batch + (d.C,) + tuple(colors.shape[-3:]))
    rgb = spherical_harmonics(sh_degree, dirs, coeffs, masks=visible)
    return (rgb + 0.5).clamp_min(0.0)


def _plain_colors(d, colors, per_camera, packed, rows):
    if packed:
        b, c, g = rows
        return colors.reshape(d.B, d.C, d.N, -1)[b, c, g] if per_camera else colors.reshape(d.B, d.N, -1)[b, g]
    if per_camera:
        return colors
    # a stride-0 view: the compositing kernels read one row per Gaussian for every camera (no [C,N,D] copy)
    return colors.unsqueeze(-3).expand(d.batch + (d.C,) + tuple(colors.shape[-2:]))


def _attach_depth(render_mode, colors, depths, backgrounds, lead_shape):
    """RGB+D / RGB+ED append the depth as one more channel, D / ED render it alone (rendering.py:614-629); the extra
    channel's background is zero."""
    if render_mode not in _DEPTH_MODES:
        return colors, backgrounds
    depth_col = depths.unsqueeze(-1)
    zero_bg = None if backgrounds is None else torch.zeros(lead_shape + (1,), device=backgrounds.device,
                                                           dtype=backgrounds.dtype)
    if render_mode.startswith("RGB"):
        colors = torch.cat([colors, depth_col], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, zero_bg], dim=-1)
    else:
        colors, backgrounds = depth_col, zero_bg
    return colors, backgrounds


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
    """Rasterize N 3D Gaussians to C image planes.

    Returns (render_colors [..., C, H, W, X], render_alphas [..., C, H, W, 1], meta); `meta` carries the reference's keys
    (rendering.py:455-468, 651-665) and `meta["means2d"]` is a grad-tracking non-leaf ([..., C, N, 2], or [nnz, 2] when
    packed) so that `retain_grad()` / `.absgrad` work as in the reference's training loop.

    Lifetime note (distributed=True, packed=True, no gradients): the per-splat `meta` tensors (means2d, radii, depths,
    conics, opacities, camera_ids, gaussian_ids) are then views of this rank's persistent peer-memory receive arrays
    (distributed.PeerSplatExchange) and are overwritten by the NEXT distributed rasterization() call on this process
    group; clone what must outlive it.
    """
    # ---- 0. validation -------------------------------------------------------------------------------------------------
    d = _Dims(means, viewmats)
    d.expect(means, (d.N, 3), "means")
    d.expect(opacities, (d.N,), "opacities")
    d.expect(viewmats, (d.C, 4, 4), "viewmats")
    d.expect(Ks, (d.C, 3, 3), "Ks")
    if covars is not None:
        d.expect(covars, (d.N, 3, 3), "covars")
        rows6, cols6 = (0, 0, 0, 1, 1, 2), (0, 1, 2, 1, 2, 2)
        covars = covars[..., rows6, cols6]  # upper triangle, as the projection operator expects
        quats = scales = None
    else:
        d.expect(quats, (d.N, 4), "quats")
        d.expect(scales, (d.N, 3), "scales")
    assert render_mode in _MODES, render_mode
    assert tile_size == 16, "tile_size must be 16 (the only value the reference exercises, rendering.py:184-185)"
    _reject_3dgut(with_ut, with_eval3d, radial_coeffs, tangential_coeffs, thin_prism_coeffs, ftheta_coeffs,
                  rolling_shutter, viewmats_rs)
    per_camera_colors = _check_colors(d, colors, sh_degree, distributed)
    assert not (absgrad and distributed), "AbsGrad is not supported in distributed mode."
    rigid = make_rigid(cluster_ids, body_quats, body_trans, body_centers, means)
    if rigid is not None:
        assert tuple(cluster_ids.shape) == (d.N,), cluster_ids.shape
    device = means.device

    shard = None
    if distributed:
        # Gaussian-sharded scene (rendering.py:366-381): every rank projects its Gaussians to ALL cameras, then the
        # projected splats travel to the rank that owns the camera.
        from .distributed import GaussianShardExchange

        assert d.batch == (), "Distributed mode does not support batch dimensions"
        shard = GaussianShardExchange(d.N, d.C, device)
        viewmats, Ks = shard.gather_cameras(viewmats, Ks)
        d.C = viewmats.shape[0]

    # Gaussian-sharded render without gradients: packed rows go to their cameras' ranks through NVLink peer memory, one
    # kernel per rank and no NCCL call on the data path (distributed.PeerSplatExchange); training keeps the
    # differentiable all-to-all below.
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (means, quats, scales, covars, opacities, colors, viewmats))
    peer_path = peer_grad = False
    if shard is not None and packed and means.is_cuda:
        from .distributed import PeerSplatExchange

        # same host, peer access between all devices, <= 16 ranks -- else the NCCL route below (probed once per group)
        usable = PeerSplatExchange.enabled and PeerSplatExchange.usable(shard.group, device)
        peer_path = usable and not needs_grad
        # training: the same exchange as an autograd node whose backward is the transposed peer-memory exchange
        peer_grad = usable and needs_grad and PeerSplatExchange.differentiable

    # ---- 1. project (rigid transform fused in; without gradients also the SH colours) -------------------------------------
    from . import _C
    from .wrapper import _CAMERA_MODELS

    # No gradient to carry: the view-dependent colours are evaluated INSIDE the projection kernel for the moved means (no
    # eager-torch rigid transform, no torch.linalg.inv, no [C,N,3] dirs tensor, no separate SH pass) -- what main.py:328-339
    # renders with (sh_degree=3).  Training keeps the differentiable spherical_harmonics() operator below.
    fused_sh = (sh_degree is not None and not needs_grad and not per_camera_colors and means.is_cuda
                and colors.dtype == torch.float32)
    sh_arg = (colors.contiguous(), int(sh_degree)) if fused_sh else None
    sh_rgb = None
    if peer_path or (fused_sh and packed):
        out = _C.projection_ewa_3dgs_packed_fwd(
            means.contiguous(), None if covars is None else covars.contiguous(),
            None if quats is None else quats.contiguous(), None if scales is None else scales.contiguous(),
            opacities.contiguous(), viewmats.contiguous(), Ks.contiguous(), width, height, eps2d, near_plane,
            far_plane, radius_clip, rasterize_mode == "antialiased", _CAMERA_MODELS[camera_model], rigid, None, sh_arg)
        indptr, batch_ids, camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations = out[:9]
        sh_rgb = out[9] if fused_sh else None
        projected = None
    elif fused_sh:
        radii, means2d, depths, conics, compensations, sh_rgb = _C.projection_ewa_3dgs_fused_fwd(
            means.contiguous(), None if covars is None else covars.contiguous(),
            None if quats is None else quats.contiguous(), None if scales is None else scales.contiguous(),
            opacities.contiguous(), viewmats.contiguous(), Ks.contiguous(), width, height, eps2d, near_plane, far_plane,
            radius_clip, rasterize_mode == "antialiased", _CAMERA_MODELS[camera_model], rigid, None, sh_arg)
        projected = None
    else:
        projected = fully_fused_projection(
            means, covars, quats, scales, viewmats, Ks, width, height, eps2d=eps2d, packed=packed, near_plane=near_plane,
            far_plane=far_plane, radius_clip=radius_clip, sparse_grad=sparse_grad,
            calc_compensations=(rasterize_mode == "antialiased"), camera_model=camera_model, opacities=opacities,
            rigid=rigid)
    if peer_path:
        rows = (batch_ids, camera_ids, gaussian_ids)
        alpha_in = None  # opacity x compensation is formed inside the exchange kernel
        image_ids = camera_ids
    elif packed:
        if projected is not None:
            batch_ids, camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations = projected
        rows = (batch_ids, camera_ids, gaussian_ids)
        alpha_in = opacities.reshape(d.B, d.N)[batch_ids, gaussian_ids]  # [nnz]
        image_ids = batch_ids * d.C + camera_ids
    else:
        if projected is not None:
            radii, means2d, depths, conics, compensations = projected
        batch_ids = camera_ids = gaussian_ids = image_ids = rows = None
        alpha_in = opacities.unsqueeze(-2).expand(d.batch + (d.C, d.N))  # stride-0 view, consumed without a copy
    if compensations is not None and not peer_path:
        alpha_in = alpha_in * compensations
    # ---- 2. shade ----------------------------------------------------------------------------------------------------------
    if sh_rgb is not None:
        shaded = sh_rgb
    elif peer_path and sh_degree is None and not per_camera_colors:
        shaded = None  # the exchange kernel gathers the colour row of each splat itself
    elif sh_degree is None:
        shaded = _plain_colors(d, colors, per_camera_colors, packed, rows)
    else:
        shaded = _view_dependent_colors(d, colors, per_camera_colors, sh_degree, means, rigid, viewmats, radii, packed, rows)

    # ---- 2b. exchange (Gaussian-sharded scenes only) ---------------------------------------------------------------------
    n_images = d.B * d.C
    if peer_path:
        per_row = shaded is not None
        table = shaded if per_row else colors
        peer = PeerSplatExchange.get(shard.group, device, int(table.shape[-1]))
        radii, means2d, depths, conics, alpha_in, shaded, camera_ids, gaussian_ids = peer.exchange(
            shard.local_cameras, indptr, camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations,
            opacities, False, table, per_row, shard.gaussian_base)
        d.C = shard.local_cameras
        n_images = d.C
        image_ids = camera_ids
        batch_ids = torch.zeros_like(camera_ids)
    elif peer_grad:
        n_ch = int(shaded.shape[-1])
        peer = PeerSplatExchange.get(shard.group, device, n_ch)
        edges = torch.arange(d.C + 1, device=device, dtype=camera_ids.dtype)
        indptr = torch.searchsorted(camera_ids.contiguous(), edges).to(torch.int32)  # rows are ordered by camera
        radii, means2d, depths, conics, alpha_in, shaded, camera_ids, gaussian_ids = peer.exchange_differentiable(
            shard.local_cameras, indptr, camera_ids, gaussian_ids, radii, means2d, depths, conics, alpha_in,
            shaded.reshape(-1, n_ch), shard.gaussian_base)
        d.C = shard.local_cameras
        n_images = d.C
        image_ids = camera_ids
        batch_ids = torch.zeros_like(camera_ids)
    elif shard is not None:
        out = shard.exchange(packed, radii, means2d, depths, conics, alpha_in, shaded, camera_ids, gaussian_ids)
        radii, means2d, depths, conics, alpha_in, shaded, camera_ids, gaussian_ids = out
        d.C = shard.local_cameras
        n_images = d.C
        if packed:
            image_ids = camera_ids
            batch_ids = torch.zeros_like(camera_ids)
    # the splats this rank composites (after the exchange, as in rendering.py:651-665)
    meta: Dict = dict(batch_ids=batch_ids, camera_ids=camera_ids, gaussian_ids=gaussian_ids, radii=radii, means2d=means2d,
                      depths=depths, conics=conics, opacities=alpha_in)

    shaded, backgrounds = _attach_depth(render_mode, shaded, depths, backgrounds, d.batch + (d.C,))

    # ---- 3. bin: tile intersection, ordering, per-tile offsets ------------------------------------------------------------
    tile_w, tile_h = -(-width // tile_size), -(-height // tile_size)
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(
        means2d, radii, depths, tile_size, tile_w, tile_h, segmented=segmented, packed=packed, n_images=n_images,
        image_ids=image_ids, gaussian_ids=gaussian_ids)
    isect_offsets = isect_offset_encode(isect_ids, n_images, tile_w, tile_h).reshape(d.batch + (d.C, tile_h, tile_w))
    meta.update(tile_width=tile_w, tile_height=tile_h, tiles_per_gauss=tiles_per_gauss, isect_ids=isect_ids,
                flatten_ids=flatten_ids, isect_offsets=isect_offsets, width=width, height=height, tile_size=tile_size,
                n_batches=d.B, n_cameras=d.C)

    # ---- 4. composite (channels beyond `channel_chunk` in slices; only the first slice's alpha is kept, 668-720) ----------
    n_ch = shaded.shape[-1]
    pieces, render_alphas = [], None
    step = max(int(channel_chunk), 1)
    for lo in range(0, n_ch, step):
        hi = min(lo + step, n_ch)
        whole = lo == 0 and hi == n_ch
        part = shaded if whole else shaded[..., lo:hi]
        part_bg = backgrounds if (backgrounds is None or whole) else backgrounds[..., lo:hi]
        img, alpha = rasterize_to_pixels(means2d, conics, part, alpha_in, width, height, tile_size, isect_offsets,
                                         flatten_ids, backgrounds=part_bg, packed=packed, absgrad=absgrad)
        pieces.append(img)
        if render_alphas is None:
            render_alphas = alpha
    render_colors = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=-1)

    if render_mode in ("ED", "RGB+ED"):  # expected depth = accumulated depth / accumulated alpha (760-768)
        expected = render_colors[..., -1:] / render_alphas.clamp(min=1e-10)
        render_colors = torch.cat([render_colors[..., :-1], expected], dim=-1)
    return render_colors, render_alphas, meta
