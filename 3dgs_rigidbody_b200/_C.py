"""Operator boundary: a stand-in for the reference's pybind module `_C` (gsplat/cuda/ext.cpp:6-104).

Same function names and the same POSITIONAL signatures as the hot-path subset of `_C` (C++ declarations in
gsplat/cuda/include/Ops.h:42-168, 186-204, 223-263: fused and packed projection fwd/bwd, spherical harmonics, tile
intersection + offsets, compositing fwd/bwd), so `gsplat.cuda._backend._C` can be replaced by this module
(see INTEGRATION.md).  Each function validates its inputs like the reference host launchers (CHECK_INPUT ->
RuntimeError), allocates the outputs with torch exactly where the reference does, and calls the hand-written sm_100a
kernels through the C ABI (include/rigidsplat.h) on the current CUDA stream.  There is no CPU / eager fallback.

Extension over the reference: the projection operators (fused and packed, fwd and bwd) take an optional trailing `rigid`
argument (`RigidPoses`) that fuses main.py's apply_transform() into the projection.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import RigidSplatError  # noqa: F401  (re-export)

# gsplat/cuda/include/Common.h:46-51, exported with export_values() (ext.cpp:8-13)
PINHOLE, ORTHO, FISHEYE, FTHETA = 0, 1, 2, 3


class CameraModelType:
    PINHOLE, ORTHO, FISHEYE, FTHETA = 0, 1, 2, 3


@dataclass
class RigidPoses:
    """Per-Gaussian cluster ids + per-body poses (replaces one apply_transform() call per body, main.py:183-228)."""

    cluster_ids: Tensor  # int32 [N]; < 0 = static
    body_quats: Tensor  # float32 [K,4] wxyz
    body_trans: Tensor  # float32 [K,3]
    body_centers: Optional[Tensor] = None  # float32 [K,3]; None = rotate about the origin

    def validate(self, N: int, device) -> None:
        K = self.body_quats.shape[0]
        _check(self.cluster_ids, "cluster_ids", torch.int32, (N,), device)
        _check(self.body_quats, "body_quats", torch.float32, (K, 4), device)
        _check(self.body_trans, "body_trans", torch.float32, (K, 3), device)
        if self.body_centers is not None:
            _check(self.body_centers, "body_centers", torch.float32, (K, 3), device)
        if K < 1:
            raise RuntimeError("RigidPoses: at least one body is required")

    def fill(self, r) -> None:
        r.cluster_ids = self.cluster_ids.data_ptr()
        r.body_quats = self.body_quats.data_ptr()
        r.body_trans = self.body_trans.data_ptr()
        r.body_centers = self.body_centers.data_ptr() if self.body_centers is not None else None
        r.K = self.body_quats.shape[0]


def _check(t: Tensor, name: str, dtype=None, shape=None, device=None) -> None:
    if not isinstance(t, Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    if device is not None and t.device != device:
        raise RuntimeError(f"{name} must be on {device}, got {t.device}")


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cam(camera_model) -> int:
    return int(camera_model)


# ---------------------------------------------------------------------------------------------------------------------
# projection (Ops.h:42-88; csrc/Projection.cpp:104-281)
# ---------------------------------------------------------------------------------------------------------------------
def projection_ewa_3dgs_fused_fwd(
    means: Tensor,  # [..., N, 3]
    covars: Optional[Tensor],  # [..., N, 6]
    quats: Optional[Tensor],  # [..., N, 4]
    scales: Optional[Tensor],  # [..., N, 3]
    opacities: Optional[Tensor],  # [..., N]
    viewmats: Tensor,  # [..., C, 4, 4]
    Ks: Tensor,  # [..., C, 3, 3]
    image_width: int,
    image_height: int,
    eps2d: float,
    near_plane: float,
    far_plane: float,
    radius_clip: float,
    calc_compensations: bool,
    camera_model: int,
    rigid: Optional[RigidPoses] = None,
    _tile_count: Optional[Tuple[int, int, int]] = None,
    _sh: Optional[Tuple[Tensor, int]] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Optional[Tensor]]:
    """With `_sh = (coeffs [..., N, K, 3], degree)` (an extension: not part of the reference's operator) the view-dependent
    colours max(SH(degree, moved mean - camera origin) + 0.5, 0) of the visible rows are evaluated by the same kernel and
    returned as a sixth output [..., C, N, 3] (rendering.py:491-525 without the dirs / inverse / colours passes)."""
    lib = _lib.load()
    _check(means, "means", torch.float32)
    _check(viewmats, "viewmats", torch.float32)
    _check(Ks, "Ks", torch.float32)
    if covars is not None:
        _check(covars, "covars", torch.float32)
    else:
        if quats is None or scales is None:
            raise RuntimeError("either covars or (quats, scales) must be provided")
        _check(quats, "quats", torch.float32)
        _check(scales, "scales", torch.float32)
    if opacities is not None:
        _check(opacities, "opacities", torch.float32)
    N = means.shape[-2]
    C = viewmats.shape[-3]
    B = means.numel() // (N * 3) if N > 0 else 1
    batch_dims = tuple(means.shape[:-2])
    dev = means.device
    if rigid is not None:
        rigid.validate(N, dev)
    with torch.cuda.device(dev):
        radii = torch.empty(batch_dims + (C, N, 2), dtype=torch.int32, device=dev)
        means2d = torch.empty(batch_dims + (C, N, 2), dtype=torch.float32, device=dev)
        depths = torch.empty(batch_dims + (C, N), dtype=torch.float32, device=dev)
        conics = torch.empty(batch_dims + (C, N, 3), dtype=torch.float32, device=dev)
        compensations = (
            torch.empty(batch_dims + (C, N), dtype=torch.float32, device=dev) if calc_compensations else None
        )
        a = _lib.rs_project_fwd_args()
        a.B, a.C, a.N = B, C, N
        a.image_width, a.image_height = int(image_width), int(image_height)
        a.camera_model = _cam(camera_model)
        a.eps2d, a.near_plane, a.far_plane, a.radius_clip = eps2d, near_plane, far_plane, radius_clip
        a.means, a.covars, a.quats, a.scales = _ptr(means), _ptr(covars), _ptr(None if covars is not None else quats), _ptr(None if covars is not None else scales)
        a.opacities, a.viewmats, a.Ks = _ptr(opacities), _ptr(viewmats), _ptr(Ks)
        if rigid is not None:
            rigid.fill(a.rigid)
        a.radii, a.means2d, a.depths, a.conics = _ptr(radii), _ptr(means2d), _ptr(depths), _ptr(conics)
        a.compensations = _ptr(compensations)
        sh_colors = None
        if _sh is not None:
            coeffs, degree = _sh
            _check(coeffs, "sh coefficients", torch.float32, tuple(batch_dims) + (N, coeffs.shape[-2], 3), dev)
            sh_colors = torch.zeros(batch_dims + (C, N, 3), dtype=torch.float32, device=dev)
            a.sh_coeffs, a.sh_colors = _ptr(coeffs), _ptr(sh_colors)
            a.sh_degree, a.sh_K = int(degree), coeffs.shape[-2]
        _lib.check(lib.rs_project_fwd(ctypes.byref(a), _stream()))
    if _sh is not None:
        return radii, means2d, depths, conics, compensations, sh_colors
    return radii, means2d, depths, conics, compensations


def projection_ewa_3dgs_fused_bwd(
    means: Tensor,
    covars: Optional[Tensor],
    quats: Optional[Tensor],
    scales: Optional[Tensor],
    viewmats: Tensor,
    Ks: Tensor,
    image_width: int,
    image_height: int,
    eps2d: float,
    camera_model: int,
    radii: Tensor,
    conics: Tensor,
    compensations: Optional[Tensor],
    v_means2d: Tensor,
    v_depths: Tensor,
    v_conics: Tensor,
    v_compensations: Optional[Tensor],
    viewmats_requires_grad: bool,
    rigid: Optional[RigidPoses] = None,
) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
    lib = _lib.load()
    for t, n in ((means, "means"), (viewmats, "viewmats"), (Ks, "Ks"), (conics, "conics"), (v_means2d, "v_means2d"),
                 (v_depths, "v_depths"), (v_conics, "v_conics")):
        _check(t, n, torch.float32)
    _check(radii, "radii", torch.int32)
    N = means.shape[-2]
    C = viewmats.shape[-3]
    B = means.numel() // (N * 3) if N > 0 else 1
    dev = means.device
    if rigid is not None:
        rigid.validate(N, dev)
    with torch.cuda.device(dev):
        v_means = torch.zeros_like(means)
        v_covars = v_quats = v_scales = None
        if covars is not None:
            _check(covars, "covars", torch.float32)
            v_covars = torch.zeros_like(covars)
        else:
            _check(quats, "quats", torch.float32)
            _check(scales, "scales", torch.float32)
            v_quats = torch.zeros_like(quats)
            v_scales = torch.zeros_like(scales)
        v_viewmats = torch.zeros_like(viewmats) if viewmats_requires_grad else None
        a = _lib.rs_project_bwd_args()
        a.B, a.C, a.N = B, C, N
        a.image_width, a.image_height = int(image_width), int(image_height)
        a.camera_model = _cam(camera_model)
        a.eps2d = eps2d
        a.means, a.covars = _ptr(means), _ptr(covars)
        a.quats, a.scales = _ptr(None if covars is not None else quats), _ptr(None if covars is not None else scales)
        a.viewmats, a.Ks = _ptr(viewmats), _ptr(Ks)
        if rigid is not None:
            rigid.fill(a.rigid)
        a.radii, a.conics, a.compensations = _ptr(radii), _ptr(conics), _ptr(compensations)
        a.v_means2d, a.v_depths, a.v_conics = _ptr(v_means2d), _ptr(v_depths), _ptr(v_conics)
        a.v_compensations = _ptr(v_compensations)
        a.v_means, a.v_covars, a.v_quats, a.v_scales = _ptr(v_means), _ptr(v_covars), _ptr(v_quats), _ptr(v_scales)
        a.v_viewmats = _ptr(v_viewmats)
        _lib.check(lib.rs_project_bwd(ctypes.byref(a), _stream()))
    return v_means, v_covars, v_quats, v_scales, v_viewmats


# ---------------------------------------------------------------------------------------------------------------------
# packed (COO) projection (Ops.h:98-151; csrc/Projection.cpp:283-547)
# ---------------------------------------------------------------------------------------------------------------------
def projection_ewa_3dgs_packed_fwd(
    means: Tensor,  # [..., N, 3]
    covars: Optional[Tensor],  # [..., N, 6]
    quats: Optional[Tensor],  # [..., N, 4]
    scales: Optional[Tensor],  # [..., N, 3]
    opacities: Optional[Tensor],  # [..., N]
    viewmats: Tensor,  # [..., C, 4, 4]
    Ks: Tensor,  # [..., C, 3, 3]
    image_width: int,
    image_height: int,
    eps2d: float,
    near_plane: float,
    far_plane: float,
    radius_clip: float,
    calc_compensations: bool,
    camera_model: int,
    rigid: Optional[RigidPoses] = None,
    _capacity: Optional[int] = None,
    _sh: Optional[Tuple[Tensor, int]] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Optional[Tensor]]:
    """-> (indptr i32 [B*C+1], batch_ids, camera_ids, gaussian_ids i64 [nnz], radii i32 [nnz,2], means2d [nnz,2],
    depths [nnz], conics [nnz,3], compensations [nnz] or None), rows in (batch, camera, gaussian) order.

    One projection pass (the reference runs two, csrc/Projection.cpp:331-413): the rows are written into buffers of
    `_capacity` rows (default: every pair, which cannot overflow) and the returned tensors are views of the first nnz
    rows; reading nnz is the one host sync of the call, as in the reference (Projection.cpp:369)."""
    lib = _lib.load()
    _check(means, "means", torch.float32)
    _check(viewmats, "viewmats", torch.float32)
    _check(Ks, "Ks", torch.float32)
    if covars is not None:
        _check(covars, "covars", torch.float32)
    else:
        if quats is None or scales is None:
            raise RuntimeError("either covars or (quats, scales) must be provided")
        _check(quats, "quats", torch.float32)
        _check(scales, "scales", torch.float32)
    if opacities is not None:
        _check(opacities, "opacities", torch.float32)
    N = means.shape[-2]
    C = viewmats.shape[-3]
    B = means.numel() // (N * 3) if N > 0 else 1
    dev = means.device
    if rigid is not None:
        rigid.validate(N, dev)
    total = B * C * N
    with torch.cuda.device(dev):
        indptr = torch.empty(B * C + 1, dtype=torch.int32, device=dev)
        nnz_dev = torch.empty(1, dtype=torch.int64, device=dev)
        cap = total if _capacity is None else min(int(_capacity), total)
        while True:
            ids = torch.empty((3, cap), dtype=torch.int64, device=dev)
            radii = torch.empty((cap, 2), dtype=torch.int32, device=dev)
            means2d = torch.empty((cap, 2), dtype=torch.float32, device=dev)
            depths = torch.empty((cap,), dtype=torch.float32, device=dev)
            conics = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            compensations = torch.zeros((cap,), dtype=torch.float32, device=dev) if calc_compensations else None
            sh_colors = torch.empty((cap, 3), dtype=torch.float32, device=dev) if _sh is not None else None
            ws = torch.empty(max(int(lib.rs_project_packed_workspace_bytes(B, C, N)), 16), dtype=torch.uint8, device=dev)
            pa = _lib.rs_project_packed_fwd_args()
            a = pa.proj
            a.B, a.C, a.N = B, C, N
            a.image_width, a.image_height = int(image_width), int(image_height)
            a.camera_model = _cam(camera_model)
            a.eps2d, a.near_plane, a.far_plane, a.radius_clip = eps2d, near_plane, far_plane, radius_clip
            a.means, a.covars = _ptr(means), _ptr(covars)
            a.quats, a.scales = _ptr(None if covars is not None else quats), _ptr(None if covars is not None else scales)
            a.opacities, a.viewmats, a.Ks = _ptr(opacities), _ptr(viewmats), _ptr(Ks)
            if rigid is not None:
                rigid.fill(a.rigid)
            a.radii, a.means2d, a.depths, a.conics = _ptr(radii), _ptr(means2d), _ptr(depths), _ptr(conics)
            a.compensations = _ptr(compensations)
            if _sh is not None:  # view-dependent colours of the packed rows, evaluated by the same kernel (extension)
                _check(_sh[0], "sh coefficients", torch.float32, None, dev)
                a.sh_coeffs, a.sh_colors = _ptr(_sh[0]), _ptr(sh_colors)
                a.sh_degree, a.sh_K = int(_sh[1]), _sh[0].shape[-2]
            pa.capacity = cap
            pa.indptr = _ptr(indptr)
            pa.batch_ids, pa.camera_ids, pa.gaussian_ids = ids[0].data_ptr(), ids[1].data_ptr(), ids[2].data_ptr()
            pa.nnz = _ptr(nnz_dev)
            pa.workspace = _ptr(ws)
            _lib.check(lib.rs_project_packed_fwd(ctypes.byref(pa), _stream()))
            nnz = int(nnz_dev.item())
            if nnz <= cap:
                break
            cap = nnz  # the caller's capacity hint was too small: run again with room for every row
    comp = compensations[:nnz] if compensations is not None else None
    out = (indptr, ids[0, :nnz], ids[1, :nnz], ids[2, :nnz], radii[:nnz], means2d[:nnz], depths[:nnz], conics[:nnz], comp)
    return out + (sh_colors[:nnz],) if _sh is not None else out


def projection_ewa_3dgs_packed_bwd(
    means: Tensor,
    covars: Optional[Tensor],
    quats: Optional[Tensor],
    scales: Optional[Tensor],
    viewmats: Tensor,
    Ks: Tensor,
    image_width: int,
    image_height: int,
    eps2d: float,
    camera_model: int,
    batch_ids: Tensor,  # [nnz] int64
    camera_ids: Tensor,  # [nnz] int64
    gaussian_ids: Tensor,  # [nnz] int64
    conics: Tensor,  # [nnz, 3]
    compensations: Optional[Tensor],  # [nnz]
    v_means2d: Tensor,  # [nnz, 2]
    v_depths: Tensor,  # [nnz]
    v_conics: Tensor,  # [nnz, 3]
    v_compensations: Optional[Tensor],  # [nnz]
    viewmats_requires_grad: bool,
    sparse_grad: bool,
    rigid: Optional[RigidPoses] = None,
) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
    """-> (v_means, v_covars, v_quats, v_scales, v_viewmats); [nnz, ...] rows when sparse_grad, else dense accumulators
    shaped like the inputs (csrc/Projection.cpp:415-547)."""
    lib = _lib.load()
    for t, n in ((means, "means"), (viewmats, "viewmats"), (Ks, "Ks"), (conics, "conics"), (v_means2d, "v_means2d"),
                 (v_depths, "v_depths"), (v_conics, "v_conics")):
        _check(t, n, torch.float32)
    for t, n in ((batch_ids, "batch_ids"), (camera_ids, "camera_ids"), (gaussian_ids, "gaussian_ids")):
        _check(t, n, torch.int64)
    nnz = gaussian_ids.shape[0]
    N = means.shape[-2]
    C = viewmats.shape[-3]
    B = means.numel() // (N * 3) if N > 0 else 1
    dev = means.device
    if rigid is not None:
        rigid.validate(N, dev)

    def out_like(t: Tensor) -> Tensor:
        if sparse_grad:
            return torch.zeros((nnz,) + tuple(t.shape[-1:]), dtype=t.dtype, device=dev)
        return torch.zeros_like(t)

    with torch.cuda.device(dev):
        v_means = out_like(means)
        v_covars = v_quats = v_scales = None
        if covars is not None:
            _check(covars, "covars", torch.float32)
            v_covars = out_like(covars)
        else:
            _check(quats, "quats", torch.float32)
            _check(scales, "scales", torch.float32)
            v_quats = out_like(quats)
            v_scales = out_like(scales)
        v_viewmats = torch.zeros_like(viewmats) if viewmats_requires_grad else None
        if nnz == 0:  # no visible pair: every gradient is zero (the row pointers are empty)
            return v_means, v_covars, v_quats, v_scales, v_viewmats
        a = _lib.rs_project_bwd_args()
        a.B, a.C, a.N = B, C, N
        a.image_width, a.image_height = int(image_width), int(image_height)
        a.camera_model = _cam(camera_model)
        a.eps2d = eps2d
        a.means, a.covars = _ptr(means), _ptr(covars)
        a.quats, a.scales = _ptr(None if covars is not None else quats), _ptr(None if covars is not None else scales)
        a.viewmats, a.Ks = _ptr(viewmats), _ptr(Ks)
        if rigid is not None:
            rigid.fill(a.rigid)
        a.radii, a.conics, a.compensations = None, _ptr(conics), _ptr(compensations)
        a.v_means2d, a.v_depths, a.v_conics = _ptr(v_means2d), _ptr(v_depths), _ptr(v_conics)
        a.v_compensations = _ptr(v_compensations)
        a.v_means, a.v_covars, a.v_quats, a.v_scales = _ptr(v_means), _ptr(v_covars), _ptr(v_quats), _ptr(v_scales)
        a.v_viewmats = _ptr(v_viewmats)
        a.batch_ids, a.camera_ids, a.gaussian_ids = _ptr(batch_ids), _ptr(camera_ids), _ptr(gaussian_ids)
        a.nnz = nnz
        a.sparse_grad = 1 if sparse_grad else 0
        _lib.check(lib.rs_project_bwd(ctypes.byref(a), _stream()))
    return v_means, v_covars, v_quats, v_scales, v_viewmats


# ---------------------------------------------------------------------------------------------------------------------
# tile intersection (Ops.h:186-204; csrc/Intersect.cpp:15-168)
# ---------------------------------------------------------------------------------------------------------------------
# The per-tile offsets fall out of the depth-ordered binning for free (32-bit tile keys, rs_isect_sorted.tile_offsets).  They
# are kept for the one isect_ids tensor the last sorted intersect_tile() returned; intersect_offset() hands them out when it is
# called with that very tensor object, unmodified -- which is what rasterization() does (rendering.py:880-892: isect_tiles ->
# isect_offset_encode) -- instead of re-deriving them from the 64-bit ids.
_OFFSETS_CACHE: dict = {}


def _remember_offsets(isect_ids: Tensor, offsets: Tensor, I: int, tile_width: int, tile_height: int) -> None:
    import weakref

    _OFFSETS_CACHE.clear()
    _OFFSETS_CACHE["entry"] = (weakref.ref(isect_ids), isect_ids._version, int(I), int(tile_width), int(tile_height), offsets)


def intersect_tile(
    means2d: Tensor,  # [..., N, 2] or [nnz, 2]
    radii: Tensor,  # [..., N, 2] or [nnz, 2]
    depths: Tensor,  # [..., N] or [nnz]
    image_ids: Optional[Tensor],  # [nnz] int64 (packed)
    gaussian_ids: Optional[Tensor],  # [nnz] int64 (packed)
    I: int,
    tile_size: int,
    tile_width: int,
    tile_height: int,
    sort: bool,
    segmented: bool,
) -> Tuple[Tensor, Tensor, Tensor]:
    lib = _lib.load()
    _check(means2d, "means2d", torch.float32)
    _check(radii, "radii", torch.int32)
    _check(depths, "depths", torch.float32)
    packed = means2d.dim() == 2
    if packed:
        if image_ids is None or gaussian_ids is None:
            raise RuntimeError("When packed is set, image_ids and gaussian_ids must be provided.")
        _check(image_ids, "image_ids", torch.int64)
        _check(gaussian_ids, "gaussian_ids", torch.int64)
    dev = means2d.device
    n_elems = means2d.numel() // 2
    N = 0 if packed else means2d.shape[-2]
    with torch.cuda.device(dev):
        tiles_per_gauss = torch.empty(depths.shape, dtype=torch.int32, device=dev)
        nb = lib.rs_isect_num_blocks(n_elems)
        scratch = torch.empty(nb + 2, dtype=torch.int32, device=dev)  # block sums (+ total) + n_isects
        a = _lib.rs_isect_args()
        a.n_elems, a.N, a.I = n_elems, N, int(I)
        a.tile_size, a.tile_width, a.tile_height = int(tile_size), int(tile_width), int(tile_height)
        a.means2d, a.radii, a.depths = _ptr(means2d), _ptr(radii), _ptr(depths)
        a.image_ids = _ptr(image_ids) if packed else None
        a.tiles_per_gauss = _ptr(tiles_per_gauss)
        a.block_sums = scratch.data_ptr()
        a.n_isects = scratch.data_ptr() + 4 * (nb + 1)
        a.capacity = 0
        n_isects = 0
        s = _stream()
        footprints = None
        if n_elems > 0:
            if sort:  # counts + one 16-byte tile footprint per row: the depth-ordered emission gathers one record per row
                footprints = torch.empty((n_elems, 4), dtype=torch.int32, device=dev)
                _lib.check(lib.rs_isect_footprints(ctypes.byref(a), None, None, footprints.data_ptr(), s))
            else:
                _lib.check(lib.rs_isect_count(ctypes.byref(a), s))
            _lib.check(lib.rs_isect_scan(ctypes.byref(a), s))
            out = ctypes.c_int64(0)
            # the one host sync of the compat path (csrc/Intersect.cpp:79-80)
            _lib.check(lib.rs_isect_count_total(ctypes.byref(a), s, ctypes.byref(out)))
            n_isects = out.value
        isect_ids = torch.empty(n_isects, dtype=torch.int64, device=dev)
        flatten_ids = torch.empty(n_isects, dtype=torch.int32, device=dev)
        if n_isects > 0:
            a.isect_ids, a.flatten_ids, a.capacity = _ptr(isect_ids), _ptr(flatten_ids), n_isects
            if not sort:
                _lib.check(lib.rs_isect_emit(ctypes.byref(a), s))
            else:
                # `segmented` sorts inside each image segment (IntersectTile.cu:368-379); the order is the same as one
                # stable sort over (image | tile | depth), which is what rs_isect_sorted produces.
                sa = _lib.rs_isect_sorted_args()
                ctypes.memmove(ctypes.byref(sa.isect), ctypes.byref(a), ctypes.sizeof(a))
                ws_bytes = lib.rs_isect_sorted_workspace_bytes(n_elems, n_isects)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                offsets = torch.empty((int(I), int(tile_height), int(tile_width)), dtype=torch.int32, device=dev)
                sa.tile_offsets = offsets.data_ptr()
                sa.workspace, sa.workspace_bytes = ws.data_ptr(), ws_bytes
                sa.tile_footprints = footprints.data_ptr()
                _lib.check(lib.rs_isect_sorted(ctypes.byref(sa), s))
                _remember_offsets(isect_ids, offsets, I, tile_width, tile_height)
    return tiles_per_gauss, isect_ids, flatten_ids


def intersect_offset(isect_ids: Tensor, I: int, tile_width: int, tile_height: int) -> Tensor:
    lib = _lib.load()
    _check(isect_ids, "isect_ids", torch.int64)
    hit = _OFFSETS_CACHE.get("entry")
    if hit is not None and hit[0]() is isect_ids and hit[1] == isect_ids._version and hit[2:5] == (int(I), int(tile_width),
                                                                                                 int(tile_height)):
        return hit[5].clone()  # the caller owns its result (a second call must not alias the first)
    dev = isect_ids.device
    with torch.cuda.device(dev):
        offsets = torch.empty((I, tile_height, tile_width), dtype=torch.int32, device=dev)
        _lib.check(
            lib.rs_isect_offsets(
                isect_ids.data_ptr() if isect_ids.numel() else None,
                isect_ids.numel(), None, int(I), int(tile_width), int(tile_height), offsets.data_ptr(), _stream(),
            )
        )
    return offsets


# ---------------------------------------------------------------------------------------------------------------------
# compositing (Ops.h:223-263; csrc/Rasterization.cpp:20-228)
# ---------------------------------------------------------------------------------------------------------------------
def _fill_raster(a, means2d, conics, colors, opacities, backgrounds, masks, image_width, image_height, tile_size,
                 tile_offsets, flatten_ids, attr_mod_colors=0, attr_mod_opacities=0):
    _check(means2d, "means2d", torch.float32)
    _check(conics, "conics", torch.float32)
    _check(colors, "colors", torch.float32)
    _check(opacities, "opacities", torch.float32)
    _check(tile_offsets, "tile_offsets", torch.int32)
    _check(flatten_ids, "flatten_ids", torch.int32)
    if backgrounds is not None:
        _check(backgrounds, "backgrounds", torch.float32)
    if masks is not None:
        _check(masks, "masks", torch.bool)
    packed = means2d.dim() == 2
    tile_height, tile_width = tile_offsets.shape[-2:]
    I = tile_offsets.numel() // max(tile_height * tile_width, 1)
    a.I = I
    a.N = 0 if packed else means2d.shape[-2]
    a.channels = colors.shape[-1]
    a.image_width, a.image_height, a.tile_size = int(image_width), int(image_height), int(tile_size)
    a.tile_width, a.tile_height = tile_width, tile_height
    a.n_isects = flatten_ids.shape[0]
    a.n_isects_dev = None
    a.means2d, a.conics, a.colors, a.opacities = _ptr(means2d), _ptr(conics), _ptr(colors), _ptr(opacities)
    a.backgrounds, a.masks = _ptr(backgrounds), _ptr(masks)
    a.tile_offsets, a.flatten_ids = _ptr(tile_offsets), _ptr(flatten_ids)
    a.attr_mod_colors, a.attr_mod_opacities = attr_mod_colors, attr_mod_opacities
    a.n_rows = means2d.numel() // 2
    return I


def rasterize_to_pixels_3dgs_fwd(
    means2d: Tensor,  # [..., N, 2] or [nnz, 2]
    conics: Tensor,  # [..., N, 3]
    colors: Tensor,  # [..., N, channels]
    opacities: Tensor,  # [..., N]
    backgrounds: Optional[Tensor],  # [..., channels]
    masks: Optional[Tensor],  # [..., tile_height, tile_width] bool
    image_width: int,
    image_height: int,
    tile_size: int,
    tile_offsets: Tensor,  # [..., tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
    _attr_mod_colors: int = 0,
    _attr_mod_opacities: int = 0,
) -> Tuple[Tensor, Tensor, Tensor]:
    lib = _lib.load()
    dev = means2d.device
    with torch.cuda.device(dev):
        a = _lib.rs_raster_fwd_args()
        _fill_raster(a, means2d, conics, colors, opacities, backgrounds, masks, image_width, image_height, tile_size,
                     tile_offsets, flatten_ids, _attr_mod_colors, _attr_mod_opacities)
        image_dims = tuple(tile_offsets.shape[:-2])
        channels = colors.shape[-1]
        renders = torch.empty(image_dims + (image_height, image_width, channels), dtype=torch.float32, device=dev)
        alphas = torch.empty(image_dims + (image_height, image_width, 1), dtype=torch.float32, device=dev)
        last_ids = torch.empty(image_dims + (image_height, image_width), dtype=torch.int32, device=dev)
        a.render_colors, a.render_alphas, a.last_ids = _ptr(renders), _ptr(alphas), _ptr(last_ids)
        # staging records (32 B per projected splat), packed by rs_raster_fwd itself on this path
        records = torch.empty((max(a.n_rows, 1), 8), dtype=torch.float32, device=dev)
        a.records, a.records_ready = records.data_ptr(), 0
        counter = torch.zeros(1, dtype=torch.int32, device=dev)  # work counter of the persistent compositing kernel
        a.tile_counter = counter.data_ptr()
        _lib.check(lib.rs_raster_fwd(ctypes.byref(a), _stream()))
    return renders, alphas, last_ids


def rasterize_to_pixels_3dgs_bwd(
    means2d: Tensor,
    conics: Tensor,
    colors: Tensor,
    opacities: Tensor,
    backgrounds: Optional[Tensor],
    masks: Optional[Tensor],
    image_width: int,
    image_height: int,
    tile_size: int,
    tile_offsets: Tensor,
    flatten_ids: Tensor,
    render_alphas: Tensor,
    last_ids: Tensor,
    v_render_colors: Tensor,
    v_render_alphas: Tensor,
    absgrad: bool,
    _attr_mod_colors: int = 0,
    _attr_mod_opacities: int = 0,
    _ring: bool = True,  # False: no record scratch -> the barrier-per-batch kernel (kept for C callers without scratch)
) -> Tuple[Optional[Tensor], Tensor, Tensor, Tensor, Tensor]:
    lib = _lib.load()
    dev = means2d.device
    _check(render_alphas, "render_alphas", torch.float32)
    _check(last_ids, "last_ids", torch.int32)
    _check(v_render_colors, "v_render_colors", torch.float32)
    _check(v_render_alphas, "v_render_alphas", torch.float32)
    with torch.cuda.device(dev):
        a = _lib.rs_raster_bwd_args()
        _fill_raster(a.f, means2d, conics, colors, opacities, backgrounds, masks, image_width, image_height,
                     tile_size, tile_offsets, flatten_ids, _attr_mod_colors, _attr_mod_opacities)
        a.f.render_alphas, a.f.last_ids = _ptr(render_alphas), _ptr(last_ids)
        # staging records of the batch ring (32 B per projected splat), packed by rs_raster_bwd itself
        if _ring:
            records = torch.empty((max(a.f.n_rows, 1), 8), dtype=torch.float32, device=dev)
            a.f.records, a.f.records_ready = records.data_ptr(), 0
        v_means2d = torch.zeros_like(means2d)
        v_conics = torch.zeros_like(conics)
        v_colors = torch.zeros_like(colors)
        v_opacities = torch.zeros_like(opacities)
        v_means2d_abs = torch.zeros_like(means2d) if absgrad else None
        a.v_render_colors, a.v_render_alphas = _ptr(v_render_colors), _ptr(v_render_alphas)
        a.v_means2d_abs, a.v_means2d, a.v_conics = _ptr(v_means2d_abs), _ptr(v_means2d), _ptr(v_conics)
        a.v_colors, a.v_opacities = _ptr(v_colors), _ptr(v_opacities)
        _lib.check(lib.rs_raster_bwd(ctypes.byref(a), _stream()))
    return v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities


# ---------------------------------------------------------------------------------------------------------------------
# spherical harmonics (Ops.h:154-168; csrc/SphericalHarmonics.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def _fill_sh(a, degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor]) -> None:
    _check(dirs, "dirs", torch.float32)
    _check(coeffs, "coeffs", torch.float32)
    if masks is not None:
        _check(masks, "masks", torch.bool)
    a.n = dirs.numel() // 3
    a.degree, a.K = int(degree), coeffs.shape[-2]
    a.dirs, a.coeffs, a.masks = _ptr(dirs), _ptr(coeffs), _ptr(masks)


def spherical_harmonics_fwd(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor]) -> Tensor:
    lib = _lib.load()
    with torch.cuda.device(dirs.device):
        a = _lib.rs_sh_args()
        _fill_sh(a, degrees_to_use, dirs, coeffs, masks)
        colors = torch.empty_like(dirs)
        a.colors = _ptr(colors)
        _lib.check(lib.rs_sh_fwd(ctypes.byref(a), _stream()))
    return colors


def spherical_harmonics_bwd(K: int, degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor],
                            v_colors: Tensor, compute_v_dirs: bool) -> Tuple[Tensor, Optional[Tensor]]:
    lib = _lib.load()
    _check(v_colors, "v_colors", torch.float32)
    assert coeffs.shape[-2] == K, (coeffs.shape, K)
    with torch.cuda.device(dirs.device):
        a = _lib.rs_sh_args()
        _fill_sh(a, degrees_to_use, dirs, coeffs, masks)
        v_coeffs = torch.empty_like(coeffs)
        v_dirs = torch.empty_like(dirs) if compute_v_dirs else None
        a.v_colors, a.v_coeffs, a.v_dirs = _ptr(v_colors), _ptr(v_coeffs), _ptr(v_dirs)
        _lib.check(lib.rs_sh_bwd(ctypes.byref(a), _stream()))
    return v_coeffs, v_dirs
