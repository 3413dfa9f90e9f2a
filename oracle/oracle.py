"""numpy front-end of oracle.c -- the CPU restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package.  See oracle.c for the reference file:line each function follows and for how parity is
pinned (tests/golden/ fixtures generated from the reference's own python functions; reference CUDA extension under
oracle/_ref/ on the GPU box).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB = os.path.join(_OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -ffp-contract=off (explicit fmaf only) -> oracle/_build/liboracle.so"""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _LIB, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


_lib = None
_f32p = ctypes.POINTER(ctypes.c_float)


class _ProjectArgs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "C", "N", "width", "height", "camera_model")] + [
        (n, ctypes.c_float) for n in ("eps2d", "near_plane", "far_plane", "radius_clip")
    ] + [(n, ctypes.c_void_p) for n in (
        "means", "covars", "quats", "scales", "opacities", "viewmats", "Ks", "radii", "means2d", "depths", "conics",
        "compensations", "ambiguous")]


class _RasterArgs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("I", "channels", "width", "height", "tile_size", "tile_width",
                                              "tile_height")] + [("n_isects", ctypes.c_int64)] + [
        (n, ctypes.c_void_p) for n in ("means2d", "conics", "colors", "opacities", "backgrounds", "masks",
                                       "tile_offsets", "flatten_ids")
    ] + [("attr_mod_colors", ctypes.c_int32), ("attr_mod_opacities", ctypes.c_int32)] + [
        (n, ctypes.c_void_p) for n in ("render_colors", "render_alphas", "last_ids", "margin")]


class _RasterBwdArgs(ctypes.Structure):
    _fields_ = [("f", _RasterArgs)] + [(n, ctypes.c_void_p) for n in (
        "v_render_colors", "v_render_alphas", "v_means2d_abs", "v_means2d", "v_conics", "v_colors", "v_opacities")]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_isect_count.restype = ctypes.c_int64
        L.orc_isect_count.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        L.orc_isect_emit.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                     ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_sort_pairs.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_isect_offsets.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_void_p]
        L.orc_rigid_transform.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
        L.orc_project.argtypes = [ctypes.POINTER(_ProjectArgs)]
        L.orc_raster_fwd.argtypes = [ctypes.POINTER(_RasterArgs)]
        L.orc_raster_bwd.argtypes = [ctypes.POINTER(_RasterBwdArgs)]
        _lib = L
    return _lib


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(n)


def _c(a, dtype) -> Optional[np.ndarray]:
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def bit_width(x: int) -> int:
    return int(x).bit_length()


# ---------------------------------------------------------------------------------------------------------------------
def rigid_transform(means, quats, cluster_ids, body_quats, body_trans, body_centers=None):
    """main.py:173-228 generalised to per-Gaussian cluster ids.  Returns (means', quats')."""
    means = _c(means, np.float32)
    quats = _c(quats, np.float32)
    ids = _c(cluster_ids, np.int32)
    bq, bt, bc = _c(body_quats, np.float32), _c(body_trans, np.float32), _c(body_centers, np.float32)
    N = means.shape[0]
    om = np.empty_like(means)
    oq = np.empty_like(quats) if quats is not None else None
    lib().orc_rigid_transform(N, _p(means), _p(quats), _p(ids), bq.shape[0], _p(bq), _p(bt), _p(bc), _p(om), _p(oq))
    return om, oq


def project(means, quats, scales, opacities, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
            far_plane=1e10, radius_clip=0.0, calc_compensations=False, camera_model=0, covars=None) -> Dict:
    """csrc/ProjectionEWA3DGSFused.cu:41-212.  Arrays may carry leading batch dims ([..., N, 3], [..., C, 4, 4])."""
    means = _c(means, np.float32)
    viewmats, Ks = _c(viewmats, np.float32), _c(Ks, np.float32)
    N, C = means.shape[-2], viewmats.shape[-3]
    B = int(np.prod(means.shape[:-2], dtype=np.int64)) if means.ndim > 2 else 1
    bd = tuple(means.shape[:-2])
    quats, scales, covars = _c(quats, np.float32), _c(scales, np.float32), _c(covars, np.float32)
    opacities = _c(opacities, np.float32)
    out = {
        "radii": np.empty(bd + (C, N, 2), np.int32),
        "means2d": np.empty(bd + (C, N, 2), np.float32),
        "depths": np.empty(bd + (C, N), np.float32),
        "conics": np.empty(bd + (C, N, 3), np.float32),
        "compensations": np.empty(bd + (C, N), np.float32) if calc_compensations else None,
        "ambiguous": np.empty(bd + (C, N), np.uint8),
    }
    a = _ProjectArgs()
    a.B, a.C, a.N, a.width, a.height, a.camera_model = B, C, N, int(width), int(height), int(camera_model)
    a.eps2d, a.near_plane, a.far_plane, a.radius_clip = eps2d, near_plane, far_plane, radius_clip
    a.means, a.covars = _p(means), _p(covars)
    a.quats, a.scales = (None, None) if covars is not None else (_p(quats), _p(scales))
    a.opacities, a.viewmats, a.Ks = _p(opacities), _p(viewmats), _p(Ks)
    a.radii, a.means2d, a.depths, a.conics = (_p(out[k]) for k in ("radii", "means2d", "depths", "conics"))
    a.compensations, a.ambiguous = _p(out["compensations"]), _p(out["ambiguous"])
    lib().orc_project(ctypes.byref(a))
    return out


def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True, n_images=None,
                image_ids=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """csrc/Intersect.cpp:15-149: (tiles_per_gauss, isect_ids, flatten_ids); stable sort on the key bits."""
    means2d, radii, depths = _c(means2d, np.float32), _c(radii, np.int32), _c(depths, np.float32)
    packed = means2d.ndim == 2 and image_ids is not None
    n_elems = means2d.size // 2
    if packed:
        I, N = int(n_images), 1
        image_ids = _c(image_ids, np.int64)
    else:
        N = means2d.shape[-2]
        I = n_elems // N if N else 0
    tpg = np.empty(depths.shape, np.int32)
    n = lib().orc_isect_count(n_elems, _p(means2d), _p(radii), tile_size, tile_width, tile_height, _p(tpg))
    ids = np.empty(n, np.int64)
    flat = np.empty(n, np.int32)
    if n:
        lib().orc_isect_emit(n_elems, max(N, 1), I, _p(means2d), _p(radii), _p(depths), _p(image_ids) if packed else None,
                             tile_size, tile_width, tile_height, _p(ids), _p(flat))
        if sort:
            end_bit = 32 + bit_width(tile_width * tile_height) + bit_width(I)
            lib().orc_sort_pairs(n, end_bit, _p(ids), _p(flat))
    return tpg, ids, flat


def isect_offset_encode(isect_ids, n_images, tile_width, tile_height) -> np.ndarray:
    """csrc/IntersectTile.cu:209-257."""
    isect_ids = _c(isect_ids, np.int64)
    off = np.empty((n_images, tile_height, tile_width), np.int32)
    lib().orc_isect_offsets(isect_ids.size, _p(isect_ids), n_images, tile_width, tile_height, _p(off))
    return off


def _raster_args(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size, tile_offsets,
                 flatten_ids, attr_mod_colors, attr_mod_opacities):
    keep = {}
    keep["means2d"], keep["conics"] = _c(means2d, np.float32), _c(conics, np.float32)
    keep["colors"], keep["opacities"] = _c(colors, np.float32), _c(opacities, np.float32)
    keep["backgrounds"] = _c(backgrounds, np.float32)
    keep["masks"] = _c(masks, np.uint8)
    keep["tile_offsets"], keep["flatten_ids"] = _c(tile_offsets, np.int32), _c(flatten_ids, np.int32)
    th, tw = keep["tile_offsets"].shape[-2:]
    I = keep["tile_offsets"].size // (th * tw)
    a = _RasterArgs()
    a.I, a.channels, a.width, a.height = I, keep["colors"].shape[-1], int(width), int(height)
    a.tile_size, a.tile_width, a.tile_height = int(tile_size), tw, th
    a.n_isects = keep["flatten_ids"].size
    for k in ("means2d", "conics", "colors", "opacities", "backgrounds", "masks", "tile_offsets", "flatten_ids"):
        setattr(a, k, _p(keep[k]))
    a.attr_mod_colors, a.attr_mod_opacities = attr_mod_colors, attr_mod_opacities
    return a, keep, I


def rasterize_fwd(means2d, conics, colors, opacities, width, height, tile_size, tile_offsets, flatten_ids,
                  backgrounds=None, masks=None, attr_mod_colors=0, attr_mod_opacities=0):
    """csrc/RasterizeToPixels3DGSFwd.cu:62-187 -> (render_colors [I,H,W,D], alphas [I,H,W,1], last_ids [I,H,W],
    margin [I,H,W] = smallest relative distance of any threshold decision from flipping)."""
    a, keep, I = _raster_args(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
                              tile_offsets, flatten_ids, attr_mod_colors, attr_mod_opacities)
    D = a.channels
    rc = np.zeros((I, height, width, D), np.float32)
    ra = np.zeros((I, height, width, 1), np.float32)
    li = np.zeros((I, height, width), np.int32)
    mg = np.full((I, height, width), 1e30, np.float32)
    a.render_colors, a.render_alphas, a.last_ids, a.margin = _p(rc), _p(ra), _p(li), _p(mg)
    lib().orc_raster_fwd(ctypes.byref(a))
    return rc, ra, li, mg


def rasterize_bwd(means2d, conics, colors, opacities, width, height, tile_size, tile_offsets, flatten_ids,
                  render_alphas, last_ids, v_render_colors, v_render_alphas, backgrounds=None, masks=None,
                  absgrad=False, attr_mod_colors=0, attr_mod_opacities=0):
    """csrc/RasterizeToPixels3DGSBwd.cu:106-276 -> dict of float64 gradients."""
    b = _RasterBwdArgs()
    a, keep, I = _raster_args(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
                              tile_offsets, flatten_ids, attr_mod_colors, attr_mod_opacities)
    b.f = a
    ra, li = _c(render_alphas, np.float32), _c(last_ids, np.int32)
    vrc, vra = _c(v_render_colors, np.float32), _c(v_render_alphas, np.float32)
    b.f.render_alphas, b.f.last_ids = _p(ra), _p(li)
    b.v_render_colors, b.v_render_alphas = _p(vrc), _p(vra)
    out = {
        "v_means2d": np.zeros(keep["means2d"].shape, np.float64),
        "v_conics": np.zeros(keep["conics"].shape, np.float64),
        "v_colors": np.zeros(keep["colors"].shape, np.float64),
        "v_opacities": np.zeros(keep["opacities"].shape, np.float64),
        "v_means2d_abs": np.zeros(keep["means2d"].shape, np.float64) if absgrad else None,
    }
    b.v_means2d, b.v_conics, b.v_colors, b.v_opacities = (_p(out[k]) for k in ("v_means2d", "v_conics", "v_colors",
                                                                              "v_opacities"))
    b.v_means2d_abs = _p(out["v_means2d_abs"])
    lib().orc_raster_bwd(ctypes.byref(b))
    return out


def render(means, quats, scales, opacities, colors, viewmats, Ks, width, height, cluster_ids=None, body_quats=None,
           body_trans=None, body_centers=None, backgrounds=None, tile_size=16, near_plane=0.01, far_plane=1e10,
           radius_clip=0.0, eps2d=0.3) -> Dict:
    """Whole path for one batch of cameras: apply_transform (per body) -> rasterization(packed=False, sh_degree=None)
    (main.py:366-400 + rendering.py:33-770).  colors [N, D]."""
    if cluster_ids is not None:
        means, quats = rigid_transform(means, quats, cluster_ids, body_quats, body_trans, body_centers)
    pr = project(means, quats, scales, opacities, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                 radius_clip)
    C, N = pr["depths"].shape[-2:]
    tw, th = -(-width // tile_size), -(-height // tile_size)
    tpg, ids, flat = isect_tiles(pr["means2d"], pr["radii"], pr["depths"], tile_size, tw, th)
    off = isect_offset_encode(ids, C, tw, th)
    rc, ra, li, mg = rasterize_fwd(pr["means2d"].reshape(-1, 2), pr["conics"].reshape(-1, 3), colors, opacities, width,
                                   height, tile_size, off, flat, backgrounds=backgrounds, attr_mod_colors=N,
                                   attr_mod_opacities=N)
    out = dict(pr)
    out.update(tiles_per_gauss=tpg, isect_ids=ids, flatten_ids=flat, isect_offsets=off, render_colors=rc,
               render_alphas=ra, last_ids=li, margin=mg, means=means, quats=quats)
    return out
