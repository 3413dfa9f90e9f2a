"""FrameRenderer: the per-frame animate-and-render loop with no host synchronisation.

This is the executable form of the commented-out animation loop of the reference (main.py:357-409): per frame, a new
pose for every rigid body, then one render.  The reference does `apply_transform()` per body (cloning all splat
tensors, main.py:200) followed by `rasterization()` with its `.item()` sync (csrc/Intersect.cpp:80); here one C-ABI call
(`rs_render_frame`, include/rigidsplat.h) enqueues the whole pipeline on the current stream into a pre-sized workspace.

Colours are post-activation per-Gaussian values [N, D] (sh_degree=None in the reference's terms), or, with `sh_degree`,
SH coefficients [N, K, 3] evaluated inside the projection kernel for the moved means (what main.py:328-339 renders with).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._C import PINHOLE, RigidPoses, _check


class FrameRenderer:
    def __init__(
        self,
        means: Tensor,  # [N,3]
        quats: Tensor,  # [N,4]
        scales: Tensor,  # [N,3]
        opacities: Tensor,  # [N]
        colors: Tensor,  # [N,D]
        width: int,
        height: int,
        cluster_ids: Optional[Tensor] = None,  # [N] int32
        body_centers: Optional[Tensor] = None,  # [K,3]
        n_cameras: int = 1,
        max_isects: Optional[int] = None,
        near_plane: float = 0.01,
        far_plane: float = 1e10,
        radius_clip: float = 0.0,
        eps2d: float = 0.3,
        backgrounds: Optional[Tensor] = None,  # [C,D]
        camera_model: int = PINHOLE,
        sh_degree: Optional[int] = None,  # with it, `colors` holds SH coefficients [N,K,3] (main.py renders with degree 3)
        rgb8: bool = False,  # also keep the frame as uint8 [C,H,W,3] (`render_rgb8`), quantised like save_rendered_image
        tight_tiles: bool = False,  # list a (tile, splat) pair only where the splat can reach alpha >= 1/255 inside the
        # tile (rs_frame_args.tight_tiles): bit-identical images from ~20 % fewer intersections; meta() then returns a
        # subset of the reference's lists.  False: exactly the reference's lists.
    ):
        self.lib = _lib.load()
        dev = means.device
        N = means.shape[0]
        D = colors.shape[-1]
        self.sh_degree = sh_degree
        if sh_degree is not None:
            assert colors.dim() == 3 and D == 3 and (sh_degree + 1) ** 2 <= colors.shape[1], colors.shape
        _check(means, "means", torch.float32, (N, 3))
        _check(quats, "quats", torch.float32, (N, 4), dev)
        _check(scales, "scales", torch.float32, (N, 3), dev)
        _check(opacities, "opacities", torch.float32, (N,), dev)
        _check(colors, "colors", torch.float32, (N, D) if sh_degree is None else (N, colors.shape[1], 3), dev)
        if cluster_ids is not None:
            _check(cluster_ids, "cluster_ids", torch.int32, (N,), dev)
        if backgrounds is not None:
            _check(backgrounds, "backgrounds", torch.float32, (n_cameras, D), dev)
        self.means, self.quats, self.scales, self.opacities, self.colors = means, quats, scales, opacities, colors
        self.cluster_ids, self.body_centers, self.backgrounds = cluster_ids, body_centers, backgrounds
        self.N, self.D, self.C, self.W, self.H = N, D, n_cameras, int(width), int(height)
        self.near_plane, self.far_plane, self.radius_clip, self.eps2d = near_plane, far_plane, radius_clip, eps2d
        self.camera_model = camera_model
        self.tight_tiles = bool(tight_tiles)
        self.device = dev
        self.tile_size = 16
        self.tile_width = (self.W + 15) // 16
        self.tile_height = (self.H + 15) // 16
        if max_isects is None:
            max_isects = max(16 * N * n_cameras, 1 << 20)
        self._alloc(int(max_isects))
        with torch.cuda.device(dev):
            self.render_colors = torch.empty(self.C, self.H, self.W, D, dtype=torch.float32, device=dev)
            self.render_alphas = torch.empty(self.C, self.H, self.W, 1, dtype=torch.float32, device=dev)
            self.status = torch.zeros(4, dtype=torch.int32, device=dev)
            # the 8-bit frame the reference's loop writes to disk (main.py:140-171 -> torchvision save_image), produced by
            # the compositing epilogue instead of four torch passes over the float image
            self.render_rgb8 = torch.empty(self.C, self.H, self.W, 3, dtype=torch.uint8, device=dev) if rgb8 else None
        if rgb8:
            assert D >= 3, "rgb8 output needs at least 3 colour channels"
        self._args = None

    def _alloc(self, max_isects: int) -> None:
        self.max_isects = max_isects
        nbytes = self.lib.rs_frame_workspace_bytes(self.C, self.N, self.W, self.H, self.tile_size, self.D, max_isects)
        if nbytes == 0:
            raise _lib.RigidSplatError("rs_frame_workspace_bytes: invalid geometry")
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.workspace_bytes = nbytes

    def _fill(self, viewmats: Tensor, Ks: Tensor, body_quats: Optional[Tensor], body_trans: Optional[Tensor]):
        a = _lib.rs_frame_args()
        p = a.proj
        p.B, p.C, p.N = 1, self.C, self.N
        p.image_width, p.image_height = self.W, self.H
        p.camera_model = self.camera_model
        p.eps2d, p.near_plane, p.far_plane, p.radius_clip = self.eps2d, self.near_plane, self.far_plane, self.radius_clip
        p.means, p.quats, p.scales = self.means.data_ptr(), self.quats.data_ptr(), self.scales.data_ptr()
        p.covars = None
        p.opacities = self.opacities.data_ptr()
        p.viewmats, p.Ks = viewmats.data_ptr(), Ks.data_ptr()
        if self.cluster_ids is not None and body_quats is not None:
            RigidPoses(self.cluster_ids, body_quats, body_trans, self.body_centers).fill(p.rigid)
        p.tile_size, p.tile_width, p.tile_height = self.tile_size, self.tile_width, self.tile_height
        if self.sh_degree is None:
            a.colors = self.colors.data_ptr()
        else:  # colours are evaluated inside the projection kernel from the SH coefficients
            a.colors = None
            p.sh_coeffs, p.sh_degree, p.sh_K = self.colors.data_ptr(), int(self.sh_degree), self.colors.shape[1]
        a.channels = self.D
        a.colors_per_camera = 0
        a.backgrounds = self.backgrounds.data_ptr() if self.backgrounds is not None else None
        a.max_isects = self.max_isects
        a.workspace = self.workspace.data_ptr()
        a.workspace_bytes = self.workspace_bytes
        a.render_colors = self.render_colors.data_ptr()
        a.render_alphas = self.render_alphas.data_ptr()
        a.status = self.status.data_ptr()
        a.render_rgb8 = self.render_rgb8.data_ptr() if self.render_rgb8 is not None else None
        a.out_tile_offsets = None
        a.tight_tiles = 1 if self.tight_tiles else 0
        return a

    def render(self, viewmats: Tensor, Ks: Tensor, body_quats: Optional[Tensor] = None,
               body_trans: Optional[Tensor] = None, composite_stream: Optional[torch.cuda.Stream] = None
               ) -> Tuple[Tensor, Tensor]:
        """Enqueue one frame on the current stream.  Returns views of the renderer-owned output buffers
        (render_colors [C,H,W,D], render_alphas [C,H,W,1]); they are overwritten by the next call.

        With `composite_stream`, projection + binning go to the current stream and compositing to `composite_stream`
        (ordered after them); the frame is complete when `composite_stream` is."""
        _check(viewmats, "viewmats", torch.float32, (self.C, 4, 4), self.device)
        _check(Ks, "Ks", torch.float32, (self.C, 3, 3), self.device)
        if self.cluster_ids is not None:
            if body_quats is None or body_trans is None:
                raise RuntimeError("FrameRenderer.render: body_quats and body_trans are required with cluster_ids")
            K = body_quats.shape[0]
            _check(body_quats, "body_quats", torch.float32, (K, 4), self.device)
            _check(body_trans, "body_trans", torch.float32, (K, 3), self.device)
        with torch.cuda.device(self.device):
            a = self._fill(viewmats, Ks, body_quats, body_trans)
            self._args = a
            cur = torch.cuda.current_stream()
            if composite_stream is None:
                _lib.check(self.lib.rs_render_frame(ctypes.byref(a), cur.cuda_stream))
            else:
                a.stages = 1  # RS_FRAME_BIN
                _lib.check(self.lib.rs_render_frame(ctypes.byref(a), cur.cuda_stream))
                composite_stream.wait_stream(cur)
                a.stages = 2  # RS_FRAME_COMPOSITE
                _lib.check(self.lib.rs_render_frame(ctypes.byref(a), composite_stream.cuda_stream))
                a.stages = 0
        return self.render_colors, self.render_alphas

    def render_timed(self, viewmats: Tensor, Ks: Tensor, body_quats: Optional[Tensor] = None,
                     body_trans: Optional[Tensor] = None):
        """Measurement aid: one frame with CUDA events between the stages (synchronises).  Returns
        (project_ms, binning_ms, composite_ms, frame_ms)."""
        with torch.cuda.device(self.device):
            a = self._fill(viewmats, Ks, body_quats, body_trans)
            self._args = a
            ms = (ctypes.c_float * 4)()
            _lib.check(self.lib.rs_render_frame_timed(ctypes.byref(a), torch.cuda.current_stream().cuda_stream, ms))
        return tuple(float(x) for x in ms)

    # ---- introspection (each of these synchronises) ---------------------------------------------------------------
    def n_isects(self) -> int:
        return int(self.status[0].item())

    def overflowed(self) -> bool:
        return bool(self.status[1].item())

    def ensure_capacity(self, slack: float = 1.25) -> bool:
        """After a frame: grow the workspace if it overflowed.  Returns True if it was re-allocated (re-render then)."""
        n = self.n_isects()
        if n > self.max_isects:
            self._alloc(int(n * slack) + 1024)
            return True
        return False

    def _ws_tensor(self, which: int, dtype, numel: int) -> Tensor:
        ptr = self.lib.rs_frame_workspace_ptr(ctypes.byref(self._args), which)
        off = ptr - self.workspace.data_ptr()
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off : off + nbytes].view(dtype)

    def meta(self) -> dict:
        """The sorted intersection data of the last frame (`meta` of rendering.py:651-665), as views of the workspace."""
        assert self._args is not None, "render() first"
        n = min(self.n_isects(), self.max_isects)
        E = self.C * self.N
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rs_frame_export_isect_ids(ctypes.byref(self._args),
                                                          torch.cuda.current_stream().cuda_stream))
        return {
            "isect_ids": self._ws_tensor(0, torch.int64, self.max_isects)[:n],
            "flatten_ids": self._ws_tensor(1, torch.int32, self.max_isects)[:n],
            "isect_offsets": self._ws_tensor(2, torch.int32, self.C * self.tile_height * self.tile_width).view(
                self.C, self.tile_height, self.tile_width),
            "last_ids": self._ws_tensor(3, torch.int32, self.C * self.H * self.W).view(self.C, self.H, self.W),
            "tiles_per_gauss": self._ws_tensor(4, torch.int32, E).view(self.C, self.N),
            "radii": self._ws_tensor(5, torch.int32, E * 2).view(self.C, self.N, 2),
            "means2d": self._ws_tensor(6, torch.float32, E * 2).view(self.C, self.N, 2),
            "depths": self._ws_tensor(7, torch.float32, E).view(self.C, self.N),
            "conics": self._ws_tensor(8, torch.float32, E * 3).view(self.C, self.N, 3),
            "n_isects": n,
            "tile_width": self.tile_width,
            "tile_height": self.tile_height,
        }


class FramePipeline:
    """Several frames in flight on one GPU.

    Animation frames are independent units (that is also how they shard across GPUs).  At 1 M
    Gaussians most stages of a frame are short latency-bound kernels that leave most of the 148 SMs idle, while
    compositing saturates the issue slots; giving every in-flight frame its own stream and workspace lets the binning of
    frame f+1 run underneath the compositing of frame f.  `depth` FrameRenderers share the (read-only) Gaussian tensors.

        pipe = FramePipeline(3, means, quats, scales, opacities, colors, W, H, cluster_ids=ids, body_centers=c)
        for f in range(F):
            img, alpha, done = pipe.submit(viewmats, Ks, body_quats[f], body_trans[f])   # enqueue only
            ...                                                                           # consume after done.wait()
        pipe.join()                                                                       # current stream waits for all
    The returned buffers belong to the renderer that drew the frame and are overwritten `depth` submissions later.
    """

    def __init__(self, depth: int, *args, split: bool = True, **kwargs):
        assert depth >= 1
        self.renderers = [FrameRenderer(*args, **kwargs) for _ in range(depth)]
        dev = self.renderers[0].device
        with torch.cuda.device(dev):
            self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
            # projection + binning of a frame run on a HIGH-priority stream: their short kernels are then dispatched as
            # soon as a slot frees instead of queueing behind the thousands of compositing CTAs of the frames in front
            self.bin_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(depth)] if split else None
        self.device = dev
        self.count = 0

    def submit(self, viewmats: Tensor, Ks: Tensor, body_quats: Optional[Tensor] = None,
               body_trans: Optional[Tensor] = None):
        k = self.count % len(self.renderers)
        self.count += 1
        stream = self.streams[k]
        if self.bin_streams is None:
            stream.wait_stream(torch.cuda.current_stream(self.device))  # the frame's inputs were produced there
            with torch.cuda.stream(stream):
                img, alpha = self.renderers[k].render(viewmats, Ks, body_quats, body_trans)
                done = torch.cuda.Event()
                done.record(stream)
            return img, alpha, done
        front = self.bin_streams[k]
        front.wait_stream(torch.cuda.current_stream(self.device))  # the frame's inputs were produced there
        front.wait_stream(stream)  # the slot's previous frame has been composited out of this workspace
        with torch.cuda.stream(front):
            img, alpha = self.renderers[k].render(viewmats, Ks, body_quats, body_trans, composite_stream=stream)
        done = torch.cuda.Event()
        done.record(stream)
        return img, alpha, done

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def overflowed(self) -> bool:
        return any(r.overflowed() for r in self.renderers)
