"""Gaussian-sharded rendering: the one exchange step of the path (BASELINE.json config c5).

Scheme (the reference's, gsplat/rendering.py:366-381 + 527-611 on top of gsplat/distributed.py:10-257): every rank owns a
shard of the Gaussians and an equal number of cameras.  Cameras are all-gathered, every rank projects ITS Gaussians to
ALL cameras, and the projected splats are sent to the rank that owns the camera; binning and compositing are then purely
local.  Backward is the transposed exchange.

What differs from the reference: it ships each attribute in its own all-to-all (radii; then means2d, depths, conics,
opacities, colours as a list; then the two id tensors, plus a count exchange when packed).  Here every projected splat is
ONE row -- [means2d 2 | depth 1 | conic 3 | opacity 1 | colour D] float32 -- sent by ONE differentiable
`all_to_all_single`, and one more for the integer columns (radii, and the ids when packed), so a frame costs two NCCL
collectives over NVLink instead of four to eight.  Frames / cameras sharding (c2, c4) needs no collective at all and does
not come through here.

`PeerSplatExchange` is the NVLink-native form of the same step for packed rows when no gradient has to flow back (the
render / animation path): no NCCL call at all -- one kernel per rank stores its rows straight into the receive arrays of
the ranks that own the cameras, over peer-mapped memory (csrc/exchange.cu).  The NCCL route above stays for training,
where autograd needs the transposed exchange.
"""
from __future__ import annotations

import ctypes
import os
import socket
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.distributed.nn.functional as dist_fn
from torch import Tensor

from . import _lib


def _world() -> Tuple[int, int]:
    assert dist.is_available() and dist.is_initialized(), "distributed=True needs an initialised process group"
    return dist.get_rank(), dist.get_world_size()


class GaussianShardExchange:
    """Bookkeeping + collectives of one distributed `rasterization()` call."""

    _layout_cache: Dict = {}

    def __init__(self, n_local: int, c_local: int, device: torch.device, group=None):
        self.group = group
        self.rank, self.world = _world()
        self.device = device
        # (Gaussians, cameras) of every rank.  Render / animation loops (no gradients) call this every frame with a static
        # scene: the table is then gathered once per (group, local sizes) instead of costing a collective and a host sync per
        # frame; with gradients enabled (training: densification changes the shard sizes) it is gathered on every call.
        key = (id(group) if group is not None else None, str(device), int(n_local), int(c_local))
        cached = None if torch.is_grad_enabled() else GaussianShardExchange._layout_cache.get(key)
        if cached is None:
            mine = torch.tensor([n_local, c_local], dtype=torch.int64, device=device)
            table = torch.empty(self.world * 2, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(table, mine, group=group)
            table = table.reshape(self.world, 2).cpu()
            cached = (table[:, 0].tolist(), table[:, 1].tolist())
            if not torch.is_grad_enabled():
                GaussianShardExchange._layout_cache[key] = cached
        self.n_per_rank: List[int] = list(cached[0])
        self.c_per_rank: List[int] = list(cached[1])
        # the reference requires the same number of cameras on every rank (rendering.py:374-375)
        assert len(set(self.c_per_rank)) == 1, f"every rank must own the same number of cameras, got {self.c_per_rank}"
        self.local_cameras = c_local
        self.total_cameras = sum(self.c_per_rank)
        self.n_local = n_local
        self.gaussian_base = sum(self.n_per_rank[: self.rank])  # first global index of this rank's shard

    # ---- cameras ---------------------------------------------------------------------------------------------------------
    def gather_cameras(self, viewmats: Tensor, Ks: Tensor) -> Tuple[Tensor, Tensor]:
        """[C_local, ...] on every rank -> [C_total, ...] in rank order (differentiable w.r.t. the local cameras)."""
        if self.world == 1:
            return viewmats, Ks
        flat = torch.cat([viewmats.reshape(self.local_cameras, 16), Ks.reshape(self.local_cameras, 9)], dim=1).contiguous()
        parts = dist_fn.all_gather(flat, group=self.group)
        every = torch.cat(list(parts), dim=0)
        return every[:, :16].reshape(-1, 4, 4), every[:, 16:].reshape(-1, 3, 3)

    # ---- projected splats ---------------------------------------------------------------------------------------------------
    def _swap(self, rows: Tensor, send: List[int], recv: List[int], differentiable: bool) -> Tensor:
        rows = rows.contiguous()
        if self.world == 1:
            return rows
        if differentiable:
            out = torch.empty((sum(recv),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
            return dist_fn.all_to_all_single(out, rows, output_split_sizes=recv, input_split_sizes=send, group=self.group)
        out = torch.empty((sum(recv),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out, rows, output_split_sizes=recv, input_split_sizes=send, group=self.group)
        return out

    def exchange(self, packed: bool, radii: Tensor, means2d: Tensor, depths: Tensor, conics: Tensor, opacities: Tensor,
                 colors: Tensor, camera_ids: Optional[Tensor], gaussian_ids: Optional[Tensor]):
        """Moves every projected splat to the rank owning its camera.

        Dense (packed=False): inputs are [C_total, N_local, ...]; outputs are [C_local, N_total, ...] with the Gaussians of
        rank 0 first, then rank 1, ... (so global Gaussian index = shard base + local index).
        Packed: inputs are [nnz, ...] rows ordered by camera; outputs are the rows of this rank's cameras from every source
        rank in rank order, `camera_ids` made local and `gaussian_ids` made global.
        Returns (radii, means2d, depths, conics, opacities, colors, camera_ids, gaussian_ids)."""
        D = colors.shape[-1]
        Cl = self.local_cameras
        if packed:
            # rows are ordered by camera, so the rows owned by rank r are the contiguous run of cameras [r*Cl, (r+1)*Cl):
            # two binary searches per rank instead of a histogram over all rows
            edges = torch.arange(self.world + 1, device=camera_ids.device, dtype=camera_ids.dtype) * Cl
            cuts = torch.searchsorted(camera_ids.contiguous(), edges)
            send_t = cuts[1:] - cuts[:-1]
            owner = torch.div(camera_ids, Cl, rounding_mode="floor")
            recv_t = torch.empty_like(send_t)
            if self.world > 1:
                dist.all_to_all_single(recv_t, send_t, group=self.group)
            else:
                recv_t.copy_(send_t)
            send, recv = send_t.tolist(), recv_t.tolist()  # the one host sync of the packed exchange
            nnz = means2d.shape[0]
            fl = torch.cat([means2d, depths.reshape(nnz, 1), conics, opacities.reshape(nnz, 1), colors.reshape(nnz, D)],
                           dim=1)
            assert sum(self.n_per_rank) < 2**31, "global Gaussian ids must fit int32 for the exchange"
            ints = torch.cat([radii.to(torch.int32), (camera_ids - owner * Cl).to(torch.int32).unsqueeze(1),
                              (gaussian_ids + self.gaussian_base).to(torch.int32).unsqueeze(1)], dim=1)
            fl = self._swap(fl, send, recv, True)
            ints = self._swap(ints, send, recv, False)
            return (ints[:, :2].contiguous(), fl[:, 0:2], fl[:, 2], fl[:, 3:6], fl[:, 6], fl[:, 7:],
                    ints[:, 2].long(), ints[:, 3].long())

        Ct, Nl = self.total_cameras, self.n_local
        assert means2d.shape[:2] == (Ct, Nl), means2d.shape
        send = [c * Nl for c in self.c_per_rank]
        recv = [Cl * n for n in self.n_per_rank]
        fl = torch.cat([means2d, depths.unsqueeze(-1), conics, opacities.unsqueeze(-1), colors.expand(Ct, Nl, D)],
                       dim=-1).reshape(Ct * Nl, 7 + D)
        fl = self._swap(fl, send, recv, True)
        ri = self._swap(radii.reshape(Ct * Nl, 2), send, recv, False)

        def regroup(t: Tensor) -> Tensor:  # chunks [Cl * N_s, k] per source s -> [Cl, sum N_s, k]
            chunks = torch.split(t, recv, dim=0)
            return torch.cat([c.reshape(Cl, n, t.shape[-1]) for c, n in zip(chunks, self.n_per_rank)], dim=1)

        fl, ri = regroup(fl), regroup(ri)
        return (ri.contiguous(), fl[..., 0:2], fl[..., 2], fl[..., 3:6], fl[..., 6], fl[..., 7:], None, None)


class _DeviceArray:
    """Raw device memory presented to torch through __cuda_array_interface__ (no copy; `owner` keeps it alive)."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}
        self._owner = owner


class _PeerBuffers:
    """One receive allocation per rank, mapped by every peer (legacy CUDA IPC over cudaMalloc memory)."""

    def __init__(self, lib, group, device: torch.device, capacity: int, channels: int):
        self.lib, self.capacity, self.channels = lib, capacity, channels
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        offsets = (ctypes.c_uint64 * 9)()
        _lib.check(lib.rs_exchange_layout(capacity, channels, offsets))
        self.offsets = list(offsets)
        own = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.rs_peer_alloc(self.offsets[-1], ctypes.byref(own)))
            handle = ctypes.create_string_buffer(64)
            _lib.check(lib.rs_peer_export(own, handle))
            cards: List[Optional[Tuple[bytes, int, int]]] = [None] * world
            dist.all_gather_object(cards, (handle.raw, capacity, channels), group=group)
            # the layout of the arrays follows from (capacity, channels): it must be the same on every rank
            assert all(c[1] == capacity and c[2] == channels for c in cards), f"peer buffers disagree: {[c[1:] for c in cards]}"
            self.own = own.value
            self.mapped: List[int] = []
            for s in range(world):
                if s == rank:
                    self.mapped.append(self.own)
                    continue
                p = ctypes.c_void_p()
                _lib.check(lib.rs_peer_open(cards[s][0], ctypes.byref(p)))
                self.mapped.append(p.value)
            self.table = torch.tensor(self.mapped, dtype=torch.int64, device=device)
            self.rank = rank
        dist.barrier(group=group)

    def column(self, i: int, rows: int, tail: Tuple[int, ...], typestr: str, dtype: torch.dtype, device) -> Tensor:
        if rows == 0:
            return torch.empty((0,) + tail, dtype=dtype, device=device)
        return torch.as_tensor(_DeviceArray(self.own + self.offsets[i], (rows,) + tail, typestr, self), device=device)

    def release(self, group) -> None:
        for s, p in enumerate(self.mapped):
            if s != self.rank:
                self.lib.rs_peer_close(ctypes.c_void_p(p))
        dist.barrier(group=group)  # nobody still maps this rank's allocation when it is freed
        self.lib.rs_peer_free(ctypes.c_void_p(self.own))
        self.mapped = []


class PeerSplatExchange:
    """Packed projected splats -> the ranks owning their cameras, through peer memory (rs_exchange_push / _wait).

    One instance per (process group, device, channel count), kept for the life of the process: the receive arrays are
    persistent and the returned tensors are views of them, valid until the next exchange() of the same instance.
    All ranks of the group must call exchange() the same number of times (it is a collective)."""

    _instances: Dict[Tuple, "PeerSplatExchange"] = {}
    _usable: Dict = {}  # process group -> bool (probed once, collectively)
    enabled: bool = True  # False routes no-grad packed calls through the NCCL all-to-all too (A/B measurements)
    differentiable: bool = True  # False keeps training (gradients flowing back) on the NCCL all-to-all
    initial_capacity: Optional[int] = None  # rows; None = this rank's row count of the first call (regrown on demand)
    timeout_ms: int = int(os.environ.get("RS_EXCHANGE_TIMEOUT_MS", "0"))  # 0 = the library default (60 s)

    @classmethod
    def usable(cls, group, device: torch.device) -> bool:
        """Collective probe, cached per group: the peer-memory route needs every rank on ONE host (CUDA IPC), peer access
        between every pair of devices and at most RS_EXCHANGE_MAX_WORLD ranks.  Anything else (multi-node jobs, more than
        16 ranks, GPUs without P2P) takes the NCCL all-to-all of GaussianShardExchange, as the reference does."""
        key = id(group) if group is not None else None
        if key in cls._usable:
            return cls._usable[key]
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ok = cls.enabled and world <= 16 and device.type == "cuda"
        cards: List = [None] * world
        dist.all_gather_object(cards, (socket.gethostname(), device.index if device.type == "cuda" else -1), group=group)
        ok = ok and len({c[0] for c in cards}) == 1 and all(c[1] >= 0 for c in cards)
        if ok:
            for s, (_, idx) in enumerate(cards):
                if s != rank and idx != device.index and not torch.cuda.can_device_access_peer(device.index, idx):
                    ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        cls._usable[key] = bool(flag.item())
        return cls._usable[key]

    @classmethod
    def get(cls, group, device: torch.device, channels: int) -> "PeerSplatExchange":
        key = (id(group) if group is not None else None, device.index, channels)
        if key not in cls._instances:
            cls._instances[key] = cls(group, device, channels)
        return cls._instances[key]

    def __init__(self, group, device: torch.device, channels: int):
        self.lib = _lib.load()
        self.group, self.device, self.channels = group, device, channels
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= 16, "PeerSplatExchange: at most 16 ranks (one NVLink domain)"
        self.epoch = 0
        self.buffers: Optional[_PeerBuffers] = None
        self.totals = torch.zeros(4, dtype=torch.int64, device=device)
        # transposed exchange (backward): a second peer-mapped allocation per rank + its own epoch counter
        self.grad_epoch = 0
        self.grad_buffers: Optional[_PeerBuffers] = None
        self.grad_status = torch.zeros(2, dtype=torch.int64, device=device)
        self._last_args = None

    def _regrow(self, capacity: int) -> None:
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # every rank's wait kernel has finished, so nobody still writes the old arrays
        if self.buffers is not None:
            self.buffers.release(self.group)
        self.buffers = _PeerBuffers(self.lib, self.group, self.device, capacity, self.channels)

    def exchange(self, cameras_per_rank: int, indptr: Tensor, camera_ids: Tensor, gaussian_ids: Tensor, radii: Tensor,
                 means2d: Tensor, depths: Tensor, conics: Tensor, compensations: Optional[Tensor], opacities: Tensor,
                 opacities_per_row: bool, colors: Tensor, colors_per_row: bool, gaussian_base: int):
        """-> (radii, means2d, depths, conics, opacities, colors, camera_ids (local), gaussian_ids (global)) of the rows
        this rank composites, ordered (source rank, camera, Gaussian)."""
        if self.buffers is None:
            # first call: every rank must size its arrays identically (the layout follows from the capacity), so start
            # from the largest row count any rank holds; the exchange regrows on demand from then on
            cap = torch.tensor([self.initial_capacity or max(int(camera_ids.shape[0]), 1024)], dtype=torch.int64,
                               device=self.device)
            dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=self.group)
            self._regrow(int(cap.item()))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        keep = [t.contiguous() for t in (indptr, camera_ids, gaussian_ids, radii, means2d, depths, conics, opacities, colors)]
        comp = compensations.contiguous() if compensations is not None else None
        while True:
            self.epoch += 1
            a = _lib.rs_exchange_args()
            a.world, a.rank, a.cameras_per_rank, a.channels = self.world, self.rank, cameras_per_rank, self.channels
            a.capacity, a.epoch = self.buffers.capacity, self.epoch
            a.colors_per_row, a.opacities_per_row = int(colors_per_row), int(opacities_per_row)
            a.timeout_ms = int(self.timeout_ms)
            a.peer_base = self.buffers.table.data_ptr()
            (a.indptr, a.camera_ids, a.gaussian_ids, a.radii, a.means2d, a.depths, a.conics, a.opacities,
             a.colors) = [t.data_ptr() for t in keep]
            a.compensations = comp.data_ptr() if comp is not None else None
            a.gaussian_base = gaussian_base
            a.nnz = int(camera_ids.shape[0])
            with torch.cuda.device(self.device):
                _lib.check(self.lib.rs_exchange_push(ctypes.byref(a), stream))
                _lib.check(self.lib.rs_exchange_wait(ctypes.byref(a), self.totals.data_ptr(), stream))
            got, worst, err, behind = self.totals.tolist()  # the one host sync of the exchange (sizes of the views)
            if err == 1:  # raised on EVERY rank of the group for this epoch (a rank that gave up sent no rows at all)
                raise RuntimeError(f"PeerSplatExchange rank {self.rank} epoch {self.epoch}: a peer did not arrive within "
                                   f"the spin limit (data flags behind: {behind & 0xffff:#x}, count flags behind: "
                                   f"{behind >> 16:#x}); set RS_EXCHANGE_TIMEOUT_MS to wait longer")
            if err == 0:
                self._last_args = a
                break
            self._regrow(int(worst * 1.25) + 1024)  # identical decision on every rank: `worst` comes from the full matrix
        b, dev = self.buffers, self.device
        return (b.column(5, got, (2,), "<i4", torch.int32, dev), b.column(0, got, (2,), "<f4", torch.float32, dev),
                b.column(1, got, (), "<f4", torch.float32, dev), b.column(2, got, (3,), "<f4", torch.float32, dev),
                b.column(3, got, (), "<f4", torch.float32, dev),
                b.column(4, got, (self.channels,), "<f4", torch.float32, dev),
                b.column(6, got, (), "<i8", torch.int64, dev), b.column(7, got, (), "<i8", torch.int64, dev))


    # ---- differentiable form: forward = the exchange above, backward = the transposed exchange over peer memory -----------
    def exchange_differentiable(self, cameras_per_rank: int, indptr: Tensor, camera_ids: Tensor, gaussian_ids: Tensor,
                                radii: Tensor, means2d: Tensor, depths: Tensor, conics: Tensor, opacities_rows: Tensor,
                                colors_rows: Tensor, gaussian_base: int):
        """Same rows as exchange() for per-row opacities / colours ([nnz], [nnz, D]), with gradients flowing back to this
        rank's packed rows through rs_exchange_push_grad (gsplat/distributed.py:243-248 without NCCL).  The returned
        tensors are copies (autograd keeps them until the backward pass; the receive arrays are reused by the next call)."""
        return _PeerExchangeFn.apply(self, cameras_per_rank, indptr, camera_ids, gaussian_ids, radii, gaussian_base,
                                     means2d, depths, conics, opacities_rows, colors_rows)

    def _read_counts(self) -> Tensor:
        counts = torch.empty(self.world * self.world, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rs_exchange_read_counts(ctypes.byref(self._last_args), counts.data_ptr(),
                                                        torch.cuda.current_stream(self.device).cuda_stream))
        return counts

    def _push_grads(self, counts: Tensor, grad_capacity: int, nnz_local: int, grads):
        """grads = (v_means2d, v_depths, v_conics, v_opacities, v_colors) of the rows this rank received -> the gradients of
        the rows it sent, [nnz_local, ...] each."""
        if self.grad_buffers is None or self.grad_buffers.capacity < grad_capacity:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            if self.grad_buffers is not None:
                self.grad_buffers.release(self.group)
            self.grad_buffers = _PeerBuffers(self.lib, self.group, self.device, int(grad_capacity * 1.25) + 1024, self.channels)
        self.grad_epoch += 1
        # a rank whose cameras saw nothing received no rows: its gradient tensors are empty (NULL data pointers); the kernel
        # then has no block to copy, but the library insists on valid pointers
        keep = [g.contiguous() if g.numel() > 0 else torch.zeros((1,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
                for g in grads]
        a = _lib.rs_exchange_grad_args()
        a.world, a.rank, a.channels, a.timeout_ms = self.world, self.rank, self.channels, int(self.timeout_ms)
        a.capacity, a.epoch = self.grad_buffers.capacity, self.grad_epoch
        a.peer_base, a.counts = self.grad_buffers.table.data_ptr(), counts.data_ptr()
        a.v_means2d, a.v_depths, a.v_conics, a.v_opacities, a.v_colors = [g.data_ptr() for g in keep]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rs_exchange_push_grad(ctypes.byref(a), stream))
            _lib.check(self.lib.rs_exchange_wait_grad(ctypes.byref(a), self.grad_status.data_ptr(), stream))
        err, behind = self.grad_status.tolist()
        if err != 0:
            raise RuntimeError(f"PeerSplatExchange rank {self.rank} backward epoch {self.grad_epoch}: transposed exchange failed "
                               f"(error {err}, sources behind {behind:#x})")
        b, dev, n = self.grad_buffers, self.device, nnz_local
        return (b.column(0, n, (2,), "<f4", torch.float32, dev).clone(), b.column(1, n, (), "<f4", torch.float32, dev).clone(),
                b.column(2, n, (3,), "<f4", torch.float32, dev).clone(), b.column(3, n, (), "<f4", torch.float32, dev).clone(),
                b.column(4, n, (self.channels,), "<f4", torch.float32, dev).clone())


class _PeerExchangeFn(torch.autograd.Function):
    """Row exchange over NVLink peer memory as an autograd node: backward is the transposed exchange."""

    @staticmethod
    def forward(ctx, peer: PeerSplatExchange, cameras_per_rank, indptr, camera_ids, gaussian_ids, radii, gaussian_base,
                means2d, depths, conics, opacities_rows, colors_rows):
        out = peer.exchange(cameras_per_rank, indptr, camera_ids, gaussian_ids, radii, means2d.detach(), depths.detach(),
                            conics.detach(), None, opacities_rows.detach(), True, colors_rows.detach(), True, gaussian_base)
        ctx.peer, ctx.nnz_local = peer, int(means2d.shape[0])
        ctx.counts = peer._read_counts()
        # every rank's gradient arrays share one layout, so they are sized for the largest row count any rank sent
        cap = torch.tensor([ctx.nnz_local], dtype=torch.int64, device=peer.device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=peer.group)
        ctx.grad_capacity = int(cap.item())
        radii_r, m2, d, con, op, col, cam, gid = (t.clone() for t in out)
        ctx.mark_non_differentiable(radii_r, cam, gid)
        return radii_r, m2, d, con, op, col, cam, gid

    @staticmethod
    def backward(ctx, _v_radii, v_m2, v_d, v_con, v_op, v_col, _v_cam, _v_gid):
        peer = ctx.peer
        got = int(ctx.counts.view(peer.world, peer.world)[:, peer.rank].sum().item())
        dev = peer.device

        def dense(v, tail):
            return v if v is not None else torch.zeros((got,) + tail, dtype=torch.float32, device=dev)

        grads = (dense(v_m2, (2,)), dense(v_d, ()), dense(v_con, (3,)), dense(v_op, ()), dense(v_col, (peer.channels,)))
        g_m2, g_d, g_con, g_op, g_col = peer._push_grads(ctx.counts, ctx.grad_capacity, ctx.nnz_local, grads)
        return (None,) * 7 + (g_m2, g_d, g_con, g_op, g_col)


class PeerRouteUnavailable(RuntimeError):
    """Raised on EVERY rank alike (the probe is collective) when the ranks cannot map each other's memory."""


class ShardedFrameRenderer:
    """One Gaussian-sharded frame (BASELINE config c5) with NO host synchronisation on the way: the distributed counterpart
    of animation.FrameRenderer.

    Same data flow as rasterization(distributed=True, packed=True) -- gsplat/rendering.py:366-381, 527-611: cameras are
    all-gathered, every rank projects ITS Gaussians to ALL cameras, the projected splats travel to the rank owning the camera,
    binning and compositing are local -- but every size stays on the device: the packed projection writes into capacity-sized
    row buffers (device-side nnz / indptr), the peer-memory exchange places rows from the device-side count matrix,
    rs_exchange_seal blanks the unused tail of the receive arrays, and tile binning (rs_isect_footprints: counts + one
    16-byte tile footprint per received row, so the depth-ordered emission gathers one record per row) / compositing run
    over the whole row capacity with the device-side intersection count (as rs_render_frame does on one GPU).  The reference reads back sizes
    three times per frame (projection nnz, all-to-all counts, intersection count); here the host reads NOTHING unless
    `check()` is called (one read of six integers: rows received, capacity needed, exchange error, intersections, overflow),
    so the ranks do not drift apart between frames.  Per-Gaussian colours [N, D] (sh_degree=None), no gradients.
    All ranks of the group must call render() the same number of times (it contains the exchange)."""

    def __init__(self, means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor, width: int,
                 height: int, cameras_per_rank: int, group=None, max_isects: Optional[int] = None,
                 row_capacity: Optional[int] = None, cluster_ids: Optional[Tensor] = None,
                 body_centers: Optional[Tensor] = None, near_plane: float = 0.01, far_plane: float = 1e10,
                 radius_clip: float = 0.0, eps2d: float = 0.3, tight_tiles: bool = False):
        self.lib = _lib.load()
        self.group = group
        # True: a received splat is binned only into the tiles where it can reach alpha >= 1/255 (rs_isect_footprints with
        # conics + opacities): the same pixels bit for bit from ~20 % fewer intersections.  False: the reference's lists.
        self.tight_tiles = bool(tight_tiles)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = means.device
        self.device = dev
        self.means, self.quats, self.scales, self.opacities, self.colors = (t.contiguous() for t in (means, quats, scales,
                                                                                                     opacities, colors))
        self.cluster_ids, self.body_centers = cluster_ids, body_centers
        self.N, self.D = int(means.shape[0]), int(colors.shape[-1])
        self.W, self.H, self.Cl = int(width), int(height), int(cameras_per_rank)
        self.Ct = self.Cl * self.world
        self.tile_w, self.tile_h = (self.W + 15) // 16, (self.H + 15) // 16
        self.near_plane, self.far_plane, self.radius_clip, self.eps2d = near_plane, far_plane, radius_clip, eps2d
        sizes = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([self.N], dtype=torch.int64, device=dev), group=group)
        self.gaussian_base = int(sizes[: self.rank].sum().item())
        if not PeerSplatExchange.usable(group, dev):
            raise PeerRouteUnavailable("ShardedFrameRenderer needs the peer-memory exchange (all ranks on one host, peer access "
                                       "between every pair of devices, at most 16 ranks); use rasterization(distributed=True), "
                                       "which falls back to NCCL")
        self.peer = PeerSplatExchange.get(group, dev, self.D)
        # packed projection: every (camera, Gaussian) pair fits, so the projection can never overflow
        cap_p = max(self.N * self.Ct, 1)
        self.cap_p = cap_p
        with torch.cuda.device(dev):
            self.p_ids = torch.empty((3, cap_p), dtype=torch.int64, device=dev)
            self.p_radii = torch.empty((cap_p, 2), dtype=torch.int32, device=dev)
            self.p_means2d = torch.empty((cap_p, 2), dtype=torch.float32, device=dev)
            self.p_depths = torch.empty((cap_p,), dtype=torch.float32, device=dev)
            self.p_conics = torch.empty((cap_p, 3), dtype=torch.float32, device=dev)
            self.p_indptr = torch.zeros(self.Ct + 1, dtype=torch.int32, device=dev)
            self.p_nnz = torch.zeros(1, dtype=torch.int64, device=dev)
            self.p_ws = torch.empty(max(int(self.lib.rs_project_packed_workspace_bytes(1, self.Ct, self.N)), 16),
                                    dtype=torch.uint8, device=dev)
            self.cams = torch.empty(self.Ct, 25, dtype=torch.float32, device=dev)
            self.totals = torch.zeros(4, dtype=torch.int64, device=dev)
            self.status = torch.zeros(4, dtype=torch.int32, device=dev)
            self.render_colors = torch.empty(self.Cl, self.H, self.W, self.D, dtype=torch.float32, device=dev)
            self.render_alphas = torch.empty(self.Cl, self.H, self.W, 1, dtype=torch.float32, device=dev)
            self.last_ids = torch.empty(self.Cl, self.H, self.W, dtype=torch.int32, device=dev)
            self.offsets = torch.empty(self.Cl, self.tile_h, self.tile_w, dtype=torch.int32, device=dev)
        self.row_capacity = row_capacity
        self.max_isects = max_isects
        self._alloc_rows = 0

    # ---- sizing -------------------------------------------------------------------------------------------------------------
    def _alloc_render(self, rows: int, max_isects: int) -> None:
        dev = self.device
        self._alloc_rows, self.max_isects = rows, max_isects
        with torch.cuda.device(dev):
            self.tiles_per_gauss = torch.empty(rows, dtype=torch.int32, device=dev)
            self.footprints = torch.empty((rows, 4), dtype=torch.int32, device=dev)  # 16 B per row, see rs_isect_footprints
            self.block_sums = torch.empty(self.lib.rs_isect_num_blocks(rows) + 2, dtype=torch.int32, device=dev)
            self.records = torch.empty((rows, 8), dtype=torch.float32, device=dev)
            self.flatten_ids = torch.empty(max_isects, dtype=torch.int32, device=dev)
            nbytes = int(self.lib.rs_isect_sorted_workspace_bytes(rows, max_isects))
            self.bin_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.tile_counter = torch.zeros(1, dtype=torch.int32, device=dev)

    def _size_once(self, viewmats: Tensor, Ks: Tensor, rigid) -> None:
        """First frame only: one ordinary (synchronising) sharded render tells how many rows and intersections this scene
        produces; the persistent arrays get 25 % head room on top of the largest rank's numbers."""
        from .rendering import rasterization

        kw = {}
        if rigid is not None:
            kw = dict(cluster_ids=self.cluster_ids, body_quats=rigid[0], body_trans=rigid[1], body_centers=self.body_centers)
        with torch.no_grad():
            _, _, meta = rasterization(self.means, self.quats, self.scales, self.opacities, self.colors, viewmats, Ks, self.W,
                                       self.H, packed=True, distributed=True, near_plane=self.near_plane,
                                       far_plane=self.far_plane, radius_clip=self.radius_clip, eps2d=self.eps2d, **kw)
        need = torch.tensor([meta["gaussian_ids"].numel(), meta["flatten_ids"].numel()], dtype=torch.int64, device=self.device)
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        rows, isects = (int(v) for v in need.tolist())
        rows = max(int(rows * 1.25) + 4096, int(self.row_capacity or 0))
        isects = max(int(isects * 1.25) + 65536, int(self.max_isects or 0))
        if self.peer.buffers is None or self.peer.buffers.capacity < rows:
            self.peer._regrow(rows)
        self._alloc_render(self.peer.buffers.capacity, isects)

    # ---- one frame ----------------------------------------------------------------------------------------------------------
    def render(self, viewmats: Tensor, Ks: Tensor, body_quats: Optional[Tensor] = None, body_trans: Optional[Tensor] = None):
        """Enqueue one frame on the current stream; returns views of the renderer-owned (render_colors [Cl,H,W,D],
        render_alphas [Cl,H,W,1]).  Nothing is read back; call check() when the sizes must be verified."""
        lib, dev = self.lib, self.device
        rigid = None
        if self.cluster_ids is not None:
            if body_quats is None or body_trans is None:
                raise RuntimeError("ShardedFrameRenderer.render: body_quats and body_trans are required with cluster_ids")
            rigid = (body_quats.contiguous(), body_trans.contiguous())
        if self._alloc_rows == 0 or self.peer.buffers is None or self.peer.buffers.capacity != self._alloc_rows:
            self._size_once(viewmats, Ks, rigid)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            # 1. cameras of every rank (the one NCCL call of the frame: 25 floats per camera, no host involvement)
            local = torch.cat([viewmats.reshape(self.Cl, 16), Ks.reshape(self.Cl, 9)], dim=1).contiguous()
            if self.world > 1:
                dist.all_gather_into_tensor(self.cams, local, group=self.group)
            else:
                self.cams.copy_(local)
            vm_all = self.cams[:, :16].contiguous()
            Ks_all = self.cams[:, 16:].contiguous()
            # 2. packed projection of MY Gaussians to ALL cameras (device-side nnz / indptr)
            pa = _lib.rs_project_packed_fwd_args()
            a = pa.proj
            a.B, a.C, a.N = 1, self.Ct, self.N
            a.image_width, a.image_height, a.camera_model = self.W, self.H, 0
            a.eps2d, a.near_plane, a.far_plane, a.radius_clip = self.eps2d, self.near_plane, self.far_plane, self.radius_clip
            a.means, a.quats, a.scales = self.means.data_ptr(), self.quats.data_ptr(), self.scales.data_ptr()
            a.opacities, a.viewmats, a.Ks = self.opacities.data_ptr(), vm_all.data_ptr(), Ks_all.data_ptr()
            if rigid is not None:
                from ._C import RigidPoses

                RigidPoses(self.cluster_ids, rigid[0], rigid[1], self.body_centers).fill(a.rigid)
            a.radii, a.means2d = self.p_radii.data_ptr(), self.p_means2d.data_ptr()
            a.depths, a.conics = self.p_depths.data_ptr(), self.p_conics.data_ptr()
            pa.capacity = self.cap_p
            pa.indptr = self.p_indptr.data_ptr()
            pa.batch_ids, pa.camera_ids, pa.gaussian_ids = (self.p_ids[i].data_ptr() for i in range(3))
            pa.nnz, pa.workspace = self.p_nnz.data_ptr(), self.p_ws.data_ptr()
            _lib.check(lib.rs_project_packed_fwd(ctypes.byref(pa), stream))
            # 3. exchange over peer memory: push, wait, blank the unused tail of the receive arrays
            peer, buf = self.peer, self.peer.buffers
            peer.epoch += 1
            x = _lib.rs_exchange_args()
            x.world, x.rank, x.cameras_per_rank, x.channels = self.world, self.rank, self.Cl, self.D
            x.capacity, x.epoch = buf.capacity, peer.epoch
            x.colors_per_row, x.opacities_per_row, x.timeout_ms = 0, 0, int(peer.timeout_ms)
            x.peer_base = buf.table.data_ptr()
            x.indptr, x.camera_ids, x.gaussian_ids = self.p_indptr.data_ptr(), self.p_ids[1].data_ptr(), self.p_ids[2].data_ptr()
            x.radii, x.means2d, x.depths, x.conics = (t.data_ptr() for t in (self.p_radii, self.p_means2d, self.p_depths,
                                                                            self.p_conics))
            x.compensations = None
            x.opacities, x.colors = self.opacities.data_ptr(), self.colors.data_ptr()
            x.gaussian_base, x.nnz = self.gaussian_base, self.cap_p
            _lib.check(lib.rs_exchange_push(ctypes.byref(x), stream))
            _lib.check(lib.rs_exchange_wait(ctypes.byref(x), self.totals.data_ptr(), stream))
            _lib.check(lib.rs_exchange_seal(ctypes.byref(x), self.totals.data_ptr(), stream))
            peer._last_args = x
            col = lambda i: buf.own + buf.offsets[i]
            rows = buf.capacity
            # 4. tile binning over the row capacity (rows beyond the received ones have zero radii)
            ia = _lib.rs_isect_args()
            ia.n_elems, ia.N, ia.I = rows, 0, self.Cl
            ia.tile_size, ia.tile_width, ia.tile_height = 16, self.tile_w, self.tile_h
            ia.means2d, ia.radii, ia.depths, ia.image_ids = col(0), col(5), col(1), col(6)
            ia.tiles_per_gauss, ia.block_sums = self.tiles_per_gauss.data_ptr(), self.block_sums.data_ptr()
            ia.n_isects, ia.overflow = self.status.data_ptr(), self.status.data_ptr() + 4
            ia.isect_ids, ia.flatten_ids, ia.capacity = None, self.flatten_ids.data_ptr(), self.max_isects
            # counts + one 16-byte footprint per row: the emission then gathers one record per row instead of three
            _lib.check(lib.rs_isect_footprints(ctypes.byref(ia), col(2) if self.tight_tiles else None,
                                               col(3) if self.tight_tiles else None, self.footprints.data_ptr(), stream))
            sa = _lib.rs_isect_sorted_args()
            ctypes.memmove(ctypes.byref(sa.isect), ctypes.byref(ia), ctypes.sizeof(ia))
            sa.tile_footprints = self.footprints.data_ptr()
            sa.tile_offsets = self.offsets.data_ptr()
            sa.workspace, sa.workspace_bytes = self.bin_ws.data_ptr(), self.bin_ws.numel()
            _lib.check(lib.rs_isect_sorted(ctypes.byref(sa), stream))
            # 5. compositing of my cameras
            self.tile_counter.zero_()
            r = _lib.rs_raster_fwd_args()
            r.I, r.N, r.channels = self.Cl, 0, self.D
            r.image_width, r.image_height, r.tile_size = self.W, self.H, 16
            r.tile_width, r.tile_height = self.tile_w, self.tile_h
            r.n_isects, r.n_isects_dev = self.max_isects, self.status.data_ptr()
            r.means2d, r.conics, r.colors, r.opacities = col(0), col(2), col(4), col(3)
            r.tile_offsets, r.flatten_ids = self.offsets.data_ptr(), self.flatten_ids.data_ptr()
            r.render_colors, r.render_alphas = self.render_colors.data_ptr(), self.render_alphas.data_ptr()
            r.last_ids = self.last_ids.data_ptr()
            r.records, r.records_ready, r.n_rows = self.records.data_ptr(), 0, rows
            r.tile_counter = self.tile_counter.data_ptr()
            _lib.check(lib.rs_raster_fwd(ctypes.byref(r), stream))
        return self.render_colors, self.render_alphas

    def check(self) -> Dict[str, int]:
        """The one host read: sizes and flags of the LAST frame.  Raises when the exchange failed; returns a dict with
        `regrow` = True when a capacity was exceeded (the arrays are then re-sized collectively: render the frame again)."""
        got, worst, err, behind = self.totals.tolist()
        n_isects, overflow = self.status[:2].tolist()
        if err == 1:
            raise RuntimeError(f"ShardedFrameRenderer rank {self.rank}: a peer did not arrive within the spin limit "
                               f"(flags behind {behind:#x})")
        flag = torch.tensor([1 if (err == 2 or overflow) else 0, worst, n_isects], dtype=torch.int64, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        regrow, worst_all, isects_all = (int(v) for v in flag.tolist())
        if regrow:
            rows = max(int(worst_all * 1.25) + 4096, self.peer.buffers.capacity)
            if rows > self.peer.buffers.capacity:
                self.peer._regrow(rows)
            self._alloc_render(self.peer.buffers.capacity, max(int(isects_all * 1.25) + 65536, self.max_isects))
        return dict(rows=got, rows_needed=worst_all, n_isects=n_isects, regrow=bool(regrow))
