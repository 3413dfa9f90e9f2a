"""Identity-feature training glue: the contrastive clustering loss applied to the rendered feature map.

`cgc_contrastive_clustering_loss` has the signature and the results of the reference's function of the same name
(examples/utils.py:828-904, called at examples/simple_trainer.py:960-965 on `feat_map[0]`, the [H, W, 16] output of the c3
compositing pass) -- including its treatment of background pixels, which index the cluster table with -1 and therefore
count as pixels of the LAST foreground cluster in the loss (utils.py:878-883) while staying out of its centroid.  The
per-pixel work runs in hand-written kernels behind the C ABI (rs_cgc_fwd / rs_cgc_bwd, csrc/cgc.cu): three passes over the
feature map forward, one more backward, instead of ~25 torch kernels each way.  Only the integer mask is touched by torch
ops here (`cluster_tables`, cacheable per training image).  There is no CPU path.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import _lib

MAX_DIM, MAX_CLUSTERS = 32, 128  # RS_CGC_MAX_DIM / RS_CGC_MAX_CLUSTERS of include/rigidsplat.h


@dataclass
class ClusterTables:
    """What the loss needs from the instance mask alone."""

    target: Tensor  # int32 [P]: cluster whose temperature / logit row the pixel uses, -1 = inactive
    member: Tensor  # int32 [P]: cluster whose centroid the pixel belongs to, -1 = none
    n_member: Tensor  # float32 [K]
    n_active: Tensor  # float32 [K]
    n_active_pixels: int
    K: int


def cluster_tables(instance_mask: Tensor, min_cluster_size: int = 30) -> Optional[ClusterTables]:
    """Per-pixel cluster indices among the valid foreground clusters (ids != 0 with >= min_cluster_size pixels, in
    ascending id order: utils.py:845-871).  None when fewer than two clusters qualify (the reference then returns 0)."""
    m = instance_mask.reshape(-1)
    uniq = torch.unique(m)
    fg = uniq[uniq != 0]
    if fg.numel() < 2:
        return None
    is_fg = m != 0
    cl = torch.where(is_fg, torch.searchsorted(fg, m), torch.full_like(m, -1))
    counts = torch.bincount(cl[is_fg], minlength=fg.numel())
    valid = counts >= min_cluster_size
    K = int(valid.sum())
    if K < 2:
        return None
    vmap = torch.full((fg.numel(),), -1, dtype=torch.long, device=m.device)
    vmap[valid] = torch.arange(K, device=m.device)
    member = torch.where(is_fg, vmap[cl.clamp_min(0)], torch.full_like(cl, -1))
    # the reference looks background pixels up with index -1, i.e. in the LAST foreground cluster (utils.py:878-883)
    target = torch.where(is_fg, member, vmap[-1].expand_as(member))
    active = target >= 0
    return ClusterTables(
        target=target.to(torch.int32).contiguous(), member=member.to(torch.int32).contiguous(),
        n_member=counts[valid].to(torch.float32).contiguous(),
        n_active=torch.bincount(target[active], minlength=K).to(torch.float32).contiguous(),
        n_active_pixels=int(active.sum()), K=K)


def _args(features: Tensor, tb: ClusterTables, eps: float, ws: Tensor):
    a = _lib.rs_cgc_args()
    a.P, a.A, a.D, a.K = features.shape[0], tb.n_active_pixels, features.shape[1], tb.K
    a.eps = eps
    a.features = features.data_ptr()
    a.target, a.member = tb.target.data_ptr(), tb.member.data_ptr()
    a.n_member, a.n_active = tb.n_member.data_ptr(), tb.n_active.data_ptr()
    a.ws = ws.data_ptr()
    return a


class _CgcLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features: Tensor, tb: ClusterTables, eps: float) -> Tensor:
        lib = _lib.load()
        features = features.contiguous()
        with torch.cuda.device(features.device):
            ws = torch.empty(int(lib.rs_cgc_workspace_floats(tb.K, features.shape[1])), dtype=torch.float32,
                             device=features.device)
            a = _args(features, tb, eps, ws)
            a.accumulate_grad = 1 if ctx.needs_input_grad[0] else 0
            _lib.check(lib.rs_cgc_fwd(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(features, ws)
        ctx.tb, ctx.eps = tb, eps
        return ws[-1].clone()

    @staticmethod
    def backward(ctx, grad_loss: Tensor):
        features, ws = ctx.saved_tensors
        lib = _lib.load()
        with torch.cuda.device(features.device):
            v = torch.empty_like(features)
            g = grad_loss.to(torch.float32).reshape(1).contiguous()
            a = _args(features, ctx.tb, ctx.eps, ws)
            a.grad_loss, a.v_features = g.data_ptr(), v.data_ptr()
            _lib.check(lib.rs_cgc_bwd(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
        return v, None, None


def cgc_contrastive_clustering_loss(feature_map: Tensor, instance_mask: Tensor, min_cluster_size: int = 30,
                                    eps: float = 1e-6, tables: Optional[ClusterTables] = None) -> Tensor:
    """Contrastive clustering loss of a rendered feature map [H, W, D] against an instance-id mask [H, W] (0 = background).
    `tables` may carry a cached `cluster_tables(instance_mask, min_cluster_size)`."""
    H, W, D = feature_map.shape
    if not feature_map.is_cuda or feature_map.dtype != torch.float32:
        raise RuntimeError("cgc_contrastive_clustering_loss: feature_map must be a float32 CUDA tensor")
    if D > MAX_DIM:
        raise RuntimeError(f"cgc_contrastive_clustering_loss: {D} feature channels exceed the kernel limit of {MAX_DIM}")
    tb = tables if tables is not None else cluster_tables(instance_mask.to(feature_map.device), min_cluster_size)
    if tb is None:  # fewer than two usable clusters (utils.py:849-850, 866-867)
        return torch.tensor(0.0, device=feature_map.device, requires_grad=True)
    if tb.K > MAX_CLUSTERS:
        raise RuntimeError(f"cgc_contrastive_clustering_loss: {tb.K} clusters exceed the kernel limit of {MAX_CLUSTERS}")
    return _CgcLoss.apply(feature_map.reshape(H * W, D), tb, float(eps))


# ---------------------------------------------------------------------------------------------------------------------
# Segmentation head (examples/simple_trainer.py:442-446, 946-947), fused: rs_seghead_fwd / rs_seghead_bwd
# ---------------------------------------------------------------------------------------------------------------------
def _seg_args(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor):
    a = _lib.rs_seghead_args()
    a.N, a.D, a.H = x.shape[0], x.shape[1], w1.shape[0]
    a.x, a.w1, a.b1, a.w2, a.b2 = (t.data_ptr() for t in (x, w1, b1, w2, b2))
    return a


class _SegHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        lib = _lib.load()
        x, w1, b1, w2, b2 = (t.contiguous() for t in (x, w1, b1, w2, b2))
        with torch.cuda.device(x.device):
            y = torch.empty_like(x)
            a = _seg_args(x, w1, b1, w2, b2)
            a.y = y.data_ptr()
            _lib.check(lib.rs_seghead_fwd(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(x, w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, v_y):
        x, w1, b1, w2, b2 = ctx.saved_tensors
        lib = _lib.load()
        v_y = v_y.contiguous()
        with torch.cuda.device(x.device):
            v_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            grads = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
            a = _seg_args(x, w1, b1, w2, b2)
            a.v_y = v_y.data_ptr()
            a.v_x = v_x.data_ptr() if v_x is not None else None
            a.v_w1, a.v_b1, a.v_w2, a.v_b2 = (g.data_ptr() for g in grads)
            _lib.check(lib.rs_seghead_bwd(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
        return (v_x,) + tuple(g if need else None for g, need in zip(grads, ctx.needs_input_grad[1:]))


def segmentation_head_forward(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    """y = W2 relu(W1 x + b1) + b2 for every row of x [N, D] (torch.nn.Linear weight layout), differentiable; the [N, H]
    hidden layer never reaches HBM.  Backward supports D = 16, H = 64 (the reference's head)."""
    for t, n in ((x, "x"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError(f"segmentation_head_forward: {n} must be a float32 CUDA tensor")
    D, H = x.shape[-1], w1.shape[0]
    if tuple(w1.shape) != (H, D) or tuple(b1.shape) != (H,) or tuple(w2.shape) != (D, H) or tuple(b2.shape) != (D,):
        raise RuntimeError("segmentation_head_forward: parameter shapes must be [H,D], [H], [D,H], [D]")
    return _SegHead.apply(x.reshape(-1, D), w1, b1, w2, b2).reshape(x.shape)


class SegmentationHead(torch.nn.Sequential):
    """Drop-in for the reference's `torch.nn.Sequential(Linear(identity_dim, 64), ReLU(), Linear(64, identity_dim))`
    (examples/simple_trainer.py:442-446): same sub-modules, parameter names (`0.weight`, `0.bias`, `2.weight`, `2.bias`),
    initialisation and state_dict -- an optimizer or a checkpoint of the reference works unchanged -- but forward / backward
    run as one fused kernel each (rs_seghead_fwd / rs_seghead_bwd) instead of two / four library GEMMs around a
    materialised [N, 64] hidden tensor."""

    def __init__(self, identity_dim: int = 16, hidden: int = 64):
        super().__init__(torch.nn.Linear(identity_dim, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, identity_dim))

    def forward(self, x: Tensor) -> Tensor:
        return segmentation_head_forward(x, self[0].weight, self[0].bias, self[2].weight, self[2].bias)
