"""Differentiable torch form of the rigid transform (NOT a fallback for the CUDA path: it runs on whatever device its
tensors live on and the product only ever hands it CUDA tensors).

`apply_rigid_torch` is the multi-body generalisation of main.py:183-228; `rasterization()` uses it for one thing only --
the moved means that give the SH view directions on the operator path (the frame path evaluates SH inside the projection
kernel) -- and the tests use it to build "reference pipeline" inputs (apply_transform per body, then the un-fused
projection).  Torch restatements that exist purely as checkers are not kept in this package.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


def quat_multiply(q1: Tensor, q2: Tensor) -> Tensor:
    """Hamilton product, wxyz (main.py:173-181)."""
    w1, x1, y1, z1 = q1.unbind(-1)
    w2, x2, y2, z2 = q2.unbind(-1)
    return torch.stack(
        (
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ),
        -1,
    )


def normalized_quat_to_rotmat(quat: Tensor) -> Tensor:
    """gsplat/utils.py:109-134."""
    w, x, y, z = torch.unbind(quat, dim=-1)
    mat = torch.stack(
        [
            1 - 2 * (y**2 + z**2), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x**2 + z**2), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x**2 + y**2),
        ],
        dim=-1,
    )
    return mat.reshape(quat.shape[:-1] + (3, 3))


def apply_rigid_torch(means: Tensor, quats: Optional[Tensor], rigid) -> Tuple[Tensor, Optional[Tensor]]:
    """mean' = R_k (mean - c_k) + c_k + t_k, quat' = q_k (x) quat for every Gaussian with cluster id k >= 0."""
    ids = rigid.cluster_ids.long()
    moving = ids >= 0
    k = ids.clamp_min(0)
    q = rigid.body_quats / torch.linalg.norm(rigid.body_quats, dim=-1, keepdim=True)
    R = normalized_quat_to_rotmat(q)[k]  # [N,3,3]
    c = rigid.body_centers[k] if rigid.body_centers is not None else torch.zeros_like(means[..., :3])
    t = rigid.body_trans[k]
    moved = torch.einsum("nij,...nj->...ni", R, means - c) + c + t
    means_out = torch.where(moving[:, None], moved, means)
    quats_out = None
    if quats is not None:
        quats_out = torch.where(moving[:, None], quat_multiply(q[k].expand_as(quats), quats), quats)
    return means_out, quats_out
