"""Rigid-body pose helpers: the data contract between clustering / physics and the renderer.

Reference semantics: main.py:183-228 `apply_transform(splats, translation, rotation_quat)` moves ONE body per call by
cloning every splat tensor; `cluster_groups.npz` (examples/load_identity_encodings.py:478-491, 566-568) stores, per
object id, the list of Gaussian indices of that cluster (main.py:280-297 consumes key "1").  Here the same information
is a dense per-Gaussian `cluster_ids[N]` (int32, -1 = static background) plus per-body pose tables consumed inside the
projection kernel.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from ._C import RigidPoses


def cluster_ids_from_groups(groups: Mapping[str, Sequence[int]], num_gaussians: int,
                            device: Optional[torch.device] = None) -> Tuple[Tensor, Dict[int, str]]:
    """Dense int32 cluster ids from a `cluster_groups.npz`-style mapping {object id -> Gaussian indices}.

    The "background" key (and every Gaussian not listed) maps to -1.  Returns (cluster_ids [N], {body index -> key})."""
    ids = np.full(num_gaussians, -1, dtype=np.int32)
    names: Dict[int, str] = {}
    k = 0
    for key in sorted(groups.keys(), key=lambda s: (not str(s).lstrip("-").isdigit(), str(s))):
        if str(key) == "background":
            continue
        idx = np.asarray(groups[key], dtype=np.int64).reshape(-1)
        ids[idx] = k
        names[k] = str(key)
        k += 1
    t = torch.from_numpy(ids)
    return (t.to(device) if device is not None else t), names


def body_centers(means: Tensor, cluster_ids: Tensor, K: int) -> Tensor:
    """Per-body mean of the Gaussian centres: the pivot apply_transform() rotates about (main.py:210)."""
    valid = cluster_ids >= 0
    idx = cluster_ids[valid].long()
    sums = torch.zeros(K, 3, dtype=means.dtype, device=means.device).index_add_(0, idx, means[valid])
    cnt = torch.zeros(K, dtype=means.dtype, device=means.device).index_add_(
        0, idx, torch.ones_like(idx, dtype=means.dtype))
    return sums / cnt.clamp_min(1.0)[:, None]


def make_rigid(cluster_ids: Optional[Tensor], body_quats: Optional[Tensor], body_trans: Optional[Tensor],
               body_centers: Optional[Tensor] = None) -> Optional[RigidPoses]:
    """Build the RigidPoses argument of the projection (None when no cluster ids are given)."""
    if cluster_ids is None:
        assert body_quats is None and body_trans is None, "body poses given without cluster_ids"
        return None
    assert body_quats is not None and body_trans is not None, "cluster_ids requires body_quats and body_trans"
    dev = cluster_ids.device
    return RigidPoses(
        cluster_ids.to(torch.int32).contiguous(),
        body_quats.to(device=dev, dtype=torch.float32).contiguous(),
        body_trans.to(device=dev, dtype=torch.float32).contiguous(),
        None if body_centers is None else body_centers.to(device=dev, dtype=torch.float32).contiguous(),
    )
