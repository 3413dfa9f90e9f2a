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
    """Per-body mean of the Gaussian centres: the pivot apply_transform() rotates about (main.py:210).

    Evaluated exactly as the reference does for each body -- `means[body].mean(dim=0)` on the body's Gaussians in index
    order -- so the pivots (and therefore the moved means) are bit-identical to the per-body apply_transform() chain; it is
    a per-scene precomputation (K small reductions), not a per-frame step."""
    out = torch.zeros(K, 3, dtype=means.dtype, device=means.device)
    for k in range(K):
        idx = torch.nonzero(cluster_ids == k).squeeze(-1)
        if idx.numel() > 0:
            out[k] = means[idx].mean(dim=0)
    return out


_mean_centers = body_centers  # make_rigid() has a parameter of the same name


def make_rigid(cluster_ids: Optional[Tensor], body_quats: Optional[Tensor], body_trans: Optional[Tensor],
               body_centers: Optional[Tensor] = None, means: Optional[Tensor] = None) -> Optional[RigidPoses]:
    """Build the RigidPoses argument of the projection (None when no cluster ids are given).

    Without `body_centers` the pivot of every body is the mean of its Gaussian centres, as apply_transform() does
    (main.py:210), computed here from `means` ([N,3]; one extra pass over the means per call -- an animation loop should
    compute `body_centers(...)` once and pass it).  The poses are not differentiable: the kernels chain the gradients of
    the means / quats through the pose but produce none for the pose itself, so a pose that requires grad is rejected
    instead of silently getting no gradient."""
    if cluster_ids is None:
        assert body_quats is None and body_trans is None, "body poses given without cluster_ids"
        return None
    assert body_quats is not None and body_trans is not None, "cluster_ids requires body_quats and body_trans"
    if torch.is_grad_enabled() and (body_quats.requires_grad or body_trans.requires_grad):
        raise RuntimeError("body_quats / body_trans require grad, but the fused rigid transform has no pose gradient; "
                           "detach them (gradients still flow to means / quats through the pose)")
    dev = cluster_ids.device
    if body_centers is None and means is not None and means.dim() == 2:
        body_centers = _mean_centers(means.detach(), cluster_ids.to(means.device), body_quats.shape[0])
    return RigidPoses(
        cluster_ids.to(torch.int32).contiguous(),
        body_quats.to(device=dev, dtype=torch.float32).contiguous(),
        body_trans.to(device=dev, dtype=torch.float32).contiguous(),
        None if body_centers is None else body_centers.to(device=dev, dtype=torch.float32).contiguous(),
    )


# ---------------------------------------------------------------------------------------------------------------------
# Data contract either side of the renderer: clustering output -> cluster ids, per-body rigid parameters,
# physics poses -> pose tables.  Host-side glue (numpy / torch), nothing here is on the per-frame path.
# ---------------------------------------------------------------------------------------------------------------------
def load_cluster_groups(path: str, num_gaussians: int, device: Optional[torch.device] = None) -> Tuple[Tensor, Dict[int, str]]:
    """Reads the `cluster_groups` archive written by examples/load_identity_encodings.py:566-568 (np.savez_compressed of
    {str(object id) -> Gaussian indices, "background" -> indices}; main.py:280-297 loads it with np.load) and returns the
    dense (cluster_ids [N] int32, {body index -> object id}) form the projection kernel consumes."""
    with np.load(path, allow_pickle=False) as z:
        groups = {k: z[k] for k in z.files}
    return cluster_ids_from_groups(groups, num_gaussians, device)


def instance_mask_path(data_dir: str, image_name: str) -> str:
    """Where the per-image instance-id map of an image lives (examples/datasets/colmap.py:498-512; written by
    utils/instance_maps_to_npy.py:7-40): `<data_dir>/masks/instance_ids_npy/<stem>_instance_id.npy`, [H, W] ints, 0 = bg."""
    import os

    stem = os.path.splitext(os.path.basename(image_name))[0]
    return os.path.join(data_dir, "masks", "instance_ids_npy", f"{stem}_instance_id.npy")


def body_properties(means: Tensor, scales: Tensor, opacities: Tensor, cluster_ids: Tensor, K: int):
    """Rigid-body parameters of every cluster (README.md:12, workflow step 2): mass, centre of mass and inertia tensor about
    the centre of mass, treating each Gaussian as a point mass m_g = opacity_g * prod(scale_g) (its alpha-weighted volume
    up to a constant) plus the inertia of its own ellipsoid about its principal axes being neglected.

    Returns (mass [K], com [K,3], inertia [K,3,3]); bodies without Gaussians get zeros."""
    valid = cluster_ids >= 0
    idx = cluster_ids[valid].long()
    m = (opacities[valid] * scales[valid].prod(dim=-1)).to(torch.float64)
    x = means[valid].to(torch.float64)
    mass = torch.zeros(K, dtype=torch.float64, device=means.device).index_add_(0, idx, m)
    com = torch.zeros(K, 3, dtype=torch.float64, device=means.device).index_add_(0, idx, x * m[:, None])
    com = com / mass.clamp_min(1e-300)[:, None]
    r = x - com[idx]
    r2 = (r * r).sum(-1)
    per = m[:, None, None] * (r2[:, None, None] * torch.eye(3, dtype=torch.float64, device=means.device) - r[:, :, None] * r[:, None, :])
    inertia = torch.zeros(K, 3, 3, dtype=torch.float64, device=means.device).index_add_(0, idx, per)
    return mass.to(means.dtype), com.to(means.dtype), inertia.to(means.dtype)


class PoseStream:
    """Per-frame rigid poses of K bodies, [F, K, 7] = (quat wxyz | position of the body origin), as a physics engine
    reports them for bodies whose local origin is their rest-pose centre of mass (README.md:13-14, steps 3-4).

    frame(f) returns the (body_quats [K,4], body_trans [K,3]) tables of rasterization()/FrameRenderer for the pivot
    `centers` = rest-pose centres: x' = R (x - c) + c + t with t = p_f - c."""

    def __init__(self, poses, rest_centers):
        poses = torch.as_tensor(poses, dtype=torch.float32)
        rest_centers = torch.as_tensor(rest_centers, dtype=torch.float32)
        assert poses.dim() == 3 and poses.shape[-1] == 7, poses.shape
        assert tuple(rest_centers.shape) == (poses.shape[1], 3), rest_centers.shape
        self.quats = poses[..., :4].contiguous()
        self.trans = (poses[..., 4:] - rest_centers.to(poses.device)[None]).contiguous()
        self.centers = rest_centers

    @classmethod
    def load(cls, path: str) -> "PoseStream":
        """npz with `poses` [F,K,7] and `rest_centers` [K,3]."""
        with np.load(path, allow_pickle=False) as z:
            return cls(z["poses"], z["rest_centers"])

    def save(self, path: str) -> None:
        poses = torch.cat([self.quats, self.trans + self.centers[None]], dim=-1)
        np.savez_compressed(path, poses=poses.cpu().numpy(), rest_centers=self.centers.cpu().numpy())

    def __len__(self) -> int:
        return self.quats.shape[0]

    @property
    def num_bodies(self) -> int:
        return self.quats.shape[1]

    def to(self, device) -> "PoseStream":
        self.quats, self.trans, self.centers = self.quats.to(device), self.trans.to(device), self.centers.to(device)
        return self

    def frame(self, f: int) -> Tuple[Tensor, Tensor]:
        return self.quats[f], self.trans[f]
