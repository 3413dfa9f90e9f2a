"""Plain-torch helpers that restate small pieces of host-side glue (NOT a fallback for the CUDA path).

`apply_rigid_torch` is the multi-body generalisation of main.py:183-228 in differentiable torch ops; the product uses
it only to obtain the moved means for SH view directions, and the tests use it to build "reference pipeline" inputs
(apply_transform per body, then the un-fused projection).
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


# ---- spherical harmonics in torch: the checker of csrc/sh.cu (tests) and the CPU-side host logic test --------------
_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)
_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
       -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def sh_bases(degree: int, dirs: Tensor) -> Tensor:
    """Real SH basis values [..., (degree+1)^2] at unit directions `dirs` [..., 3]."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, _C0)]
    if degree >= 1:
        out += [-_C1 * y, _C1 * z, -_C1 * x]
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
    if degree >= 3:
        out += [
            _C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
            _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
            _C3[6] * x * (xx - 3 * yy),
        ]
    if degree >= 4:
        out += [
            _C4[0] * xy * (xx - yy), _C4[1] * yz * (3 * xx - yy), _C4[2] * xy * (7 * zz - 1),
            _C4[3] * yz * (7 * zz - 3), _C4[4] * (zz * (35 * zz - 30) + 3), _C4[5] * xz * (7 * zz - 3),
            _C4[6] * (xx - yy) * (7 * zz - 1), _C4[7] * xz * (xx - 3 * yy),
            _C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)),
        ]
    return torch.stack(out, dim=-1)


def spherical_harmonics_torch(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    assert 0 <= degrees_to_use <= 4, degrees_to_use
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    batch_dims = dirs.shape[:-1]
    assert dirs.shape == batch_dims + (3,), dirs.shape
    assert coeffs.dim() == len(batch_dims) + 2 and coeffs.shape[:-2] == batch_dims and coeffs.shape[-1] == 3, coeffs.shape
    nb = (degrees_to_use + 1) ** 2
    bases = sh_bases(degrees_to_use, F.normalize(dirs, p=2, dim=-1))  # [..., nb]
    colors = (bases[..., None] * coeffs[..., :nb, :]).sum(dim=-2)
    if masks is not None:
        assert masks.shape == batch_dims, masks.shape
        colors = torch.where(masks[..., None], colors, torch.zeros_like(colors))
    return colors


def cgc_loss_and_grad_torch(feature_map: Tensor, instance_mask: Tensor, min_cluster_size: int = 30, eps: float = 1e-6):
    """Plain-torch restatement of cgc_contrastive_clustering_loss (examples/utils.py:828-904) AND of its gradient with respect
    to the feature map, written out by hand: the formulas the CUDA kernels implement (csrc/cgc.cu).  Used by the tests only
    (pinned against the reference function's own autograd by tests/golden/cgc_loss.npz).  Returns (loss, dL/dfeature_map)."""
    H, W, D = feature_map.shape
    x = feature_map.reshape(-1, D)
    m = instance_mask.reshape(-1)
    zero = (torch.zeros((), dtype=x.dtype, device=x.device), torch.zeros_like(feature_map))
    nrm = x.norm(dim=-1).clamp_min(1e-12)
    f = x / nrm[:, None]
    fg = torch.unique(m)
    fg = fg[fg != 0]
    if fg.numel() < 2:
        return zero
    K = fg.numel()
    cl = torch.where(m != 0, torch.searchsorted(fg, m), torch.full_like(m, -1))
    S = torch.zeros(K, D, dtype=x.dtype, device=x.device).index_add_(0, cl[cl >= 0], f[cl >= 0])
    n = torch.bincount(cl[cl >= 0], minlength=K).to(x.dtype)
    valid = n >= min_cluster_size
    if int(valid.sum()) < 2:
        return zero
    Kv = int(valid.sum())
    vmap = torch.full((K,), -1, dtype=torch.long, device=x.device)
    vmap[valid] = torch.arange(Kv, device=x.device)
    t = vmap[cl]  # the reference's quirk: cl == -1 (background) indexes the LAST cluster
    member = torch.where(cl >= 0, vmap[cl.clamp_min(0)], torch.full_like(cl, -1))
    nv = n[valid]
    mk = S[valid] / nv[:, None]
    mn = mk.norm(dim=-1).clamp_min(1e-12)
    c = mk / mn[:, None]
    act = t >= 0
    A = int(act.sum())
    fa, ta = f[act], t[act]
    rows = torch.arange(A, device=x.device)
    s = fa @ c.T
    na = torch.bincount(ta, minlength=Kv).to(x.dtype)
    phi_raw = torch.zeros(Kv, dtype=x.dtype, device=x.device).index_add_(0, ta, s[rows, ta]) / na.clamp_min(1)
    phi = phi_raw.clamp_min(eps)
    tau = phi[ta]
    logits = s / tau[:, None]
    loss = (torch.logsumexp(logits, 1) - logits[rows, ta]).mean()
    onehot = F.one_hot(ta, Kv).to(x.dtype)
    g = (torch.softmax(logits, 1) - onehot) / tau[:, None] / A
    dtau = -(g * s).sum(1) / tau
    h = torch.zeros(Kv, dtype=x.dtype, device=x.device).index_add_(0, ta, dtau) * (phi_raw > eps).to(x.dtype)
    G = g + onehot * (h / na.clamp_min(1))[ta][:, None]
    u = G.T @ fa
    v = (u - (u * c).sum(1, keepdim=True) * c) / (mn * nv)[:, None]
    df = torch.zeros_like(f)
    df[act] = G @ c
    mem = member >= 0
    df[mem] += v[member[mem]]
    dx = (df - (df * f).sum(1, keepdim=True) * f) / nrm[:, None]
    return loss, dx.reshape(H, W, D)
