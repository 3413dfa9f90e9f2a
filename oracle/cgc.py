"""CPU/GPU-agnostic torch restatement of the reference's contrastive clustering loss AND of its gradient, written out by
hand: the formulas the CUDA kernels of 3dgs_rigidbody_b200/csrc/cgc.cu implement.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/, never by the product package.
Follows `cgc_contrastive_clustering_loss` (/root/reference/examples/utils.py:828-904), including its background quirk
(utils.py:878-883: background pixels index the cluster table with -1, i.e. the LAST foreground cluster).  Pinned against the
reference function's own autograd by tests/golden/cgc_loss.npz (tests/golden/make_golden_cgc.py):
tests/test_identity_loss.py::test_torch_restatement_matches_reference_autograd.
"""
import torch
import torch.nn.functional as F
from torch import Tensor


def cgc_loss_and_grad_torch(feature_map: Tensor, instance_mask: Tensor, min_cluster_size: int = 30, eps: float = 1e-6):
    """-> (loss, dL/dfeature_map) for a feature map [H, W, D] and an instance-id mask [H, W] (0 = background)."""
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
