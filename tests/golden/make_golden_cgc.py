"""Generates tests/golden/cgc_loss.npz from the REFERENCE's own function: cgc_contrastive_clustering_loss
(/root/reference/examples/utils.py:828-904, extracted by AST because the module imports packages that are absent here)
and its autograd gradient with respect to the feature map, float32 on CPU.  Run in the build container only:
    python tests/golden/make_golden_cgc.py
Cases: (a) last foreground cluster valid -> background pixels join it (the index -1 quirk); (b) last cluster too small ->
background inactive; (c) 8 channels, many clusters; (d) fewer than two valid clusters -> loss 0."""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/examples/utils.py").read()
fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "cgc_contrastive_clustering_loss"][0]
ns = {"torch": torch, "F": F}
exec(compile(ast.Module([fn], []), "reference_utils", "exec"), ns)
ref = ns["cgc_contrastive_clustering_loss"]


def boxes_mask(H, W, boxes):
    m = torch.zeros(H, W, dtype=torch.long)
    for (y0, y1, x0, x1, v) in boxes:
        m[y0:y1, x0:x1] = v
    return m


def case(seed, H, W, D, boxes, min_size=30):
    torch.manual_seed(seed)
    mask = boxes_mask(H, W, boxes)
    centers = 2.0 * torch.randn(64, D)
    x = (torch.randn(H, W, D) + centers[mask.clamp_max(63)]).requires_grad_()
    loss = ref(x, mask, min_cluster_size=min_size)
    if loss.grad_fn is not None:
        loss.backward()
        grad = x.grad
    else:
        grad = torch.zeros_like(x)
    return dict(x=x.detach().numpy(), mask=mask.numpy().astype(np.int64), loss=np.float32(loss.item()), grad=grad.numpy(),
                min_size=np.int64(min_size))


cases = {
    "a": case(0, 96, 128, 16, [(5, 40, 5, 60, 3), (50, 90, 10, 100, 7), (2, 6, 100, 104, 9), (60, 92, 104, 124, 12)]),
    "b": case(1, 96, 128, 16, [(5, 40, 5, 60, 3), (50, 90, 10, 100, 7), (2, 6, 100, 104, 9), (60, 63, 104, 107, 12)]),
    "c": case(2, 64, 80, 8, [(4 * i, 4 * i + 8, 6 * j, 6 * j + 10, 1 + i * 5 + j) for i in range(6) for j in range(5)] +
              [(50, 64, 0, 80, 40)]),
    "d": case(3, 32, 32, 16, [(0, 16, 0, 16, 5), (20, 22, 20, 22, 6)]),
}
out = {}
for name, c in cases.items():
    for k, v in c.items():
        out[f"{name}_{k}"] = v
    print(name, "loss", float(c["loss"]), "max |grad|", float(np.abs(c["grad"]).max()))
np.savez_compressed(os.path.join(HERE, "cgc_loss.npz"), **out)
