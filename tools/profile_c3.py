#!/usr/bin/env python
"""One c3 training step at a time (1 M Gaussians, 16 identity-feature channels, forward + backward at 1080p through
rasterization() with the rigid poses fused in, then the contrastive clustering loss and the fused segmentation head): the
target of the `ncu --set full` captures of the backward kernels.

    ncu --set full --clock-control none --import-source on -k regex:"rs_raster_bwd|rs_project_bwd|rs_cgc|rs_seghead" -c 12 \
        -o gpurun_out/prof_c3 python tools/profile_c3.py --steps 2"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rs = importlib.import_module("3dgs_rigidbody_b200")
    dev = "cuda:0"
    W, H, D = bench.WIDTH, bench.HEIGHT, 16
    sc = bench.make_domino_scene(device=dev)
    g = torch.Generator(device=dev).manual_seed(42)
    feats = torch.randn(sc["means"].shape[0], D, device=dev, generator=g).requires_grad_()
    w = torch.rand(1, H, W, D, device=dev, generator=g)
    leaves = [sc[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities")]
    mask = torch.zeros(H, W, dtype=torch.long, device=dev)
    for a in range(4):
        for b in range(6):
            mask[20 + a * 260:20 + a * 260 + 240, 20 + b * 315:20 + b * 315 + 290] = 1 + a * 6 + b
    tables = rs.cluster_tables(mask, 30)
    torch.manual_seed(0)
    head = rs.SegmentationHead(16, 64).to(dev)
    for step in range(args.steps):
        bq, bt = bench.domino_poses(bench.N_BODIES, frame=60 + step, device=dev, centers=sc["body_centers"])
        for t in leaves + [feats] + list(head.parameters()):
            t.grad = None
        processed = head(feats)  # examples/simple_trainer.py:946-947
        img, _, _ = rs.rasterization(*leaves, processed, sc["viewmats"], sc["Ks"], W, H, packed=False, cluster_ids=sc["cluster_ids"],
                                     body_quats=bq, body_trans=bt, body_centers=sc["body_centers"])
        loss = (img * w).sum() + rs.cgc_contrastive_clustering_loss(img[0], mask, tables=tables)
        loss.backward()
        torch.cuda.synchronize()
    print("ok", float(loss))


if __name__ == "__main__":
    main()
