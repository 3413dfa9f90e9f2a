#!/usr/bin/env python
"""One c2 frame at a time through FrameRenderer on a single stream: the target of the `ncu --set full` captures.

    ncu --set full --clock-control none --import-source on -s <12 * warm-up frames> -c 12 -o gpurun_out/prof \
        python tools/profile_frame.py --frames 4
A frame is 12 kernel launches (projection incl. the depth statistics; depth order: count, scan, scatter, sort; tile count,
scan, emission; 2 tile-key radix passes; offsets; compositing), with the tight tile lists bench.py uses;
`--packed` profiles the packed projection instead (1 launch per frame)."""
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
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--packed", action="store_true")
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--reference-tile-lists", action="store_true", help="the reference's tile lists instead of the tight ones")
    args = ap.parse_args()
    rs = importlib.import_module("3dgs_rigidbody_b200")
    dev = "cuda:0"
    sc = bench.make_domino_scene(device=dev)
    colors = sc["colors"] if args.channels == 3 else torch.rand(sc["means"].shape[0], args.channels, device=dev)
    if args.packed:
        for f in range(args.frames):
            bq, bt = bench.domino_poses(bench.N_BODIES, frame=60 + f, device=dev, centers=sc["body_centers"])
            rigid = rs._C.RigidPoses(sc["cluster_ids"], bq, bt, sc["body_centers"])
            out = rs._C.projection_ewa_3dgs_packed_fwd(sc["means"], None, sc["quats"], sc["scales"], sc["opacities"],
                                                       sc["viewmats"], sc["Ks"], bench.WIDTH, bench.HEIGHT, 0.3, 0.01, 1e10,
                                                       0.0, False, rs._C.PINHOLE, rigid)
        print("nnz", out[1].numel())
        return
    fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], colors, bench.WIDTH, bench.HEIGHT,
                          cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=24_000_000,
                          tight_tiles=not args.reference_tile_lists)
    for f in range(args.frames):
        bq, bt = bench.domino_poses(bench.N_BODIES, frame=60 + f, device=dev, centers=sc["body_centers"])
        torch.cuda.synchronize()
        fr.render(sc["viewmats"], sc["Ks"], bq, bt)
    torch.cuda.synchronize()
    print("n_isects", fr.n_isects())


if __name__ == "__main__":
    main()
