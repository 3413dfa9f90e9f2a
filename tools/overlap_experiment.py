#!/usr/bin/env python
"""Where does the pipelined frame time go?  Times, on one B200, `depth` frames in flight of (a) the whole frame, (b) only
projection + binning, (c) only compositing (of an already binned frame).  If (b) + (c) ~= (a) the two halves do not
overlap (they compete for SM residency) and only less resource-time per half helps."""
import argparse
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--frames", type=int, default=240)
    args = ap.parse_args()
    rs = importlib.import_module("3dgs_rigidbody_b200")
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    lib = _lib.load()
    dev = "cuda:0"
    sc = bench.make_domino_scene(device=dev)
    q_np, t_np = bench.domino_poses_np(bench.N_BODIES, None, sc["body_centers"].cpu().numpy())
    q_all, t_all = torch.from_numpy(q_np).to(dev), torch.from_numpy(t_np).to(dev)
    frs = [rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], bench.WIDTH, bench.HEIGHT,
                            cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=24_000_000)
           for _ in range(args.depth)]
    streams = [torch.cuda.Stream() for _ in range(args.depth)]

    def run(stages, n):
        for i in range(n):
            k = i % args.depth
            f = i % 240
            with torch.cuda.stream(streams[k]):
                a = frs[k]._fill(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
                a.stages = stages
                _lib.check(lib.rs_render_frame(ctypes.byref(a), streams[k].cuda_stream))

    for stages, name in ((0, "whole frame"), (1, "projection + binning only"), (2, "compositing only"), (0, "whole frame")):
        run(0, 2 * args.depth)  # every workspace holds a binned frame
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(stages, args.frames)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        print(f"depth {args.depth}  {name:28s} {e0.elapsed_time(e1) / args.frames * 1e3:8.1f} us/frame")


if __name__ == "__main__":
    main()
