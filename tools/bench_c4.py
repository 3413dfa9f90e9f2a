#!/usr/bin/env python
"""c4 (BASELINE configs[3]): 6 M Gaussians, 500 rigid bodies, 8 ring cameras at 3840x2160; the cameras of every animation
frame are sharded over the ranks (camera c -> rank c % world), Gaussians replicated, NO collective on the data path.

    python tools/bench_c4.py                                   # 1 GPU renders all 8 cameras of every frame
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/bench_c4.py
Prints one JSON line on rank 0: camera-frames per second over all ranks (CUDA events, max over ranks)."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24, help="animation frames (each = 8 cameras)")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--in-flight", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    rs = importlib.import_module("3dgs_rigidbody_b200")
    W, H, N, K, C = 3840, 2160, 6_000_000, 500, 8
    sc = bench.make_domino_scene(N, K, device=dev, width=W, height=H, n_cameras=C)
    q_np, t_np = bench.domino_poses_np(K, None, sc["body_centers"].cpu().numpy())
    q_all, t_all = torch.from_numpy(q_np).to(dev), torch.from_numpy(t_np).to(dev)
    # camera c of animation frame f goes to rank (c + f) % world: every rank sees every viewpoint over the animation, so a
    # viewpoint with more intersections than the others does not pin one rank
    cams_of = lambda f: [c for c in range(C) if (c + f) % world == rank]
    pipe = rs.FramePipeline(args.in_flight, sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                            cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=96_000_000)

    def run(frames):
        for f in frames:
            for c in cams_of(f):
                pipe.submit(sc["viewmats"][c:c + 1], sc["Ks"][c:c + 1], q_all[f % 240], t_all[f % 240])
        pipe.join()

    run(range(60, 60 + args.warmup))
    torch.cuda.synchronize()
    assert not pipe.overflowed(), "max_isects too small for c4"
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(range(60 + args.warmup, 60 + args.warmup + args.frames))
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    n_isects = torch.tensor([float(pipe.renderers[0].n_isects())], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_isects, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"workload": "c4: 6M Gaussians, 500 bodies, 8 ring cameras 3840x2160, cameras sharded over ranks",
                          "n_gpus": world, "animation_frames": args.frames, "camera_frames": args.frames * C,
                          "ms_per_animation_frame": round(ms / args.frames, 3),
                          "camera_frames_per_s": round(args.frames * C / (ms * 1e-3), 2),
                          "n_isects_sample_camera_mean": int(float(n_isects[0]) / world)}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
