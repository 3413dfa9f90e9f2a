#!/usr/bin/env python
"""c5: Gaussian-sharded distributed render -- 20 M Gaussians split over the ranks, one 1080p camera per rank, projected
splats exchanged with the fused all-to-all of 3dgs_rigidbody_b200/distributed.py (NCCL over NVLink).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29533 \
        tools/bench_c5.py [--gaussians 20000000] [--steps 20] [--packed 1]
Prints one JSON line on rank 0 (frames/s = cameras rendered per second over all ranks; max over ranks, CUDA events)."""
import argparse
import importlib
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=20_000_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--packed", type=int, default=1)
    ap.add_argument("--exchange", choices=("peer", "nccl"), default="peer",
                    help="packed rows: NVLink peer-memory kernel (rs_exchange_push) or the NCCL all-to-all route")
    ap.add_argument("--profile", type=int, default=0, help="print the top CUDA ops of one step on rank 0 (torch.profiler)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rs = importlib.import_module("3dgs_rigidbody_b200")
    importlib.import_module("3dgs_rigidbody_b200.distributed").PeerSplatExchange.enabled = args.exchange == "peer"
    W, H = 1920, 1080
    n_local = args.gaussians // world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    # uniform in a 40 x 40 x 4 slab
    means = (torch.rand(n_local, 3, device=dev, generator=g) - 0.5) * torch.tensor([40.0, 40.0, 4.0], device=dev)
    quats = torch.nn.functional.normalize(torch.randn(n_local, 4, device=dev, generator=g), dim=-1)
    scales = torch.rand(n_local, 3, device=dev, generator=g) * 0.02
    opac = torch.rand(n_local, device=dev, generator=g)
    colors = torch.rand(n_local, 3, device=dev, generator=g)
    ang = 2 * math.pi * rank / max(world, 1)
    vm = torch.from_numpy(bench.look_at((30 * math.cos(ang), 30 * math.sin(ang), 12.0), (0, 0, 0))[None]).to(dev)
    f = 0.5 * W / math.tan(math.radians(30.0))
    Ks = torch.tensor([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]], dtype=torch.float32, device=dev)

    def step():
        with torch.no_grad():
            return rs.rasterization(means, quats, scales, opac, colors, vm, Ks, W, H, packed=bool(args.packed),
                                    distributed=True)

    for _ in range(args.warmup):
        img, alpha, meta = step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        img, alpha, meta = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    stats = torch.tensor([float(meta["flatten_ids"].numel()), float(alpha.mean())], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if args.profile:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        if rank == 0:
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    if rank == 0:
        ms = float(t[0]) / args.steps
        print(json.dumps({"workload": "c5: Gaussian-sharded render, all-to-all of projected splats", "n_gpus": world,
                          "gaussians_total": n_local * world, "cameras": world, "packed": bool(args.packed),
                          "exchange": (args.exchange if args.packed else "nccl"),
                          "ms_per_step": round(ms, 3), "frames_per_s": round(world / (ms * 1e-3), 2),
                          "n_isects_total": int(stats[0]), "mean_alpha": round(float(stats[1]) / world, 4)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
