#!/usr/bin/env python
"""c4 alone (the same measurement bench.py reports under other_configs.c4): 6 M Gaussians, 500 bodies, 8 ring cameras at 4K,
the cameras of every animation frame sharded over the ranks.

    python tools/bench_c4.py [--frames 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/bench_c4.py"""
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
    ap.add_argument("--reference-tile-lists", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    rs = importlib.import_module("3dgs_rigidbody_b200")
    res = bench.bench_c4(rs, torch, dist, dev, rank, world, frames=args.frames, warmup=args.warmup, in_flight=args.in_flight,
                         tight_tiles=not args.reference_tile_lists)
    if rank == 0:
        print(json.dumps(res))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
