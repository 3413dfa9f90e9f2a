#!/usr/bin/env python
"""c5 alone (the same measurement bench.py reports under other_configs.c5): 20 M Gaussians sharded over the ranks, 8 ring
cameras at 1080p, projected splats exchanged over NVLink peer memory (rs_exchange_push) and over NCCL, plus the
sharded == single-GPU image check on a reduced scene.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29533 \
        tools/bench_c5.py [--gaussians 20000000] [--steps 20]"""
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
    ap.add_argument("--gaussians", type=int, default=20_000_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    rs = importlib.import_module("3dgs_rigidbody_b200")
    res = bench.bench_c5(rs, torch, dist, dev, rank, world, n_total=args.gaussians, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        print(json.dumps(res))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
