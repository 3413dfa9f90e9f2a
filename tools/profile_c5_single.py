#!/usr/bin/env python
"""The Gaussian-sharded path in ONE process (world size 1, NCCL): rasterization(distributed=True, packed=True) without and
with gradients, so that `ncu` (which must not wrap a multi-rank command) can capture rs_exchange_push_kernel,
rs_exchange_push_grad_kernel and the packed projection.  The peer "exchange" then stores into this rank's own receive
arrays: the kernels do all of their work, only the NVLink hop is missing.

    ncu --set full --clock-control none --import-source on -k regex:"rs_exchange|rs_project_fwd_staged" -c 8 \
        -o gpurun_out/prof_c5 python tools/profile_c5_single.py"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    rs = importlib.import_module("3dgs_rigidbody_b200")
    n, W, H = 4_000_000, bench.WIDTH, bench.HEIGHT
    sc = bench._c5_scene(torch, dev, n, 0, n)
    vm, Ks = bench._c5_cameras(torch, dev, 2, W, H)
    for _ in range(2):
        with torch.no_grad():
            img, alpha, meta = rs.rasterization(sc["means"], sc["quats"], sc["scales"], sc["opac"], sc["colors"], vm, Ks, W, H,
                                                packed=True, distributed=True)
    colors = sc["colors"].clone().requires_grad_(True)
    for _ in range(2):
        colors.grad = None
        img, alpha, meta = rs.rasterization(sc["means"], sc["quats"], sc["scales"], sc["opac"], colors, vm, Ks, W, H,
                                            packed=True, distributed=True)
        img.sum().backward()
    torch.cuda.synchronize()
    print("rows", int(meta["gaussian_ids"].numel()), "grad", float(colors.grad.abs().sum()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
