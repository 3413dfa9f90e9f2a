"""Gaussian-sharded rendering (c5) on real GPUs: `rasterization(distributed=True)` on every rank against a single-GPU
render of the whole scene.  Runs with as many ranks as there are GPUs (1 on the round-end box: the collectives then
degenerate to copies but the whole code path -- camera gather, exchange, id globalisation -- still executes over NCCL)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, packed, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from conftest import pinhole_cameras, synthetic_scene

        rs = importlib.import_module("3dgs_rigidbody_b200")
        W, H, N = 320, 240, 30_000
        s = synthetic_scene(3, N, K=3)
        vm, Ks = pinhole_cameras(world, W, H)
        t = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
        vm, Ks = torch.from_numpy(vm).to(dev), torch.from_numpy(Ks).to(dev)
        rigid = dict(body_quats=t["body_quats"], body_trans=t["body_trans"], body_centers=t["body_centers"])
        # whole scene, this rank's camera, one GPU
        full, full_a, _ = rs.rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], vm[rank:rank + 1],
                                           Ks[rank:rank + 1], W, H, packed=packed, cluster_ids=t["cluster_ids"], **rigid)
        # Gaussian shard (contiguous block, so global ids keep the single-GPU order), own camera
        lo, hi = rank * N // world, (rank + 1) * N // world
        means = t["means"][lo:hi].clone().requires_grad_(True)
        img, alpha, meta = rs.rasterization(means, t["quats"][lo:hi], t["scales"][lo:hi], t["opacities"][lo:hi],
                                            t["colors"][lo:hi], vm[rank:rank + 1], Ks[rank:rank + 1], W, H, packed=packed,
                                            distributed=True, cluster_ids=t["cluster_ids"][lo:hi], **rigid)
        assert img.shape == full.shape
        err = float((img - full).abs().max())
        assert err <= 1e-4, err
        assert float((alpha - full_a).abs().max()) <= 1e-4
        assert meta["n_cameras"] == 1
        img.sum().backward()  # transposed exchange
        assert means.grad is not None and bool(torch.isfinite(means.grad).all()) and float(means.grad.abs().sum()) > 0
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("packed", [False, True])
def test_gaussian_sharded_render_matches_single_gpu(rs, packed):
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, packed, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
