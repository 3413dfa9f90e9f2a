"""Gaussian-sharded rendering (c5) on real GPUs: `rasterization(distributed=True)` on every rank against a single-GPU
render of the whole scene.  Runs with as many ranks as there are GPUs (up to 8; 1 on the round-end box: the collectives then
degenerate to copies but the whole code path -- camera gather, exchange, id globalisation -- still executes over NCCL; the
multi-rank runs are kept under profiles/ and bench.py --gpus N carries its own sharded == single-GPU check)."""
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
    world = min(torch.cuda.device_count(), 8)
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


def _peer_worker(rank, world, port, q):
    """no_grad + packed: the NVLink peer-memory exchange (rs_exchange_push / _wait) must deliver exactly the rows the NCCL
    all-to-all delivers -- same order, same values -- for several frames, through a forced regrow, with per-Gaussian
    colours, per-row (SH) colours and antialiased compensations."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from conftest import pinhole_cameras, synthetic_scene

        rs = importlib.import_module("3dgs_rigidbody_b200")
        dmod = importlib.import_module("3dgs_rigidbody_b200.distributed")
        dmod.PeerSplatExchange.initial_capacity = 64  # far too small: the first frame must regrow on every rank
        W, H, N, Cl = 320, 240, 30_000, 2
        s = synthetic_scene(3, N, K=3)
        vm, Ks = pinhole_cameras(world * Cl, W, H)
        t = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
        vm, Ks = torch.from_numpy(vm).to(dev), torch.from_numpy(Ks).to(dev)
        mine = slice(rank * Cl, (rank + 1) * Cl)
        lo, hi = rank * N // world, (rank + 1) * N // world
        g = torch.Generator(device=dev).manual_seed(7)
        sh = torch.randn(N, 16, 3, device=dev, generator=g) * 0.2
        for frame in range(3):
            rigid = dict(cluster_ids=t["cluster_ids"][lo:hi], body_quats=t["body_quats"],
                         body_trans=t["body_trans"] + 0.05 * frame, body_centers=t["body_centers"])
            for variant in ("rgb", "sh", "aa"):
                colors = sh[lo:hi] if variant == "sh" else t["colors"][lo:hi]
                kw = dict(packed=True, distributed=True, sh_degree=3 if variant == "sh" else None,
                          rasterize_mode="antialiased" if variant == "aa" else "classic",
                          render_mode="RGB+ED" if variant == "aa" else "RGB", **rigid)
                args = (t["quats"][lo:hi], t["scales"][lo:hi], t["opacities"][lo:hi], colors, vm[mine], Ks[mine], W, H)
                with torch.no_grad():
                    img_p, alpha_p, meta_p = rs.rasterization(t["means"][lo:hi], *args, **kw)
                    keep = {k: meta_p[k].clone() for k in ("gaussian_ids", "camera_ids", "radii", "means2d", "opacities")}
                means = t["means"][lo:hi].clone().requires_grad_(True)  # grad needed -> NCCL all-to-all route
                img_n, alpha_n, meta_n = rs.rasterization(means, *args, **kw)
                assert img_p.shape == (Cl, H, W, 4 if variant == "aa" else 3)
                for k, v in keep.items():
                    assert torch.equal(v, meta_n[k].detach()), (frame, variant, k)
                if variant == "sh":  # no-grad: SH evaluated inside the projection kernel; grad: the separate SH operator
                    assert float((img_p - img_n.detach()).abs().max()) <= 2e-5, (frame, variant)
                else:
                    assert torch.equal(img_p, img_n.detach()), (frame, variant)
                assert torch.equal(alpha_p, alpha_n.detach()), (frame, variant)
        # (no per-rank content assertion here: a rank whose cameras look away from the scene legitimately receives no rows,
        # and a rank leaving early would strand the others inside a collective)
        # nothing visible anywhere (a near plane beyond the whole scene culls every Gaussian for every camera of the ring,
        # whatever the world size): zero rows sent and received on both routes
        gone = t["means"][lo:hi]
        args = (t["quats"][lo:hi], t["scales"][lo:hi], t["opacities"][lo:hi], t["colors"][lo:hi], vm[mine], Ks[mine], W, H)
        with torch.no_grad():
            img_p, alpha_p, meta_p = rs.rasterization(gone, *args, packed=True, distributed=True, near_plane=1e6, far_plane=1e7)
        img_n, alpha_n, _ = rs.rasterization(gone.clone().requires_grad_(True), *args, packed=True, distributed=True,
                                             near_plane=1e6, far_plane=1e7)
        assert meta_p["gaussian_ids"].numel() == 0 and float(alpha_p.abs().max()) == 0.0
        assert torch.equal(img_p, img_n.detach()) and torch.equal(alpha_p, alpha_n.detach())
        peer = next(iter(dmod.PeerSplatExchange._instances.values()))
        assert peer.buffers.capacity > 64 and peer.epoch >= 10
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_peer_memory_exchange_equals_nccl_exchange(rs):
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def _peer_grad_worker(rank, world, port, q):
    """Training through the peer-memory exchange: forward rows pushed over NVLink, gradients pushed back by the transposed
    exchange (rs_exchange_push_grad) -- against the differentiable NCCL all-to-all route on the same inputs, two steps in a row
    (the second reuses the persistent gradient arrays), plus the single-GPU gradients of this rank's shard."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from conftest import pinhole_cameras, synthetic_scene

        rs = importlib.import_module("3dgs_rigidbody_b200")
        dmod = importlib.import_module("3dgs_rigidbody_b200.distributed")
        W, H, N, Cl = 256, 192, 24_000, 2
        s = synthetic_scene(5, N, K=3)
        vm, Ks = pinhole_cameras(world * Cl, W, H)
        t = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
        vm, Ks = torch.from_numpy(vm).to(dev), torch.from_numpy(Ks).to(dev)
        mine = slice(rank * Cl, (rank + 1) * Cl)
        lo, hi = rank * N // world, (rank + 1) * N // world
        g = torch.Generator(device=dev).manual_seed(11)
        wgt = torch.rand(Cl, H, W, 3, device=dev, generator=g)
        rigid = dict(cluster_ids=t["cluster_ids"][lo:hi], body_quats=t["body_quats"], body_trans=t["body_trans"],
                     body_centers=t["body_centers"])
        names = ("means", "quats", "scales", "opacities", "colors")

        def step(route_peer, shift):
            dmod.PeerSplatExchange.differentiable = route_peer
            leaves = [t[k][lo:hi].clone().requires_grad_(True) for k in names]
            leaves[0].data += shift
            img, alpha, meta = rs.rasterization(*leaves, vm[mine], Ks[mine], W, H, packed=True, distributed=True, **rigid)
            ((img * wgt).sum() + alpha.sum()).backward()
            return img.detach(), [l.grad.clone() for l in leaves]

        for shift in (0.0, 0.05):
            img_n, grads_n = step(False, shift)
            img_p, grads_p = step(True, shift)
            assert torch.equal(img_p, img_n), shift
            for name, gp, gn in zip(names, grads_p, grads_n):
                scale = max(float(gn.abs().max()), 1e-12)
                assert float((gp - gn).abs().max()) <= 2e-4 * scale, (shift, name, float((gp - gn).abs().max()), scale)
                assert float(gn.abs().sum()) > 0 or name == "quats"
        peer = next(p for p in dmod.PeerSplatExchange._instances.values() if p.grad_buffers is not None)
        assert peer.grad_epoch == 2
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_peer_memory_backward_equals_nccl_backward(rs):
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def _syncfree_worker(rank, world, port, q):
    """ShardedFrameRenderer (no host read between projection and image) against rasterization(distributed=True, packed=True)
    on the same shards: bit-equal frames over an animation, sizes confirmed by the single check() read."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from conftest import pinhole_cameras, synthetic_scene

        rs = importlib.import_module("3dgs_rigidbody_b200")
        W, H, N, Cl = 320, 240, 40_000, 2
        s = synthetic_scene(8, N, K=3)
        vm, Ks = pinhole_cameras(world * Cl, W, H)
        t = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
        vm, Ks = torch.from_numpy(vm).to(dev), torch.from_numpy(Ks).to(dev)
        mine = slice(rank * Cl, (rank + 1) * Cl)
        lo, hi = rank * N // world, (rank + 1) * N // world
        shard = [t[k][lo:hi].contiguous() for k in ("means", "quats", "scales", "opacities", "colors")]
        ids = t["cluster_ids"][lo:hi].contiguous()
        fr = rs.ShardedFrameRenderer(*shard, W, H, Cl, cluster_ids=ids, body_centers=t["body_centers"])
        fr_tight = rs.ShardedFrameRenderer(*shard, W, H, Cl, cluster_ids=ids, body_centers=t["body_centers"], tight_tiles=True)
        for frame in range(4):
            bt = t["body_trans"] + 0.07 * frame
            img, alpha = fr.render(vm[mine], Ks[mine], t["body_quats"], bt)
            got_img, got_alpha = img.clone(), alpha.clone()
            info = fr.check()
            assert not info["regrow"], info
            with torch.no_grad():
                want, want_a, meta = rs.rasterization(*shard, vm[mine], Ks[mine], W, H, packed=True, distributed=True,
                                                      cluster_ids=ids, body_quats=t["body_quats"], body_trans=bt,
                                                      body_centers=t["body_centers"])
            assert info["rows"] == meta["gaussian_ids"].numel() and info["n_isects"] == meta["flatten_ids"].numel(), (frame, info)
            assert torch.equal(got_img, want) and torch.equal(got_alpha, want_a), frame
            # tight tile lists (rs_isect_footprints with conics + opacities): fewer intersections, the same pixels
            img_t, alpha_t = fr_tight.render(vm[mine], Ks[mine], t["body_quats"], bt)
            info_t = fr_tight.check()
            assert not info_t["regrow"] and info_t["rows"] == info["rows"] and info_t["n_isects"] <= info["n_isects"], info_t
            assert torch.equal(img_t, want) and torch.equal(alpha_t, want_a), frame
        # an intersection workspace that is too small is reported by check() -- on every rank -- and re-sized; the frame
        # rendered again is the right one
        fr._alloc_render(fr._alloc_rows, 1024)
        fr.render(vm[mine], Ks[mine], t["body_quats"], bt)
        info = fr.check()
        assert info["regrow"] and fr.max_isects >= info["n_isects"] > 1024, info
        img2, alpha2 = fr.render(vm[mine], Ks[mine], t["body_quats"], bt)
        assert not fr.check()["regrow"]
        assert torch.equal(img2, want) and torch.equal(alpha2, want_a)
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sync_free_sharded_frame_equals_distributed_rasterization(rs):
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_syncfree_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
