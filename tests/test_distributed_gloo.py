"""World-size-2 `gloo` tests (CPU) of the multi-rank host logic: the Gaussian-shard exchange of
3dgs_rigidbody_b200/distributed.py (c5) and the frame sharding of bench.py (c2/c4).  The collectives run for real over
gloo; only the CUDA kernels are absent."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_problem(world, n_per_rank, cams_per_rank, D, seed=0, sparse=False):
    g = torch.Generator().manual_seed(seed)
    Ct, Nt = world * cams_per_rank, sum(n_per_rank)
    radii = torch.randint(0, 30, (Ct, Nt, 2), generator=g, dtype=torch.int32)
    if sparse:  # the last rank's Gaussians are visible nowhere (it packs zero rows) and its cameras see nothing at all
        radii[:, Nt - n_per_rank[-1]:] = 0
        radii[(world - 1) * cams_per_rank:] = 0
    return dict(
        radii=radii,
        means2d=torch.randn(Ct, Nt, 2, generator=g), depths=torch.rand(Ct, Nt, generator=g) + 1,
        conics=torch.randn(Ct, Nt, 3, generator=g), opacities=torch.rand(Ct, Nt, generator=g),
        colors=torch.rand(Nt, D, generator=g))


def _worker(rank, world, port, n_per_rank, cams, D, q, sparse=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex_mod = importlib.import_module("3dgs_rigidbody_b200.distributed")
        G = _global_problem(world, n_per_rank, cams, D, sparse=sparse)
        lo = sum(n_per_rank[:rank])
        hi = lo + n_per_rank[rank]
        ex = ex_mod.GaussianShardExchange(n_per_rank[rank], cams, torch.device("cpu"))
        assert ex.n_per_rank == list(n_per_rank) and ex.gaussian_base == lo and ex.total_cameras == world * cams
        # cameras
        vm = torch.arange(cams * 16, dtype=torch.float32).reshape(cams, 4, 4) + 1000 * rank
        Ks = torch.arange(cams * 9, dtype=torch.float32).reshape(cams, 3, 3) + 1000 * rank
        vm_all, Ks_all = ex.gather_cameras(vm, Ks)
        assert vm_all.shape == (world * cams, 4, 4) and torch.equal(vm_all[rank * cams:(rank + 1) * cams], vm)
        assert float(vm_all[(1 - rank) * cams, 0, 0]) == 1000 * (1 - rank)
        assert torch.equal(Ks_all[rank * cams:(rank + 1) * cams], Ks)
        my_cams = slice(rank * cams, (rank + 1) * cams)

        # ---- dense mode: [C_total, N_local] in -> [C_local, N_total] out -------------------------------------------------
        m2 = G["means2d"][:, lo:hi].clone().requires_grad_(True)
        col = G["colors"][lo:hi].clone().requires_grad_(True)
        out = ex.exchange(False, G["radii"][:, lo:hi].contiguous(), m2, G["depths"][:, lo:hi], G["conics"][:, lo:hi],
                          G["opacities"][:, lo:hi], col[None].expand(world * cams, -1, -1), None, None)
        radii, means2d, depths, conics, opac, colors, cid, gid = out
        assert cid is None and gid is None
        assert torch.equal(radii, G["radii"][my_cams])
        assert torch.equal(means2d, G["means2d"][my_cams]) and torch.equal(depths, G["depths"][my_cams])
        assert torch.equal(conics, G["conics"][my_cams]) and torch.equal(opac, G["opacities"][my_cams])
        assert torch.equal(colors, G["colors"][None].expand(cams, -1, -1))
        # backward = transposed exchange: every (camera, local Gaussian) row receives its cotangent back
        w = torch.arange(world * cams, dtype=torch.float32)[my_cams].view(cams, 1, 1) + 1.0
        (means2d * w).sum().backward(retain_graph=True)
        expect = (torch.arange(world * cams, dtype=torch.float32) + 1.0).view(-1, 1, 1).expand(-1, hi - lo, 2)
        assert torch.equal(m2.grad, expect)
        colors.sum().backward()
        assert torch.equal(col.grad, torch.full_like(col, float(world * cams)))

        # ---- packed mode ---------------------------------------------------------------------------------------------------
        vis = (G["radii"][:, lo:hi] > 0).all(-1)
        cam_ids, g_ids = torch.nonzero(vis, as_tuple=True)
        pick = lambda t: t[:, lo:hi][cam_ids, g_ids]
        out = ex.exchange(True, pick(G["radii"]), pick(G["means2d"]), pick(G["depths"]), pick(G["conics"]),
                          pick(G["opacities"]), G["colors"][lo:hi][g_ids], cam_ids, g_ids)
        radii, means2d, depths, conics, opac, colors, cid, gid = out
        # expected: visible rows of my cameras from every source rank in rank order, global gaussian ids, local camera ids
        vis_all = (G["radii"] > 0).all(-1)
        exp_c, exp_g = [], []
        for s in range(world):
            slo = sum(n_per_rank[:s])
            c_, g_ = torch.nonzero(vis_all[my_cams, slo:slo + n_per_rank[s]], as_tuple=True)
            exp_c.append(c_)
            exp_g.append(g_ + slo)
        exp_c, exp_g = torch.cat(exp_c), torch.cat(exp_g)
        assert torch.equal(cid, exp_c) and torch.equal(gid, exp_g)
        gc = exp_c + rank * cams
        assert torch.equal(radii, G["radii"][gc, exp_g]) and torch.equal(means2d, G["means2d"][gc, exp_g])
        assert torch.equal(depths, G["depths"][gc, exp_g]) and torch.equal(conics, G["conics"][gc, exp_g])
        assert torch.equal(opac, G["opacities"][gc, exp_g]) and torch.equal(colors, G["colors"][exp_g])
        if sparse:  # zero rows packed on the last rank, zero rows received by it: the edge that must not strand a rank
            assert (cam_ids.numel() == 0) == (rank == world - 1) and (gid.numel() == 0) == (rank == world - 1)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_per_rank,cams,D", [((700, 500), 2, 3), ((64, 64), 1, 16)])
def test_gaussian_shard_exchange_world2_gloo(n_per_rank, cams, D):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_per_rank, cams, D, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_gaussian_shard_exchange_with_empty_ranks_gloo():
    """Packed exchange when one rank has nothing to send and nothing to receive (its Gaussians are culled everywhere, its
    cameras look away): sizes of zero must travel through the count exchange and both all-to-alls."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, (300, 200), 2, 3, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_frame_sharding_is_a_partition():
    """bench.py / FrameRenderer multi-GPU mode: rank r renders frames r, r+P, ... -- every frame exactly once."""
    sys.path.insert(0, ROOT)
    import bench

    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in bench.frames_of_rank(r, world, bench.N_FRAMES // world))
        assert seen == list(range(bench.N_FRAMES // world * world))
    q, t = bench.domino_poses_np(20, [0, 100, 239])
    assert q.shape == (3, 20, 4) and t.shape == (3, 20, 3)
    assert np.allclose(np.linalg.norm(q, axis=-1), 1.0, atol=1e-6)
    assert np.allclose(t[0, 1:], 0.0, atol=1e-6)  # frame 0: bodies 1.. have not started tipping
