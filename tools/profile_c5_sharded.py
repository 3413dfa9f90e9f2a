#!/usr/bin/env python
"""Per-kernel times of one sync-free Gaussian-sharded frame (c5: 20 M Gaussians over the ranks, 8 cameras at 1080p) on every
rank, with one CUDA event behind each kernel launch (rs_profile_begin / rs_profile_end) -- the numbers quoted in DESIGN.md
section 5 ("Sync-free Gaussian-sharded frame").  TIGHT=0 uses the reference's tile lists.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/profile_c5_sharded.py"""
import importlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rs = importlib.import_module("3dgs_rigidbody_b200"); _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
dmod = importlib.import_module("3dgs_rigidbody_b200.distributed")
n_total, n_cams, W, H = 20_000_000, 8, 1920, 1080
cl = n_cams // world; mine = slice(rank * cl, (rank + 1) * cl)
lo, hi = rank * n_total // world, (rank + 1) * n_total // world
sc = bench._c5_scene(torch, dev, n_total, lo, hi)
vm, Ks = bench._c5_cameras(torch, dev, n_cams, W, H)
fr = dmod.ShardedFrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opac"], sc["colors"], W, H, cl, tight_tiles=bool(int(os.environ.get("TIGHT", "1"))))
for _ in range(3): fr.render(vm[mine], Ks[mine])
print(rank, fr.check(), flush=True)
torch.cuda.synchronize(); dist.barrier()
prof = _lib.profile_kernels(lambda: fr.render(vm[mine], Ks[mine]), torch.cuda.current_stream().cuda_stream)
dist.barrier()
for r in range(world):
    if r == rank:
        print("rank", rank, [(n, round(ms, 3)) for n, ms in prof], "sum", round(sum(ms for _, ms in prof), 3), flush=True)
    dist.barrier()
dist.destroy_process_group()
