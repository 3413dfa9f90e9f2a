#!/usr/bin/env python
"""c3 forward + backward (1 M Gaussians, 16 channels, 1080p) through rasterization(): step time over 20 steps and the
per-kernel times of one step, backward included (autograd forced onto the calling thread so that rs_profile_begin / _end
see its launches).  RIGIDSPLAT_LIB=<variant .so> times another build (tools/build_variant_lib.sh).

    python tools/time_c3.py"""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rs = importlib.import_module("3dgs_rigidbody_b200")
_lib = importlib.import_module("3dgs_rigidbody_b200._lib")
dev = "cuda:0"
W, H, D = bench.WIDTH, bench.HEIGHT, 16
sc = bench.make_domino_scene(device=dev)
g = torch.Generator(device=dev).manual_seed(42)
feats = torch.randn(sc["means"].shape[0], D, device=dev, generator=g).requires_grad_()
w = torch.rand(1, H, W, D, device=dev, generator=g)
leaves = [sc[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities")]
def step(f):
    bq, bt = bench.domino_poses(bench.N_BODIES, frame=60 + f, device=dev, centers=sc["body_centers"])
    for t in leaves + [feats]: t.grad = None
    img, _, _ = rs.rasterization(*leaves, feats, sc["viewmats"], sc["Ks"], W, H, packed=False, cluster_ids=sc["cluster_ids"], body_quats=bq, body_trans=bt, body_centers=sc["body_centers"])
    (img * w).sum().backward()
for f in range(5): step(f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for f in range(20): step(5 + f)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("RIGIDSPLAT_LIB", "default"), "c3 step ms", e0.elapsed_time(e1) / 20, "grad checksum", float(feats.grad.abs().sum()), float(leaves[0].grad.abs().sum()))
with torch.autograd.set_multithreading_enabled(False):
    prof = _lib.profile_kernels(lambda: step(30), torch.cuda.current_stream().cuda_stream)
print([(n, round(ms, 4)) for n, ms in prof], "sum", sum(ms for _, ms in prof))
