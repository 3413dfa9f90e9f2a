#!/usr/bin/env python
"""Per-stage timing of the REFERENCE's own CUDA kernels (oracle/_ref/gsplat_ref_cuda.so, built from /root/reference in
place) against this library on the same B200, same tensors ("the number to beat").

    python tools/compare_ref_cuda.py [--config c2|c3] [--reps 20] > profiles/rNN_vs_reference_cuda_<config>.json

c2: 1 M Gaussians, 20 bodies, 1080p, RGB, forward only.  Reference chain = torch rigid transform (one fused torch
    expression for all bodies -- cheaper than the reference's per-body apply_transform() with its tensor clones) ->
    projection_ewa_3dgs_fused_fwd -> intersect_tile (cub sort, host sync) -> intersect_offset -> rasterize_to_pixels_3dgs_fwd.
c3: same scene, 16 feature channels, forward + backward of compositing and projection.
Measurement infrastructure only (uses oracle/_ref); CUDA events, warm-up, median.
"""
import argparse
import importlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "gsplat_ref_cuda.so")
    spec = importlib.util.spec_from_file_location("gsplat_ref_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c2", "c3"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--frame", type=int, default=120)
    args = ap.parse_args()
    import __graft_entry__ as ge

    ge.build()
    rs = importlib.import_module("3dgs_rigidbody_b200")
    ref = load_ref()
    from test_gpu_vs_reference_cuda import _torch_rigid

    dev = "cuda:0"
    D = 3 if args.config == "c2" else 16
    W, H = bench.WIDTH, bench.HEIGHT
    sc = bench.make_domino_scene(bench.N_GAUSS, bench.N_BODIES, device=dev, channels=D)
    bq, bt = bench.domino_poses(bench.N_BODIES, frame=args.frame, device=dev, centers=sc["body_centers"])
    N = bench.N_GAUSS
    tw, th = (W + 15) // 16, (H + 15) // 16
    C_ = rs._C
    res = {"config": args.config, "gaussians": N, "channels": D, "width": W, "height": H, "frame": args.frame,
           "reps": args.reps, "unit": "ms (median, CUDA events)", "reference": {}, "ours": {}}

    # ---------------- reference chain ----------------
    t, (m_t, q_t) = timeit(lambda: _torch_rigid(sc["means"], sc["quats"], sc["cluster_ids"], bq, bt, sc["body_centers"]),
                           args.reps)
    res["reference"]["rigid transform (torch, all bodies in one expression)"] = t
    pa = (m_t, None, q_t, sc["scales"], sc["opacities"], sc["viewmats"], sc["Ks"], W, H, 0.3, 0.01, 1e10, 0.0, False)
    t, pr = timeit(lambda: ref.projection_ewa_3dgs_fused_fwd(*pa, ref.PINHOLE), args.reps)
    res["reference"]["projection_ewa_3dgs_fused_fwd"] = t
    radii, means2d, depths, conics, _ = pr
    t, isec = timeit(lambda: ref.intersect_tile(means2d, radii, depths, None, None, 1, 16, tw, th, True, False), args.reps)
    res["reference"]["intersect_tile (count + cumsum + emit + cub sort, 1 host sync)"] = t
    tpg, ids, flat = isec
    t, off = timeit(lambda: ref.intersect_offset(ids, 1, tw, th), args.reps)
    res["reference"]["intersect_offset"] = t
    colors = sc["colors"][None].contiguous()
    opac = sc["opacities"][None].contiguous()
    ra = (means2d, conics, colors, opac, None, None, W, H, 16, off, flat)
    t, (rc, ralpha, last) = timeit(lambda: ref.rasterize_to_pixels_3dgs_fwd(*ra), args.reps)
    res["reference"]["rasterize_to_pixels_3dgs_fwd"] = t
    res["n_isects"] = int(ids.numel())
    if args.config == "c3":
        g = torch.Generator(device=dev).manual_seed(0)
        v_rc = torch.randn(rc.shape, device=dev, generator=g)
        v_ra = torch.randn(ralpha.shape, device=dev, generator=g)
        t, gr = timeit(lambda: ref.rasterize_to_pixels_3dgs_bwd(*ra, ralpha, last, v_rc, v_ra, False), args.reps)
        res["reference"]["rasterize_to_pixels_3dgs_bwd"] = t
        _, v_m2, v_con, _, _ = gr
        v_d = torch.zeros_like(depths)
        t, _ = timeit(lambda: ref.projection_ewa_3dgs_fused_bwd(
            m_t, None, q_t, sc["scales"], sc["viewmats"], sc["Ks"], W, H, 0.3, ref.PINHOLE, radii, conics, None, v_m2, v_d,
            v_con, None, False), args.reps)
        res["reference"]["projection_ewa_3dgs_fused_bwd"] = t
    res["reference"]["total"] = float(sum(res["reference"].values()))

    # ---------------- ours: operator path (same boundaries as the reference) ----------------
    rp = rs.RigidPoses(sc["cluster_ids"], bq, bt, sc["body_centers"])
    po = (sc["means"], None, sc["quats"], sc["scales"], sc["opacities"], sc["viewmats"], sc["Ks"], W, H, 0.3, 0.01, 1e10,
          0.0, False)
    t, pr_o = timeit(lambda: C_.projection_ewa_3dgs_fused_fwd(*po, C_.PINHOLE, rp), args.reps)
    res["ours"]["rigid + projection (one kernel)"] = t
    radii_o, means2d_o, depths_o, conics_o, _ = pr_o
    t, isec_o = timeit(lambda: C_.intersect_tile(means2d_o, radii_o, depths_o, None, None, 1, 16, tw, th, True, False),
                       args.reps)
    res["ours"]["intersect_tile (depth-ordered binning, 1 host sync)"] = t
    _, ids_o, flat_o = isec_o
    t, off_o = timeit(lambda: C_.intersect_offset(ids_o, 1, tw, th), args.reps)
    res["ours"]["intersect_offset"] = t
    ro = (means2d_o, conics_o, colors, opac, None, None, W, H, 16, off_o, flat_o)
    t, (rc_o, ra_o, last_o) = timeit(lambda: C_.rasterize_to_pixels_3dgs_fwd(*ro), args.reps)
    res["ours"]["rasterize_to_pixels_3dgs_fwd (incl. record packing)"] = t
    if args.config == "c3":
        t, gr_o = timeit(lambda: C_.rasterize_to_pixels_3dgs_bwd(*ro, ra_o, last_o, v_rc, v_ra, False), args.reps)
        res["ours"]["rasterize_to_pixels_3dgs_bwd"] = t
        _, v_m2o, v_cono, _, _ = gr_o
        t, _ = timeit(lambda: C_.projection_ewa_3dgs_fused_bwd(
            sc["means"], None, sc["quats"], sc["scales"], sc["viewmats"], sc["Ks"], W, H, 0.3, C_.PINHOLE, radii_o, conics_o,
            None, v_m2o, torch.zeros_like(depths_o), v_cono, None, False, rp), args.reps)
        res["ours"]["rigid + projection bwd"] = t
    res["ours"]["total (operator path)"] = float(sum(res["ours"].values()))
    # ---------------- ours: frame path (what the animation loop runs) ----------------
    if args.config == "c2":
        fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                              cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=24_000_000)
        t, _ = timeit(lambda: fr.render(sc["viewmats"], sc["Ks"], bq, bt), args.reps)
        res["ours"]["frame path (rs_render_frame: one sync-free call)"] = t
        res["speedup_frame_path_vs_reference_chain"] = res["reference"]["total"] / t
    res["speedup_operator_path_vs_reference_chain"] = res["reference"]["total"] / res["ours"]["total (operator path)"]
    res["same_sorted_ids"] = bool(torch.equal(ids, ids_o)) if ids.numel() == ids_o.numel() else False
    res["max_abs_image_diff"] = float((rc - rc_o).abs().max())
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
