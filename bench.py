#!/usr/bin/env python
"""bench.py -- frames/s of the rigid-animated render path (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): synthetic domino scene, 1 M Gaussians in 20 rigid bodies,
240-frame animation, one 1920x1080 pinhole camera.  A "step" is one frame: rigid animate -> EWA project ->
tile-intersect + radix sort -> front-to-back compositing (the commented-out loop of the reference, main.py:357-409:
apply_transform() per body + rasterization()).

  value     frames/s over all ranks, everything resident in HBM (poses [240,K,.] pre-generated on the device),
            K frames enqueued back to back through FramePipeline (`--in-flight` frames per GPU, one stream + workspace
            each, so the latency-bound binning of one frame overlaps the compositing of another), CUDA events on the
            launching stream, max over ranks.
  e2e       frames/s through the public per-frame API (FrameRenderer.render -> C ABI rs_render_frame) with HOST inputs
            and HOST results: per step the frame's poses + camera are copied from pinned host memory and the rendered
            float32 image [H,W,3] is copied back to pinned host memory (what the reference's loop keeps,
            main.py:387-400); copies are inside the timed region.  At ~25 MB per frame this leg is PCIe-bound.
  roofline  the dominant HBM-bound kernel of the step (radix-sort scatter pass), timed live with CUDA events.
  stages    per-stage CUDA-event times of one frame through the separate C-ABI entry points, with achieved GB/s against
            the algorithmic bytes of SURVEY.md section 8(d) -- explains `value`.
  cpu_baseline  the CPU port of the reference path (oracle/, OpenMP, all host threads) on a bounded sample of frames.

Multi-GPU: frames shard across ranks (rank r renders frames r, r+P, ...), Gaussians replicated, no collective on the
data path ("weak" scaling: every rank renders K frames).  `--impl reference` times the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GAUSS = 1_000_000
N_BODIES = 20
N_FRAMES = 240
WIDTH, HEIGHT = 1920, 1080
WORKLOAD = "c2: synthetic domino scene, 1M Gaussians, 20 rigid bodies, 240-frame animation, 1 camera 1920x1080"


# ----------------------------------------------------------------------------------------------------------------------
# synthetic domino scene: numpy, seeded, identical on every rank and for both arms
# ----------------------------------------------------------------------------------------------------------------------
HALF_EXTENTS = np.array([0.05, 0.25, 0.5], np.float32)  # thin along x, standing along z (z up)
SPACING = 0.35
GRID_COLS, GRID_ROW_SPACING = 20, 0.9


def _quat_about_y(theta):
    """wxyz quaternion of a rotation by theta about +y."""
    return np.stack([np.cos(theta / 2), np.zeros_like(theta), np.sin(theta / 2), np.zeros_like(theta)], -1)


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """world->camera 4x4 (OpenCV convention: +z forward, +x right, +y down)."""
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f], 0)
    vm = np.eye(4)
    vm[:3, :3] = R
    vm[:3, 3] = -R @ eye
    return vm.astype(np.float32)


def make_domino_scene_np(n_gauss=N_GAUSS, n_bodies=N_BODIES, width=WIDTH, height=HEIGHT, seed=42, s_max=0.02,
                         permute=True, channels=3, n_cameras=1):
    rng = np.random.default_rng(seed)
    per = n_gauss // n_bodies
    ids = np.minimum(np.arange(n_gauss) // per, n_bodies - 1).astype(np.int32)
    u = rng.random((n_gauss, 3), dtype=np.float32) * 2 - 1
    means = u * HALF_EXTENTS
    # bodies along a line (c2) or, beyond 40 bodies, on a 20-wide grid (c4: 20 x 25)
    means[:, 0] += SPACING * (ids % GRID_COLS if n_bodies > 40 else ids)
    if n_bodies > 40:
        means[:, 1] += GRID_ROW_SPACING * (ids // GRID_COLS)
    means[:, 2] += HALF_EXTENTS[2]  # bottom face on z = 0
    quats = rng.standard_normal((n_gauss, 4), dtype=np.float32)
    quats /= np.linalg.norm(quats, axis=1, keepdims=True)
    scales = rng.random((n_gauss, 3), dtype=np.float32) * s_max
    opacities = rng.random(n_gauss, dtype=np.float32)
    colors = rng.random((n_gauss, channels), dtype=np.float32)
    if permute:  # the kernels must not assume ids sorted by body
        p = rng.permutation(n_gauss)
        means, quats, scales, opacities, colors, ids = means[p], quats[p], scales[p], opacities[p], colors[p], ids[p]
    centers = np.zeros((n_bodies, 3), np.float32)
    for k in range(n_bodies):
        centers[k] = means[ids == k].mean(0)
    row_mid = 0.5 * SPACING * (min(n_bodies, GRID_COLS) - 1)
    if n_cameras == 1:
        viewmats = look_at((-2.0, -3.0, 1.5), (row_mid * 0.6, 0.0, 0.4))[None]
    else:  # ring of cameras around the scene centre (c4)
        cy = 0.5 * GRID_ROW_SPACING * ((n_bodies - 1) // GRID_COLS) if n_bodies > 40 else 0.0
        rad = 1.2 * max(row_mid, cy) + 4.0
        viewmats = np.stack([look_at((row_mid + rad * math.cos(2 * math.pi * c / n_cameras),
                                      cy + rad * math.sin(2 * math.pi * c / n_cameras), 2.5), (row_mid, cy, 0.4))
                             for c in range(n_cameras)])
    f = 0.5 * width / math.tan(math.radians(30.0))
    Ks = np.tile(np.array([[[f, 0, width / 2], [0, f, height / 2], [0, 0, 1]]], np.float32), (n_cameras, 1, 1))
    return dict(means=means, quats=quats, scales=scales, opacities=opacities, colors=colors, cluster_ids=ids,
                body_centers=centers, viewmats=viewmats, Ks=Ks)


def domino_poses_np(n_bodies=N_BODIES, frames=None, centers=None):
    """Pose stream [F,K,4] wxyz / [F,K,3]: body k tips about its bottom edge (x = x_k + hx, z = 0) by
    theta = clamp((f - 8k)/24, 0, 1) * 80 deg.  Translations are relative to rotation about `centers` (the reference's
    apply_transform pivot, main.py:210): t = R (c - e) + e - c."""
    frames = np.arange(N_FRAMES) if frames is None else np.atleast_1d(np.asarray(frames))
    k = np.arange(n_bodies)
    theta = np.clip((frames[:, None] - 8.0 * (k % 40)[None, :]) / 24.0, 0.0, 1.0) * math.radians(80.0)
    q = _quat_about_y(theta).astype(np.float32)  # [F,K,4]
    col = k % GRID_COLS if n_bodies > 40 else k
    row_y = GRID_ROW_SPACING * (k // GRID_COLS) if n_bodies > 40 else np.zeros(n_bodies)
    if centers is None:
        centers = np.stack([SPACING * col, row_y, np.full(n_bodies, HALF_EXTENTS[2])], -1)
    e = np.stack([SPACING * col + HALF_EXTENTS[0], row_y, np.zeros(n_bodies)], -1)  # pivot edge
    c, s = np.cos(theta), np.sin(theta)
    d = (centers - e)[None]  # [1,K,3]
    Rd = np.stack([c * d[..., 0] + s * d[..., 2], np.broadcast_to(d[..., 1], c.shape), -s * d[..., 0] + c * d[..., 2]], -1)
    t = (Rd + e[None] - centers[None]).astype(np.float32)
    return q, t


def frames_of_rank(rank, world, count, start=0):
    """Animation frames rendered by `rank`: start + rank, start + rank + world, ... (`count` of them, modulo 240)."""
    return [(start + rank + world * i) % N_FRAMES for i in range(count)]


def make_domino_scene(n_gauss=N_GAUSS, n_bodies=N_BODIES, device="cuda:0", **kw):
    import torch

    sc = make_domino_scene_np(n_gauss, n_bodies, **kw)
    return {k: torch.from_numpy(v).to(device) for k, v in sc.items()}


def domino_poses(n_bodies=N_BODIES, frame=0, device="cuda:0", centers=None):
    import torch

    if centers is not None and not isinstance(centers, np.ndarray):
        centers = centers.cpu().numpy()
    q, t = domino_poses_np(n_bodies, [frame], centers)
    return torch.from_numpy(q[0]).to(device), torch.from_numpy(t[0]).to(device)


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0 - 0.1) and (t1 is None or t <= t1 + 0.2)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU port of the reference path (oracle/), all host threads
# ----------------------------------------------------------------------------------------------------------------------
def run_cpu_port(sc, frames, budget_s=None):
    """Times oracle.render() (apply_transform per body -> projection -> isect/sort/offsets -> compositing, CPU, OpenMP)
    on the given animation frames.  Returns (seconds per frame list, threads)."""
    from oracle import oracle

    oracle.build()
    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    q, t = domino_poses_np(sc["body_centers"].shape[0], frames, sc["body_centers"])
    times = []
    t_begin = time.perf_counter()
    for i, _ in enumerate(frames):
        t0 = time.perf_counter()
        oracle.render(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], sc["viewmats"], sc["Ks"],
                      WIDTH, HEIGHT, cluster_ids=sc["cluster_ids"], body_quats=q[i], body_trans=t[i],
                      body_centers=sc["body_centers"])
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    return times, oracle.num_threads()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc = make_domino_scene_np()
    frames = [(60 + 7 * i) % N_FRAMES for i in range(args.warmup + args.steps)]
    times, threads = run_cpu_port(sc, frames)
    timed = times[args.warmup:]
    fps = len(timed) / sum(timed)
    sample = f"{len(timed)} full frames of the 240-frame animation after {args.warmup} warm-up frame(s)"
    line = {
        "impl": "reference", "metric": "frames/sec (1M Gaussians, 1080p, rigid-animated)", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(timed) / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "gaussians": N_GAUSS, "bodies": N_BODIES, "width": WIDTH, "height": HEIGHT,
                   "arm": "CPU port of the reference path (oracle/oracle.c, OpenMP); the reference's own python CPU "
                          "path cannot travel to the GPU box and cannot composite without its CUDA extension"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def stage_breakdown(fr, sc, q_all, t_all, frames, peak_gbs, peak_src, n_tiles):
    """Per-stage CUDA-event times of the SAME frames through the frame path itself (rs_render_frame_timed records events
    between the stages of rs_render_frame).  Returns (stages, roofline): `stages` lists every stage with its algorithmic
    bytes (SURVEY.md section 8d formulas, with this scene's measured M) and achieved GB/s; `roofline` is the dominant
    kernel of the step (compositing: one launch of rs_raster_fwd_kernel per frame)."""
    ms = np.array([fr.render_timed(sc["viewmats"], sc["Ks"], q_all[f], t_all[f]) for f in frames], np.float64)
    mean = ms.mean(0)
    N, D, C = fr.N, fr.D, fr.C
    E, P = C * N, C * fr.W * fr.H
    M = fr.n_isects()
    key_bits = 32 + int(n_tiles).bit_length() + int(C).bit_length()
    ref_passes = (key_bits + 7) // 8
    tile_passes = (key_bits - 32 + 7) // 8
    alg = {
        "rigid+project": N * 48 + E * 32,
        # what the reference's count + cumsum + emission + 64-bit cub sort + offsets move
        "binning": E * 32 + (E * 20 + M * 12) + (M * 8 + ref_passes * 2 * M * 12) + (M * 8 + n_tiles * C * 4),
        "composite": M * (4 + 28 + 4 * D) + P * (D + 2) * 4,
    }
    # what the depth-ordered scheme actually has to move: depth sort (hist + 4 passes of 8-byte pairs), ordered count
    # and emission, tile sort (hist + passes of 8-byte pairs), offsets
    moved_binning = E * 4 + 4 * E * 16 + E * 8 + (E * 24 + M * 8) + M * 4 + tile_passes * M * 16 + M * 4
    stages = []
    for k, name in enumerate(("rigid+project", "binning", "composite")):
        t = float(mean[k]) * 1e-3
        st = {"stage": name, "ms": round(float(mean[k]), 4), "algorithmic_MB": round(alg[name] / 1e6, 1),
              "GBps": round(alg[name] / t / 1e9, 1), "frac_of_hbm_peak": round(alg[name] / t / 1e9 / peak_gbs, 4)}
        if name == "binning":
            st["bytes_actually_moved_MB"] = round(moved_binning / 1e6, 1)
            st["GBps_actually_moved"] = round(moved_binning / t / 1e9, 1)
            st["note"] = ("algorithmic bytes = the reference's 64-bit LSD sort accounting (SURVEY 8d); the depth-ordered "
                          "scheme moves far fewer bytes, so the first fraction may exceed what the kernels stream")
        stages.append(st)
    t_c = float(mean[2]) * 1e-3
    achieved = alg["composite"] / t_c / 1e9
    roofline = {
        "bound": "hbm", "kernel": "rs_raster_fwd_kernel (compositing, 1 launch/frame: the largest share of the step)",
        "achieved": round(achieved, 1), "peak": peak_gbs, "unit": "GB/s", "frac": round(achieved / peak_gbs, 4),
        "traffic": 61.26e6, "traffic_source": "ncu --set full r01 (profiles/r01_frame_kernels_ncu_full.txt): dram read 48.80 MB + write 12.46 MB per launch",
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg["composite"], "ms_per_launch": round(float(mean[2]), 4),
        "frames_timed": int(len(frames)),
        "note": "this kernel is bound by SM issue (FP32 FMA + MUFU.EX2 + LDS), not by HBM: 66 % issue-slot utilisation, "
                "1.09e8 warp instructions per launch, 4 % DRAM throughput (profiles/r01_frame_kernels_ncu_full.txt); the gathers "
                "hit in L2, so DRAM traffic is far BELOW the algorithmic bytes; the other stages are in `stages`",
        "issue_slot_utilisation": 0.662,
    }
    return stages, roofline, float(mean[3]), M


def other_configs(rs, sc, q_all, t_all, steps=10):
    """Extra information next to the c2 headline (not the metric): the same scene through the drop-in `rasterization()` call
    (operator path, with the host reads the reference API implies), and BASELINE configs[2] (c3): 16 identity-feature
    channels, forward + backward, plus the contrastive clustering loss on the rendered map.  CUDA events, ms per step."""
    import torch

    dev = sc["means"].device

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / steps, 4)

    rigid = lambda f: dict(cluster_ids=sc["cluster_ids"], body_quats=q_all[f], body_trans=t_all[f],
                           body_centers=sc["body_centers"])
    base = (sc["means"], sc["quats"], sc["scales"], sc["opacities"])
    out = {}
    with torch.no_grad():
        out["c2_rasterization_api_ms"] = timed(lambda i=0: rs.rasterization(
            *base, sc["colors"], sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=False, **rigid(60 + i)))
        out["c2_rasterization_api_packed_ms"] = timed(lambda i=0: rs.rasterization(
            *base, sc["colors"], sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=True, **rigid(60 + i)))
    g = torch.Generator(device=dev).manual_seed(42)
    feats = torch.randn(sc["means"].shape[0], 16, device=dev, generator=g).requires_grad_()
    w = torch.rand(1, HEIGHT, WIDTH, 16, device=dev, generator=g)
    leaves = [t.clone().requires_grad_() for t in base]

    def c3_step(i=0):
        for t in leaves + [feats]:
            t.grad = None
        img, _, _ = rs.rasterization(*leaves, feats, sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=False, **rigid(60 + i))
        (img * w).sum().backward()

    out["c3_fwd_bwd_16ch_ms"] = timed(c3_step)
    # instance mask: one box per domino in screen space is not available here; a 6 x 4 grid of instances stands in
    mask = torch.zeros(HEIGHT, WIDTH, dtype=torch.long, device=dev)
    for a in range(4):
        for b in range(6):
            mask[20 + a * 260:20 + a * 260 + 240, 20 + b * 315:20 + b * 315 + 290] = 1 + a * 6 + b
    tables = rs.cluster_tables(mask, 30)
    fmap = torch.randn(HEIGHT, WIDTH, 16, device=dev, generator=g).requires_grad_()

    def cgc_step(i=0):
        fmap.grad = None
        rs.cgc_contrastive_clustering_loss(fmap, mask, tables=tables).backward()

    out["c3_contrastive_loss_fwd_bwd_ms"] = timed(cgc_step)
    return out


def ours_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if dist is not None:
        dist.barrier()
    rs = importlib.import_module("3dgs_rigidbody_b200")
    lib = importlib.import_module("3dgs_rigidbody_b200._lib").load()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak_gbs, peak_src = 6650.0, "B200_PROFILING.md fallback"

    sc_np = make_domino_scene_np()
    sc = {k: torch.from_numpy(v).to(dev) for k, v in sc_np.items()}
    q_np, t_np = domino_poses_np(N_BODIES, None, sc_np["body_centers"])
    q_all, t_all = torch.from_numpy(q_np).to(dev), torch.from_numpy(t_np).to(dev)  # [240,K,4], [240,K,3]
    fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], WIDTH, HEIGHT,
                          cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=args.max_isects)
    frames = frames_of_rank(rank, world, args.warmup + args.steps)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------------------
    pipe = rs.FramePipeline(args.in_flight, sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], WIDTH,
                            HEIGHT, cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"],
                            max_isects=args.max_isects, split=not args.no_split, rgb8=True)
    for f in frames[:args.warmup]:
        fr.render(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
        pipe.submit(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
    pipe.join()
    torch.cuda.synchronize()
    assert not fr.overflowed() and not pipe.overflowed(), "max_isects too small for the benchmark scene"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = lib.rs_launch_count() if hasattr(lib, "rs_launch_count") else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for f in frames[args.warmup:]:
        pipe.submit(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
    pipe.join()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = (lib.rs_launch_count() - n0) if n0 is not None else None
    assert not pipe.overflowed()
    n_isects_last = pipe.renderers[(pipe.count - 1) % args.in_flight].n_isects()

    # ---- end to end: host inputs -> C ABI -> host results, `in_flight` frames pipelined ---------------------------------
    K = N_BODIES
    depth = args.in_flight
    h_pose = [torch.empty(K * 7 + 16 + 9, dtype=torch.float32).pin_memory() for _ in range(depth)]
    d_pose = [torch.empty(K * 7 + 16 + 9, dtype=torch.float32, device=dev) for _ in range(depth)]
    h_img = [torch.empty(HEIGHT, WIDTH, 3, dtype=torch.float32).pin_memory() for _ in range(depth)]
    h_img8 = [torch.empty(HEIGHT, WIDTH, 3, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    vm_np, Ks_np = sc_np["viewmats"].reshape(-1), sc_np["Ks"].reshape(-1)
    slot_done = [None] * depth
    h2d_bytes = (K * 7 + 25) * 4
    d2h_bytes = HEIGHT * WIDTH * 3 * 4  # the rendered image; the reference's loop discards the alphas (main.py:387-400)

    def e2e_frame(i, f, as_rgb8):
        k = i % depth
        if slot_done[k] is not None:
            slot_done[k].synchronize()  # the host buffers of this slot hold a finished frame: "consume" it, then reuse
        # host side of the step: this frame's poses + camera, packed into one pinned buffer
        hp = h_pose[k].numpy()
        hp[:K * 4] = q_np[f].reshape(-1)
        hp[K * 4:K * 7] = t_np[f].reshape(-1)
        hp[K * 7:K * 7 + 16] = vm_np
        hp[K * 7 + 16:] = Ks_np
        with torch.cuda.stream(pipe.streams[k]):
            d_pose[k].copy_(h_pose[k], non_blocking=True)
            bq = d_pose[k][:K * 4].view(K, 4)
            bt = d_pose[k][K * 4:K * 7].view(K, 3)
            vm = d_pose[k][K * 7:K * 7 + 16].view(1, 4, 4)
            Ks = d_pose[k][K * 7 + 16:].view(1, 3, 3)
            img, alpha = pipe.renderers[k].render(vm, Ks, bq, bt)
            if as_rgb8:
                h_img8[k].copy_(pipe.renderers[k].render_rgb8[0], non_blocking=True)
            else:
                h_img[k].copy_(img[0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(pipe.streams[k])
            slot_done[k] = ev

    def e2e_run(as_rgb8):
        for i, f in enumerate(frames[:args.warmup]):
            e2e_frame(i, f, as_rgb8)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i, f in enumerate(frames[args.warmup:]):
            e2e_frame(i, f, as_rgb8)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_s = e2e_run(False)
    e2e_checksum = float(h_img[(args.steps - 1) % depth].sum())
    # the same loop reading back the 8-bit frame the reference's animation loop would store (main.py:140-171
    # save_rendered_image -> torchvision save_image quantisation), produced by the compositing epilogue: 4x fewer PCIe bytes
    e2e8_s = e2e_run(True)
    last8 = h_img8[(args.steps - 1) % depth]
    ref8 = (h_img[(args.steps - 1) % depth] * 255).add_(0.5).clamp_(0, 255).to(torch.uint8)
    rgb8_matches = bool(torch.equal(last8, ref8))

    t_ms = torch.tensor([ms, e2e_s * 1e3, e2e8_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e8_ms_max = float(t_ms[0]), float(t_ms[1]), float(t_ms[2])
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    total_frames = args.steps * world
    value = total_frames / (ms_max * 1e-3)
    line = {
        "metric": "frames/sec (1M Gaussians, 1080p, rigid-animated)", "value": round(value, 2), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "gaussians": N_GAUSS, "bodies": N_BODIES, "width": WIDTH, "height": HEIGHT,
                   "channels": 3, "n_isects_last_frame": n_isects_last, "frames_in_flight_per_gpu": args.in_flight, "sharding": f"frames round-robin over {world} rank(s), no collective",
                   "l2": "per-frame working set (Gaussians 60 MB + projected 36 MB + 2x(keys+values) >= 200 MB + images 41 MB) "
                         "exceeds the 126 MB L2 and every frame has new poses; no explicit flush"},
        "e2e": {"value": round(total_frames / (e2e_ms_max * 1e-3), 2), "unit": "frames/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "api": "FrameRenderer.render -> rs_render_frame (C ABI); pinned host poses+camera in, pinned host float32 image out, "
                       "one stream per in-flight frame (H2D, render, D2H in stream order)", "checksum_last_image": e2e_checksum},
        "e2e_rgb8": {"value": round(total_frames / (e2e8_ms_max * 1e-3), 2), "unit": "frames/s",
                     "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": HEIGHT * WIDTH * 3,
                     "note": "same loop, the result read back as the uint8 frame the reference's loop stores (main.py:140-171: "
                             "x*255+0.5, clamp, truncate), written by the compositing epilogue; extra information, `e2e` is the "
                             "float32 read-back", "equals_quantised_float_image": rgb8_matches},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_extras:
        n_tiles = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
        stages, roofline, frame_ms, M = stage_breakdown(fr, sc, q_all, t_all, frames[args.warmup:], peak_gbs, peak_src,
                                                        n_tiles)
        line["stages"] = stages
        line["roofline"] = roofline
        line["config"]["frame_ms_with_stage_events"] = round(frame_ms, 4)
        torch.cuda.empty_cache()
        budget = args.cpu_budget
        try:
            line["other_configs"] = other_configs(rs, sc, q_all, t_all)
        except Exception as e:  # extra information only: never let it take the headline line down
            line["other_configs"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
        cpu_frames = [(60 + 37 * i) % N_FRAMES for i in range(64)]  # stops at the CPU budget below
        times, threads = run_cpu_port(sc_np, cpu_frames, budget_s=budget)
        cpu_fps = len(times) / sum(times)
        line["cpu_baseline"] = {"value": round(cpu_fps, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{len(times)} full frames (animation frames 60, 97, 134, ... stride 37) of the same "
                                          f"scene through oracle/oracle.c (OpenMP), {sum(times):.1f} s of CPU work",
                                "host_cpu_count": os.cpu_count()}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--max-isects", type=int, default=24_000_000)
    ap.add_argument("--in-flight", type=int, default=4, help="frames in flight per GPU (streams + workspaces)")
    ap.add_argument("--no-split", action="store_true",
                    help="one stream per in-flight frame instead of (high-priority binning stream, compositing stream)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip stages / roofline / cpu_baseline (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 8:
            args.steps = 8  # bounded sample: each CPU frame takes seconds
        args.warmup = min(args.warmup, 1)
        return reference_arm(args)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())
