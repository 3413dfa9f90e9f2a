#!/usr/bin/env python
"""bench.py -- frames/s of the rigid-animated render path (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): synthetic domino scene, 1 M Gaussians in 20 rigid bodies,
240-frame animation, one 1920x1080 pinhole camera.  A "step" is one frame: rigid animate -> EWA project ->
tile-intersect + radix sort -> front-to-back compositing (the commented-out loop of the reference, main.py:357-409:
apply_transform() per body + rasterization()).

  value     frames/s over all ranks, everything resident in HBM (poses [240,K,.] pre-generated on the device), K frames
            enqueued back to back through FramePipeline (`--in-flight` frames per GPU, one stream + workspace each, so the
            latency-bound binning of one frame overlaps the compositing of another), CUDA events, max over ranks.
  e2e       frames/s through the public per-frame API (FrameRenderer.render -> C ABI rs_render_frame) with HOST inputs and
            HOST results: per step the frame's poses + camera are copied from pinned host memory and the rendered frame is
            copied back to pinned host memory as the 8-bit RGB image the reference's loop stores (main.py:140-171
            save_rendered_image -> torchvision save_image quantisation, done here by the compositing epilogue); copies are
            inside the timed region.  `e2e_f32` is the same loop reading back the float32 image.  Both are reported against
            the box's pinned device->host copy ceiling measured in the same run at the same rank count (`d2h_ceiling`).
  roofline  the kernel with the largest share of the step (compositing), timed live per kernel with CUDA events behind every
            launch (rs_profile_begin/_end); `kernels` lists EVERY kernel of the frame with the bytes it actually moves.
  ref_cuda  the reference's own python + CUDA kernels (baseline/_ref + oracle/_ref, unmodified) on the same B200 and the
            same frames: apply_transform() per body + rasterization() -- "the number to beat" (SURVEY.md 8d).
  cpu_baseline  the CPU port of the reference path (oracle/, OpenMP, all host threads) on a bounded sample of frames.
  other_configs  c3 (16 identity channels fwd+bwd), the drop-in rasterization() API on c2 (incl. main.py's sh_degree=3 /
            RGB+ED call), c4 (6 M Gaussians / 500 bodies / 8 cameras at 4K, cameras sharded over the ranks) and c5
            (20 M Gaussians sharded over the ranks, projected splats exchanged over NVLink peer memory and over NCCL, with a
            sharded == single-GPU image check) -- the multi-GPU configs run at every --gpus N.

Tile lists: the frame path bins a splat only into the tiles where it can reach alpha >= 1/255 (FrameRenderer(tight_tiles=True)):
bit-identical images from ~20 % fewer intersections; `run.tile_lists` re-checks that on the last timed frame and times the
same frames with the reference's lists (`--reference-tile-lists` makes those the headline).

Multi-GPU (c2): frames shard across ranks with no collective on the data path.  Every rank renders the SAME set of K
animation frames (rotated by rank), so per-rank work does not depend on N ("weak" scaling: N x K frames in total).
`--impl reference` times the oracle port on the host cores (rank 0 only).
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
    """A 240-frame job sharded over `world` ranks: rank r owns frames start + r, start + r + world, ... (FrameRenderer
    multi-GPU mode; every frame exactly once)."""
    return [(start + rank + world * i) % N_FRAMES for i in range(count)]


BENCH_FRAME_STRIDE = 7  # the K timed frames are spread over the animation (frame 5, 12, 19, ...)


def bench_frames(rank, count, start=5):
    """The frames bench.py times: the SAME set of `count` animation frames on every rank and for every --gpus N (so the
    per-rank work is identical at 1, 2, 4 and 8 GPUs and the driver's scaling efficiency compares like with like), rotated
    by rank so that the ranks are not in lock step."""
    base = [(start + BENCH_FRAME_STRIDE * i) % N_FRAMES for i in range(count)]
    r = rank % max(count, 1)
    return base[r:] + base[:r]


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


# ----------------------------------------------------------------------------------------------------------------------
# shared by both arms: the identity of the workload (identical `config` dicts => the driver's same_config check holds)
# ----------------------------------------------------------------------------------------------------------------------
def base_config():
    return {"workload": WORKLOAD, "gaussians": N_GAUSS, "bodies": N_BODIES, "width": WIDTH, "height": HEIGHT, "channels": 3,
            "frames": f"the same K animation frames on every rank: 5, {5 + BENCH_FRAME_STRIDE}, {5 + 2 * BENCH_FRAME_STRIDE}, ... "
                      "(stride 7, modulo 240), rotated by rank",
            "l2": "per-frame working set (Gaussians 60 MB + projected 68 MB + keys/values >= 120 MB + images 41 MB) exceeds "
                  "the 126 MB L2 and every frame has new poses; no explicit flush"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc = make_domino_scene_np()
    frames = bench_frames(0, args.warmup + args.steps)
    times, threads = run_cpu_port(sc, frames)
    timed = times[args.warmup:]
    fps = len(timed) / sum(timed)
    sample = f"{len(timed)} full frames of the 240-frame animation after {args.warmup} warm-up frame(s)"
    line = {
        "impl": "reference", "metric": "frames/sec (1M Gaussians, 1080p, rigid-animated)", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(timed) / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": base_config(),
        "arm": "CPU port of the reference path (oracle/oracle.c, OpenMP, all host threads); the reference's own python CPU "
               "path cannot composite without its CUDA extension (SURVEY.md 8c)",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def _timed_ms(torch, fn, steps, warm=3):
    """ms per call of fn(i): `warm` untimed calls, then `steps` calls between two CUDA events on the current stream."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def kernel_rooflines(rs, _lib, fr, sc, q_all, t_all, frames, peak_gbs, peak_src, n_tiles):
    """Per-KERNEL times of the frames through rs_render_frame itself, measured live: one CUDA event behind every kernel
    launch (rs_profile_begin / rs_profile_end).  Every kernel gets the bytes it actually has to move
    and its achieved GB/s against the measured HBM peak.  Returns (kernels, stages, roofline, frame_ms, M)."""
    import torch

    stream = torch.cuda.current_stream().cuda_stream
    per_frame = []
    for f in frames:
        rows = _lib.profile_kernels(lambda: fr.render(sc["viewmats"], sc["Ks"], q_all[f], t_all[f]), stream)
        per_frame.append(rows)
    M = fr.n_isects()
    N, D, C = fr.N, fr.D, fr.C
    E, P = C * N, C * fr.W * fr.H
    names = [n for n, _ in per_frame[0]]
    assert all([n for n, _ in rows] == names for rows in per_frame), "kernel sequence differs between frames"
    mean_ms = np.mean([[ms for _, ms in rows] for rows in per_frame], axis=0)
    tile_bits = int(n_tiles).bit_length() + int(C).bit_length()
    # bytes each launch has to move, keyed by (kernel name, occurrence index of that name inside the frame)
    n_tile_passes = (tile_bits + 7) // 8
    V = min(E, M)  # visible elements (an upper bound is enough for the byte counts: every visible element has >= 1 tile)
    seen = {}
    kernels = []
    for name, ms in zip(names, mean_ms):
        k = seen.get(name, 0)
        seen[name] = k + 1
        what, nbytes = name, None
        if name.startswith("rs_project_fwd"):
            what, nbytes = "rigid transform + EWA projection + tile count + compositing records", N * 48 + E * (84 if fr.tight_tiles else 68)
        elif name == "rs_dord_minmax_kernel":
            what, nbytes = "depth order 1/5: min / max of the visible depth bits", E * 8
        elif name == "rs_dord_count_kernel":
            what, nbytes = "depth order 2/5: histogram over 65536 range-adapted depth buckets", E * 8 + V * 4
        elif name == "rs_dord_scan_kernel":
            what, nbytes = "depth order 3/5: scan of the bucket counts, sort-group boundaries (1 CTA)", 65536 * 8
        elif name == "rs_dord_scatter_kernel":
            what, nbytes = "depth order 4/5: scatter of (depth bits, id) composites into their buckets", E * 8 + V * 12
        elif name == "rs_dord_sort_kernel":
            what, nbytes = "depth order 5/5: per-group bitonic sort in shared memory -> ids in depth order", V * 12
        elif name == "sort_pass_kernel":
            what = f"tile-key radix pass {k + 1}/{n_tile_passes} ((image|tile, id) pairs)"
            nbytes = M * 16
        elif name == "rs_bin_count_kernel":
            what, nbytes = "tile counts gathered in depth order -> block sums", E * 8
        elif name == "rs_isect_scan_kernel":
            what, nbytes = "exclusive scan of the block sums (1 CTA)", (E // 1024 + 1) * 8
        elif name == "rs_bin_emit_kernel":
            what, nbytes = "emission of (image|tile, id) pairs in depth order + tile-key histograms", E * 24 + M * 8
        elif name == "rs_isect_offsets_kernel":
            what, nbytes = "per-tile offsets from the sorted tile keys", M * 4 + n_tiles * C * 4
        elif name == "rs_raster_fwd_kernel":
            what, nbytes = "front-to-back compositing (SURVEY 8d algorithmic bytes)", M * (4 + 28 + 4 * D) + P * (D + 2) * 4
        row = {"kernel": name, "what": what, "ms": round(float(ms), 4)}
        if nbytes is not None:
            gbs = nbytes / (float(ms) * 1e-3) / 1e9
            row.update(bytes=int(nbytes), GBps=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak_gbs, 4))
        kernels.append(row)
    total = float(mean_ms.sum())
    for row in kernels:
        row["share_of_frame"] = round(row["ms"] / total, 4)
    stage_of = lambda n: ("rigid+project" if n.startswith("rs_project") else "composite" if n.startswith("rs_raster") else "binning")
    stages = []
    for st in ("rigid+project", "binning", "composite"):
        rows = [r for r in kernels if stage_of(r["kernel"]) == st]
        ms = sum(r["ms"] for r in rows)
        nb = sum(r.get("bytes", 0) for r in rows)
        stages.append({"stage": st, "ms": round(ms, 4), "kernels": len(rows), "bytes_moved_MB": round(nb / 1e6, 1),
                       "GBps": round(nb / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(nb / (ms * 1e-3) / 1e9 / peak_gbs, 4)})
    top = max(kernels, key=lambda r: r["ms"])
    traffic, traffic_src = None, "no ncu capture of this build under profiles/ (profiles/r02_kernel_traffic.json)"
    tpath = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture
        try:
            tj = json.load(open(tpath))
            if top["kernel"] in tj.get("kernels", {}):
                traffic = tj["kernels"][top["kernel"]]["dram_bytes_per_launch"]
                traffic_src = f"profiles/r02_kernel_traffic.json ({tj.get('source', 'ncu --set full')}); static profile data, not measured in this run"
        except (OSError, ValueError, KeyError):
            pass
    roofline = {
        "bound": "hbm", "kernel": f"{top['kernel']} ({top['what']}; 1 launch/frame, {top['share_of_frame']:.0%} of the frame's kernel time)",
        "achieved": top.get("GBps"), "peak": peak_gbs, "unit": "GB/s", "frac": top.get("frac_of_hbm_peak"),
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": top.get("bytes"), "ms_per_launch": top["ms"], "frames_timed": int(len(frames)),
        "timing": "CUDA events behind every kernel launch on the launching stream (rs_profile_begin/_end), mean over the timed frames",
        "limiter": "SM issue (FP32 FMA + MUFU.EX2 + LDS) and per-tile latency, not HBM: the roofline fraction of this kernel says how "
                   "far it is from being bandwidth-bound; the HBM-bound kernels of the frame are listed in `kernels`",
    }
    return kernels, stages, roofline, total, M


def d2h_ceiling(torch, dist, dev, nbytes=64 << 20, reps=8):
    """Aggregate pinned device->host copy bandwidth of this box with all ranks copying at once (GB/s): the roof of `e2e`."""
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    world = dist.get_world_size() if dist is not None else 1
    return world * nbytes * reps / (float(t[0]) * 1e-3) / 1e9


def ref_cuda_leg(sc, q_all, t_all, frames, steps=10):
    """The reference's OWN python and CUDA kernels on this GPU and these frames ("that, not the CPU, is the
    number to beat"): per frame, apply_transform() (main.py:183-228) once per body on that body's Gaussians, then
    gsplat.rendering.rasterization(packed=False) (rendering.py:33-770) on its compiled extension.  Unmodified reference
    code from baseline/_ref + oracle/_ref; measurement infrastructure only."""
    import importlib.util

    import torch

    so = os.path.join(ROOT, "oracle", "_ref", "gsplat_ref_cuda.so")
    inst = os.path.join(ROOT, "baseline", "install_ref.py")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/gsplat_ref_cuda.so not built (needs /root/reference at build time)"}
    spec = importlib.util.spec_from_file_location("install_ref", inst)
    refpy = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refpy)
    if not refpy.available():
        return {"unavailable": "baseline/_ref not installed (needs /root/reference at build time)"}
    spec = importlib.util.spec_from_file_location("gsplat_ref_cuda", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    gs = refpy.load_reference(ref)
    fns = refpy.reference_functions("main.py", ["apply_transform", "quat_multiply"])
    K = q_all.shape[1]
    body_idx = [torch.nonzero(sc["cluster_ids"] == k).squeeze(-1) for k in range(K)]
    splats = {"means": sc["means"], "quats": sc["quats"]}

    def animate(f):
        means, quats = sc["means"].clone(), sc["quats"].clone()
        for k in range(K):
            part = {n: v[body_idx[k]] for n, v in splats.items()}
            moved = fns["apply_transform"](part, t_all[f, k], q_all[f, k])
            means[body_idx[k]] = moved["means"]
            quats[body_idx[k]] = moved["quats"]
        return means, quats

    def render(m, q):
        return gs.rendering.rasterization(m, q, sc["scales"], sc["opacities"], sc["colors"], sc["viewmats"], sc["Ks"], WIDTH,
                                          HEIGHT, packed=False)

    fl = list(frames)
    with torch.no_grad():
        t_anim = _timed_ms(torch, lambda i: animate(fl[i % len(fl)]), steps)
        m, q = animate(fl[0])
        t_render = _timed_ms(torch, lambda i: render(m, q), steps)
        t_frame = _timed_ms(torch, lambda i: render(*animate(fl[i % len(fl)])), steps)
        # per operator, on the moved splats of one frame
        C_ = ref
        tw, th = (WIDTH + 15) // 16, (HEIGHT + 15) // 16
        pa = (m, None, q, sc["scales"], sc["opacities"], sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, 0.3, 0.01, 1e10, 0.0, False)
        radii, means2d, depths, conics, _ = C_.projection_ewa_3dgs_fused_fwd(*pa, C_.PINHOLE)
        tpg, ids, flat = C_.intersect_tile(means2d, radii, depths, None, None, 1, 16, tw, th, True, False)
        off = C_.intersect_offset(ids, 1, tw, th)
        colors, opac = sc["colors"][None].contiguous(), sc["opacities"][None].contiguous()
        ops = {
            "projection_ewa_3dgs_fused_fwd": _timed_ms(torch, lambda i: C_.projection_ewa_3dgs_fused_fwd(*pa, C_.PINHOLE), steps),
            "intersect_tile (count + cumsum + emit + cub 64-bit sort, 1 host sync)": _timed_ms(
                torch, lambda i: C_.intersect_tile(means2d, radii, depths, None, None, 1, 16, tw, th, True, False), steps),
            "intersect_offset": _timed_ms(torch, lambda i: C_.intersect_offset(ids, 1, tw, th), steps),
            "rasterize_to_pixels_3dgs_fwd": _timed_ms(torch, lambda i: C_.rasterize_to_pixels_3dgs_fwd(
                means2d, conics, colors, opac, None, None, WIDTH, HEIGHT, 16, off, flat), steps),
        }
    return {
        "what": "UNMODIFIED reference python (baseline/_ref: main.py apply_transform per body + gsplat.rendering.rasterization, "
                "packed=False) on the reference's own CUDA extension (oracle/_ref), same GPU, same frames; ms per frame, CUDA events",
        "frame_ms": round(t_frame, 4), "frames_per_s": round(1e3 / t_frame, 2),
        "apply_transform_per_body_ms": round(t_anim, 4), "rasterization_ms": round(t_render, 4),
        "operators_ms": {k: round(v, 4) for k, v in ops.items()}, "kernels_only_ms": round(sum(ops.values()), 4),
        "n_isects": int(ids.numel()), "steps": steps,
    }


def c2_api_configs(rs, sc, q_all, t_all, frames, steps=10):
    """The c2 scene through the drop-in `rasterization()` signature (operator path, with the host reads the reference API
    implies), including main.py:328-339's own configuration (sh_degree=3, render_mode="RGB+ED", packed=False), and
    BASELINE configs[2] (c3): 16 identity-feature channels, forward + backward, plus the contrastive clustering loss and the
    segmentation-head MLP of the identity step.  CUDA events, ms per step."""
    import torch

    dev = sc["means"].device
    fl = list(frames)
    rigid = lambda i: dict(cluster_ids=sc["cluster_ids"], body_quats=q_all[fl[i % len(fl)]], body_trans=t_all[fl[i % len(fl)]],
                           body_centers=sc["body_centers"])
    base = (sc["means"], sc["quats"], sc["scales"], sc["opacities"])
    out = {}
    with torch.no_grad():
        out["c2_rasterization_api_ms"] = round(_timed_ms(torch, lambda i: rs.rasterization(
            *base, sc["colors"], sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=False, **rigid(i)), steps), 4)
        out["c2_rasterization_api_packed_ms"] = round(_timed_ms(torch, lambda i: rs.rasterization(
            *base, sc["colors"], sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=True, **rigid(i)), steps), 4)
        g = torch.Generator(device=dev).manual_seed(7)
        sh = torch.randn(sc["means"].shape[0], 16, 3, device=dev, generator=g) * 0.2
        out["c2_main_py_call_sh3_rgb_ed_ms"] = round(_timed_ms(torch, lambda i: rs.rasterization(
            *base, sh, sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=False, sh_degree=3, render_mode="RGB+ED",
            near_plane=0.01, far_plane=1e10, **rigid(i)), steps), 4)
        del sh
    g = torch.Generator(device=dev).manual_seed(42)
    feats = torch.randn(sc["means"].shape[0], 16, device=dev, generator=g).requires_grad_()
    w = torch.rand(1, HEIGHT, WIDTH, 16, device=dev, generator=g)
    leaves = [t.clone().requires_grad_() for t in base]

    def c3_step(i=0):
        for t in leaves + [feats]:
            t.grad = None
        img, _, _ = rs.rasterization(*leaves, feats, sc["viewmats"], sc["Ks"], WIDTH, HEIGHT, packed=False, **rigid(i))
        (img * w).sum().backward()

    out["c3_fwd_bwd_16ch_ms"] = round(_timed_ms(torch, c3_step, steps), 4)
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

    out["c3_contrastive_loss_fwd_bwd_ms"] = round(_timed_ms(torch, cgc_step, steps), 4)
    # segmentation head on all N identity encodings (examples/simple_trainer.py:946-947): fused kernels vs the reference's
    # nn.Sequential (library GEMMs around a materialised [N, 64] hidden tensor)
    torch.manual_seed(0)
    ref_head = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 16)).to(dev)
    head = rs.SegmentationHead(16, 64).to(dev)
    head.load_state_dict(ref_head.state_dict())
    enc = torch.randn(sc["means"].shape[0], 16, device=dev, generator=g).requires_grad_()
    v_enc = torch.randn(sc["means"].shape[0], 16, device=dev, generator=g)

    def head_step(module):
        def fn(i=0):
            enc.grad = None
            for p in module.parameters():
                p.grad = None
            module(enc).backward(v_enc)
        return fn

    out["c3_segmentation_head_fwd_bwd_ms"] = round(_timed_ms(torch, head_step(head), steps), 4)
    out["c3_segmentation_head_fwd_bwd_torch_nn_sequential_ms"] = round(_timed_ms(torch, head_step(ref_head), steps), 4)
    return out


def bench_c4(rs, torch, dist, dev, rank, world, frames=6, warmup=1, in_flight=3, tight_tiles=True):
    """c4 (BASELINE configs[3]): 6 M Gaussians, 500 rigid bodies, 8 ring cameras at 3840x2160.  The 8 cameras of every
    animation frame are sharded over the ranks (camera c of frame f -> rank (c + f) % world: every rank sees every viewpoint,
    so the busiest one does not pin a rank), Gaussians replicated, NO collective on the data path.  camera-frames/s over all
    ranks (CUDA events, max over ranks); at world 1 one GPU renders all 8 cameras."""
    W, H, N, K, C = 3840, 2160, 6_000_000, 500, 8
    sc = make_domino_scene(N, K, device=dev, width=W, height=H, n_cameras=C)
    q_np, t_np = domino_poses_np(K, None, sc["body_centers"].cpu().numpy())
    q_all, t_all = torch.from_numpy(q_np).to(dev), torch.from_numpy(t_np).to(dev)
    cams_of = lambda f: [c for c in range(C) if (c + f) % world == rank]
    pipe = rs.FramePipeline(in_flight, sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                            cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=96_000_000,
                            tight_tiles=tight_tiles)

    def run(fs):
        for f in fs:
            for c in cams_of(f):
                pipe.submit(sc["viewmats"][c:c + 1], sc["Ks"][c:c + 1], q_all[f % N_FRAMES], t_all[f % N_FRAMES])
        pipe.join()

    run(range(60, 60 + warmup))
    torch.cuda.synchronize()
    ok = not pipe.overflowed()
    # replicas must agree: every rank renders camera 0 of frame 60 and the checksums are compared
    img, _, done = pipe.submit(sc["viewmats"][0:1], sc["Ks"][0:1], q_all[60], t_all[60])
    pipe.join()
    torch.cuda.synchronize()
    chk = torch.stack([img.double().sum(), img.double().abs().max()])
    if dist is not None:
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        replicas_equal = all(bool(torch.equal(allc[0], c)) for c in allc)
        dist.barrier()
    else:
        replicas_equal = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(range(60 + warmup, 60 + warmup + frames))
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    n_is = torch.tensor([float(pipe.renderers[0].n_isects())], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_is, op=dist.ReduceOp.SUM)
    ms = float(t[0])
    del pipe, sc
    torch.cuda.empty_cache()
    return {"workload": "c4: 6M Gaussians, 500 bodies, 8 ring cameras 3840x2160; the cameras of each animation frame sharded over the ranks, no collective" + ("; tight tile lists" if tight_tiles else "; the reference's tile lists"),
            "n_gpus": world, "animation_frames": frames, "camera_frames": frames * C,
            "ms_per_animation_frame": round(ms / frames, 3), "camera_frames_per_s": round(frames * C / (ms * 1e-3), 2),
            "n_isects_sample_camera_mean": int(float(n_is[0]) / world), "workspace_ok": bool(ok),
            "replica_checksums_equal_across_ranks": bool(replicas_equal)}


def _c5_scene(torch, dev, n_total, lo, hi, seed=1234):
    """Gaussians [lo, hi) of the c5 scene: uniform in a 40 x 40 x 4 slab.  Generated in fixed blocks of 1 M
    with a per-block seed, so any shard of any world size sees exactly the values of the single-GPU scene."""
    BLOCK = 1_000_000
    parts = {k: [] for k in ("means", "quats", "scales", "opac", "colors")}
    for b in range(lo // BLOCK, (max(hi, lo + 1) - 1) // BLOCK + 1):
        g = torch.Generator(device=dev).manual_seed(seed + b)
        n = min(BLOCK, n_total - b * BLOCK)
        blk = {
            "means": (torch.rand(n, 3, device=dev, generator=g) - 0.5) * torch.tensor([40.0, 40.0, 4.0], device=dev),
            "quats": torch.nn.functional.normalize(torch.randn(n, 4, device=dev, generator=g), dim=-1),
            "scales": torch.rand(n, 3, device=dev, generator=g) * 0.02,
            "opac": torch.rand(n, device=dev, generator=g),
            "colors": torch.rand(n, 3, device=dev, generator=g),
        }
        a, z = max(lo - b * BLOCK, 0), min(hi - b * BLOCK, n)
        for k in parts:
            parts[k].append(blk[k][a:z])
    return {k: torch.cat(v).contiguous() for k, v in parts.items()}


def _c5_cameras(torch, dev, n_cams, W, H):
    vms = np.stack([look_at((30 * math.cos(2 * math.pi * c / n_cams), 30 * math.sin(2 * math.pi * c / n_cams), 12.0), (0, 0, 0))
                    for c in range(n_cams)])
    f = 0.5 * W / math.tan(math.radians(30.0))
    Ks = np.tile(np.array([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]], np.float32), (n_cams, 1, 1))
    return torch.from_numpy(vms).to(dev), torch.from_numpy(Ks).to(dev)


class _Skip(Exception):
    pass


def bench_c5(rs, torch, dist, dev, rank, world, n_total=20_000_000, steps=5, warmup=2, n_cams=8):
    """c5 (BASELINE configs[4]): 20 M Gaussians sharded over the ranks (contiguous blocks), 8 ring cameras at 1080p owned
    8 / world per rank; every rank projects ITS Gaussians to ALL cameras and the projected splats travel to the rank owning the
    camera (rendering.py:527-611) -- over NVLink peer memory (rs_exchange_push, one kernel per rank) and, for comparison,
    over the NCCL all-to-all route.  Total work is fixed (strong scaling).  Includes a correctness flag: a reduced scene
    (400 k Gaussians, the same 8 cameras at 480x270) rendered sharded on all ranks == rendered whole on one GPU."""
    dmod = importlib.import_module("3dgs_rigidbody_b200.distributed")
    W, H = WIDTH, HEIGHT
    assert n_cams % world == 0
    cl = n_cams // world
    mine = slice(rank * cl, (rank + 1) * cl)
    out = {"workload": f"c5: {n_total // 1_000_000}M Gaussians sharded over the ranks, {n_cams} ring cameras 1920x1080 ({cl} per rank), "
                       "projected splats exchanged to the camera owners", "n_gpus": world}

    def render(sc, vm, Ks, w, h):
        with torch.no_grad():
            return rs.rasterization(sc["means"], sc["quats"], sc["scales"], sc["opac"], sc["colors"], vm, Ks, w, h, packed=True,
                                    distributed=dist is not None)

    # ---- correctness: sharded == single GPU on a reduced scene ---------------------------------------------------------
    n_small, ws, hs = 400_000, 480, 270
    vm_s, Ks_s = _c5_cameras(torch, dev, n_cams, ws, hs)
    lo, hi = rank * n_small // world, (rank + 1) * n_small // world
    small = _c5_scene(torch, dev, n_small, lo, hi, seed=99)
    for route in (("peer", "nccl") if dist is not None else ("single",)):
        dmod.PeerSplatExchange.enabled = route == "peer"
        dmod.PeerSplatExchange._usable.clear()
        img_sh, alpha_sh, _ = render(small, vm_s[mine], Ks_s[mine], ws, hs)
        if dist is not None:
            parts = [torch.empty_like(img_sh) for _ in range(world)]
            dist.all_gather(parts, img_sh.contiguous())
            img_all = torch.cat(parts, 0)
        else:
            img_all = img_sh
        if rank == 0:
            whole = _c5_scene(torch, dev, n_small, 0, n_small, seed=99)
            with torch.no_grad():
                img_one, _, _ = rs.rasterization(whole["means"], whole["quats"], whole["scales"], whole["opac"], whole["colors"],
                                                 vm_s, Ks_s, ws, hs, packed=True)
            out[f"sharded_equals_single_gpu_{route}"] = {
                "checksum_equal_to_single_gpu": bool(torch.equal(img_all, img_one)),
                "max_abs_diff": float((img_all - img_one).abs().max()), "mean_alpha_gt_0": bool(float(img_one.mean()) > 0),
                "scene": f"{n_small} Gaussians, {n_cams} cameras {ws}x{hs}"}
            del whole, img_one
    del small
    torch.cuda.empty_cache()

    # ---- timing at full size ----------------------------------------------------------------------------------------------
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    sc = _c5_scene(torch, dev, n_total, lo, hi)
    vm, Ks = _c5_cameras(torch, dev, n_cams, W, H)
    for route in (("peer", "nccl") if dist is not None else ("single",)):
        dmod.PeerSplatExchange.enabled = route == "peer"
        dmod.PeerSplatExchange._usable.clear()
        for _ in range(warmup):
            img, alpha, meta = render(sc, vm[mine], Ks[mine], W, H)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            img, alpha, meta = render(sc, vm[mine], Ks[mine], W, H)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        st = torch.tensor([float(meta["flatten_ids"].numel()), float(meta["gaussian_ids"].numel())], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(st, op=dist.ReduceOp.SUM)
        ms = float(t[0]) / steps
        out[f"{route}_exchange"] = {"ms_per_step": round(ms, 3), "camera_frames_per_s": round(n_cams / (ms * 1e-3), 2),
                                    "n_isects_total": int(st[0]), "visible_rows_total": int(st[1])}
    dmod.PeerSplatExchange.enabled = True
    dmod.PeerSplatExchange._usable.clear()
    # ---- the same frame without any host read on the way (distributed.ShardedFrameRenderer) ----------------------------
    try:
        if dist is None:
            raise _Skip("needs a torch.distributed process group (run with --gpus N >= 2)")
        fr = dmod.ShardedFrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opac"], sc["colors"], W, H, cl,
                                       tight_tiles=True)
        for _ in range(warmup):
            img_f, alpha_f = fr.render(vm[mine], Ks[mine])
        info = fr.check()
        same = bool(torch.equal(img_f, img)) if dist is not None else None  # `img`: the last frame of the route above
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fr.render(vm[mine], Ks[mine])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0]) / steps
        out["peer_exchange_sync_free"] = {"ms_per_step": round(ms, 3), "camera_frames_per_s": round(n_cams / (ms * 1e-3), 2),
                                          "host_reads_per_frame": 0, "tile_lists": "tight", "rows_received_rank0": info["rows"],
                                          "n_isects_rank0": info["n_isects"], "regrow": info["regrow"],
                                          "image_equal_to_rasterization_route": same}
        del fr
    except (_Skip, dmod.PeerRouteUnavailable) as e:  # (the second is raised on every rank alike: nobody is stranded)
        out["peer_exchange_sync_free"] = {"skipped": str(e)}
    except Exception as e:
        out["peer_exchange_sync_free"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if dist is not None:
            raise
    del sc
    torch.cuda.empty_cache()
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
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    lib = _lib.load()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak_gbs, peak_src = 6650.0, "B200_PROFILING.md fallback"

    sc_np = make_domino_scene_np()
    sc = {k: torch.from_numpy(v).to(dev) for k, v in sc_np.items()}
    q_np, t_np = domino_poses_np(N_BODIES, None, sc_np["body_centers"])
    q_all, t_all = torch.from_numpy(q_np).to(dev), torch.from_numpy(t_np).to(dev)  # [240,K,4], [240,K,3]
    tight = not args.reference_tile_lists
    fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], WIDTH, HEIGHT,
                          cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=args.max_isects,
                          tight_tiles=tight)
    frames = bench_frames(rank, args.warmup + args.steps)
    warm_frames, timed_frames = frames[:args.warmup], frames[args.warmup:]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------------------
    pipe = rs.FramePipeline(args.in_flight, sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], WIDTH,
                            HEIGHT, cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"],
                            max_isects=args.max_isects, split=not args.no_split, rgb8=True, tight_tiles=tight)
    for f in warm_frames:
        fr.render(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
        pipe.submit(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
    pipe.join()
    torch.cuda.synchronize()
    assert not fr.overflowed() and not pipe.overflowed(), "max_isects too small for the benchmark scene"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = lib.rs_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for f in timed_frames:
        pipe.submit(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
    pipe.join()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = lib.rs_launch_count() - n0
    assert not pipe.overflowed()
    n_isects_last = pipe.renderers[(pipe.count - 1) % args.in_flight].n_isects()
    # tight tile lists: the last timed frame once more from the reference's lists -- same pixels, bit for bit?
    tile_lists = {"mode": "tight" if tight else "reference"}
    if tight:
        last = pipe.renderers[(pipe.count - 1) % args.in_flight]
        plain = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], WIDTH, HEIGHT,
                                 cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=args.max_isects,
                                 rgb8=True)
        f_last = timed_frames[-1]
        img_p, alpha_p = plain.render(sc["viewmats"], sc["Ks"], q_all[f_last], t_all[f_last])
        torch.cuda.synchronize()
        tile_lists.update(
            n_isects_tight=n_isects_last, n_isects_reference_lists=plain.n_isects(),
            image_bit_identical_to_reference_lists=bool(torch.equal(img_p, last.render_colors) and
                                                        torch.equal(alpha_p, last.render_alphas) and
                                                        torch.equal(plain.render_rgb8, last.render_rgb8)),
            what="a (tile, splat) pair is listed only where the splat can reach alpha >= 1/255 at a pixel centre of the tile; "
                 "the pairs dropped are skipped at every pixel by RasterizeToPixels3DGSFwd.cu:148-149")
        del plain
        if not args.no_extras and world == 1:  # the same timed frames from the reference's lists, for comparison
            pipe_ref = rs.FramePipeline(args.in_flight, sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"],
                                        WIDTH, HEIGHT, cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"],
                                        max_isects=args.max_isects, split=not args.no_split, rgb8=True)
            for f in warm_frames:
                pipe_ref.submit(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
            pipe_ref.join()
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for f in timed_frames:
                pipe_ref.submit(sc["viewmats"], sc["Ks"], q_all[f], t_all[f])
            pipe_ref.join()
            r1.record()
            torch.cuda.synchronize()
            tile_lists["value_with_reference_lists"] = round(args.steps / (r0.elapsed_time(r1) * 1e-3), 2)
            del pipe_ref
        torch.cuda.empty_cache()

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
        for i, f in enumerate(warm_frames):
            e2e_frame(i, f, as_rgb8)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i, f in enumerate(timed_frames):
            e2e_frame(i, f, as_rgb8)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e32_s = e2e_run(False)
    e2e_checksum = float(h_img[(args.steps - 1) % depth].sum())
    e2e8_s = e2e_run(True)
    last8 = h_img8[(args.steps - 1) % depth]
    ref8 = (h_img[(args.steps - 1) % depth] * 255).add_(0.5).clamp_(0, 255).to(torch.uint8)
    rgb8_matches = bool(torch.equal(last8, ref8))
    ceiling_gbs = d2h_ceiling(torch, dist, dev)

    t_ms = torch.tensor([ms, e2e32_s * 1e3, e2e8_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e32_ms_max, e2e8_ms_max = float(t_ms[0]), float(t_ms[1]), float(t_ms[2])
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    total_frames = args.steps * world
    value = total_frames / (ms_max * 1e-3)
    e2e8_fps, e2e32_fps = total_frames / (e2e8_ms_max * 1e-3), total_frames / (e2e32_ms_max * 1e-3)
    d2h8, d2h32 = HEIGHT * WIDTH * 3, HEIGHT * WIDTH * 3 * 4
    api = ("FrameRenderer.render -> rs_render_frame (C ABI); pinned host poses+camera in, pinned host frame out, one stream per "
           "in-flight frame (H2D, render, D2H in stream order)")
    line = {
        "metric": "frames/sec (1M Gaussians, 1080p, rigid-animated)", "value": round(value, 2), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(),
        "run": {"n_isects_last_frame": n_isects_last, "tile_lists": tile_lists, "frames_in_flight_per_gpu": args.in_flight,
                "sharding": f"{world} rank(s) x the same {args.steps} frames, no collective on the data path"},
        "e2e": {"value": round(e2e8_fps, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h8,
                "api": api, "result": "uint8 [H,W,3] frame, quantised as the reference's loop stores it (main.py:140-171 "
                                      "save_rendered_image -> torchvision save_image: x*255+0.5, clamp, truncate) by the compositing epilogue",
                "equals_quantised_float_image": rgb8_matches,
                "d2h_GBps": round(e2e8_fps * d2h8 / 1e9, 2), "d2h_ceiling_GBps": round(ceiling_gbs, 2),
                "frac_of_d2h_ceiling": round(e2e8_fps * d2h8 / 1e9 / ceiling_gbs, 4)},
        "e2e_f32": {"value": round(e2e32_fps, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h32,
                    "result": "float32 [H,W,3] image", "checksum_last_image": e2e_checksum,
                    "d2h_GBps": round(e2e32_fps * d2h32 / 1e9, 2), "d2h_ceiling_GBps": round(ceiling_gbs, 2),
                    "frac_of_d2h_ceiling": round(e2e32_fps * d2h32 / 1e9 / ceiling_gbs, 4)},
        "d2h_ceiling": {"GBps": round(ceiling_gbs, 2), "how": f"{world} rank(s) copying 64 MiB device -> pinned host at once, 8 copies each, "
                                                              "CUDA events, max over ranks, in this run"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    extras = not args.no_extras
    if extras and world == 1 and rank == 0:
        n_tiles = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
        kernels, stages, roofline, frame_ms, M = kernel_rooflines(rs, _lib, fr, sc, q_all, t_all, timed_frames, peak_gbs,
                                                                  peak_src, n_tiles)
        line["kernels"] = kernels
        line["stages"] = stages
        line["roofline"] = roofline
        line["run"]["frame_ms_one_frame_in_flight_with_kernel_events"] = round(frame_ms, 4)
    del pipe
    torch.cuda.empty_cache()
    other = {}
    if extras:
        if world == 1:
            for name, fn in (("api", lambda: c2_api_configs(rs, sc, q_all, t_all, timed_frames)),
                             ("ref_cuda", lambda: ref_cuda_leg(sc, q_all, t_all, timed_frames))):
                try:
                    res = fn()
                except Exception as e:  # extra information only: never let it take the headline line down
                    res = {"error": f"{type(e).__name__}: {e}"[:300]}
                if name == "api":
                    other.update(res)
                else:
                    line["ref_cuda"] = res
                torch.cuda.empty_cache()
        del fr
        sc = None
        torch.cuda.empty_cache()
        for name, fn in (("c4", lambda: bench_c4(rs, torch, dist, dev, rank, world)),
                         ("c5", lambda: bench_c5(rs, torch, dist, dev, rank, world))):
            if name in args.skip.split(","):
                continue
            try:
                other[name] = fn()
            except Exception as e:
                other[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                if dist is not None:  # a rank that failed alone would strand the others inside a collective
                    raise
            torch.cuda.empty_cache()
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    if other:
        line["other_configs"] = other
    if extras and world == 1:
        cpu_frames = [(60 + 37 * i) % N_FRAMES for i in range(64)]  # stops at the CPU budget below
        times, threads = run_cpu_port(sc_np, cpu_frames, budget_s=args.cpu_budget)
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
    ap.add_argument("--in-flight", type=int, default=6, help="frames in flight per GPU (streams + workspaces)")
    ap.add_argument("--no-split", action="store_true",
                    help="one stream per in-flight frame instead of (high-priority binning stream, compositing stream)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip kernels / roofline / other_configs / cpu_baseline")
    ap.add_argument("--reference-tile-lists", action="store_true",
                    help="bin every tile of a splat's bounding rectangle (exactly the reference's lists) instead of the tight lists")
    ap.add_argument("--skip", default="", help="comma list of other_configs to skip: c4,c5")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())
