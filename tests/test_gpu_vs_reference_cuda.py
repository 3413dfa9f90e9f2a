"""GPU parity against the REFERENCE'S OWN CUDA kernels (oracle/_ref/gsplat_ref_cuda.so: the reference extension compiled
from the sources where they lie under /root/reference/gsplat/cuda by oracle/build_ref.py; test infrastructure only).

Both `_C` modules take the same positional arguments (gsplat/cuda/ext.cpp:6-104), so every test calls the two with the
same tensors.  Bars (BASELINE.json north_star):
  * tile counts, isect ids (keys), flatten ids (sorted order), offsets: bit-exact on identical stage inputs;
  * projection: radii exact; means2d / depths / conics <= 1e-5 relative (both sides are fast-math float32; the operation
    order differs);
  * compositing forward: images / alphas max-abs <= 1e-4 (they are in fact expected to be bit-identical: the kernel
    reproduces the reference's FMUL/FFMA association) and last_ids exact;
  * gradients: <= 2e-3 of the tensor's max magnitude (float atomics in a different order).
Also pins the CPU oracle (oracle/oracle.c) against the reference CUDA path on the same inputs.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, apply_transform_per_body, pinhole_cameras, synthetic_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def rel_err(got, want):
    scale = max(float(want.abs().max()), 1e-12)
    return float((got - want).abs().max()) / scale


def project_both(rs, ref, s, vm, Ks, W, H, comp=False, model=None, radius_clip=0.0):
    args = (T(s["means"]), None, T(s["quats"]), T(s["scales"]), T(s["opacities"]), T(vm), T(Ks), W, H, 0.3, 0.01, 1e10,
            radius_clip, comp)
    ours = rs._C.projection_ewa_3dgs_fused_fwd(*args, rs._C.PINHOLE if model is None else getattr(rs._C, model))
    theirs = ref.projection_ewa_3dgs_fused_fwd(*args, ref.PINHOLE if model is None else getattr(ref, model))
    return ours, theirs


@pytest.mark.parametrize("C,comp", [(1, False), (3, True)])
def test_projection_fwd_matches_reference_cuda(rs, ref, C, comp):
    W, H = 320, 200
    s = synthetic_scene(11, 50_000, s_max=0.08)
    vm, Ks = pinhole_cameras(C, W, H)
    ours, theirs = project_both(rs, ref, s, vm, Ks, W, H, comp=comp)
    r_o, r_t = ours[0], theirs[0]
    vis_o, vis_t = (r_o > 0).all(-1), (r_t > 0).all(-1)
    # radii / visibility: identical except where ceil() or a cull threshold sits within float noise (counted, bounded)
    differ = (r_o != r_t).any(-1)
    assert float(differ.float().mean()) < 2e-4, int(differ.sum())
    both = vis_o & vis_t & ~differ
    assert int(both.sum()) > 40_000
    for o, t, tol in ((ours[1], theirs[1], 1e-5), (ours[2], theirs[2], 1e-6), (ours[3], theirs[3], 2e-5)):
        d = (o[both] - t[both]).abs()
        assert float((d / t[both].abs().clamp_min(1.0)).max()) <= tol
    if comp:
        assert float((ours[4][both] - theirs[4][both]).abs().max()) <= 1e-5
    else:
        assert ours[4] is None


def test_projection_bwd_matches_reference_cuda(rs, ref, refpy):
    """Both backward kernels against each other AND against the float64 autograd of the reference's own torch
    implementation (gsplat/cuda/_torch_impl.py:45-75, 286-374, imported from baseline/_ref): our error against that exact
    gradient may not exceed twice the reference kernel's own error."""
    W, H = 256, 192
    C = 2
    N = 20_000
    s = synthetic_scene(5, N, s_max=0.08)
    vm, Ks = pinhole_cameras(C, W, H)
    ours, theirs = project_both(rs, ref, s, vm, Ks, W, H, comp=True)
    radii, conics, comps = theirs[0], theirs[3], theirs[4]
    g = torch.Generator(device=DEV).manual_seed(0)
    v_m2 = torch.randn(C, N, 2, device=DEV, generator=g)
    v_d = torch.randn(C, N, device=DEV, generator=g)
    v_con = torch.randn(C, N, 3, device=DEV, generator=g) * 0.1
    v_comp = torch.randn(C, N, device=DEV, generator=g)
    args = (T(s["means"]), None, T(s["quats"]), T(s["scales"]), T(vm), T(Ks), W, H, 0.3)
    tail = (radii, conics, comps, v_m2, v_d, v_con, v_comp, True)
    o = rs._C.projection_ewa_3dgs_fused_bwd(*args, rs._C.PINHOLE, *tail)
    t = ref.projection_ewa_3dgs_fused_bwd(*args, ref.PINHOLE, *tail)
    for k, name in ((0, "v_means"), (2, "v_quats"), (3, "v_scales"), (4, "v_viewmats")):
        assert rel_err(o[k], t[k]) < 2e-4, name
    # exact arbiter: float64 autograd through the reference's torch implementation, cotangents masked to the visible rows
    refpy.load_reference(rs._C)
    import importlib

    ti = importlib.import_module("gsplat.cuda._torch_impl")
    leaves = [T(s[k]).double().requires_grad_() for k in ("means", "quats", "scales")] + [T(vm).double().requires_grad_()]
    covars, _ = ti._quat_scale_to_covar_preci(leaves[1], leaves[2], True, False, triu=False)
    _, m2, dep, con, cmp_ = ti._fully_fused_projection(leaves[0], covars, leaves[3], T(Ks).double(), W, H, eps2d=0.3,
                                                       calc_compensations=True)
    vis = (radii > 0).all(-1)
    loss = ((m2 * v_m2.double())[vis].sum() + (dep * v_d.double())[vis].sum() + (con * v_con.double())[vis].sum()
            + (cmp_ * v_comp.double())[vis].sum())
    loss.backward()
    for k, leaf, name in ((0, leaves[0], "v_means"), (2, leaves[1], "v_quats"), (3, leaves[2], "v_scales"),
                          (4, leaves[3], "v_viewmats")):
        # v_viewmats is 2 x 16 sums over ALL rows, accumulated with float atomics in arbitrary order by both kernels: the two
        # errors are single draws of the same kind of noise (both ~1e-6 of the sums), so their ratio scatters from run to
        # run; ours uses one atomic per CTA instead of one per warp (smaller error on average), the bar leaves 3 x.
        bad, info = _worse_than_reference(o[k], t[k], leaf.grad, slack=3.0 if name == "v_viewmats" else 2.0)
        assert not bad, (name, info)


@pytest.mark.parametrize("W,H,C", [(256, 256, 1), (1920, 1080, 1), (500, 300, 4)])
def test_isect_sort_offsets_bit_exact_vs_reference_cuda(rs, ref, W, H, C):
    s = synthetic_scene(21, 60_000, s_max=0.06, spread=1.5)
    vm, Ks = pinhole_cameras(C, W, H)
    ours, _ = project_both(rs, ref, s, vm, Ks, W, H)
    radii, means2d, depths = ours[0], ours[1], ours[2]
    tw, th = (W + 15) // 16, (H + 15) // 16
    a = (means2d, radii, depths, None, None, C, 16, tw, th, True, False)
    tpg_o, ids_o, flat_o = rs._C.intersect_tile(*a)
    tpg_t, ids_t, flat_t = ref.intersect_tile(*a)
    assert ids_t.numel() > 100_000
    assert torch.equal(tpg_o, tpg_t)
    assert torch.equal(ids_o, ids_t)  # sorted keys
    # cub's sort is stable, so is ours: identical order even among equal (tile, depth) keys
    assert torch.equal(flat_o, flat_t)
    off_o = rs._C.intersect_offset(ids_o, C, tw, th)
    off_t = ref.intersect_offset(ids_t, C, tw, th)
    assert torch.equal(off_o, off_t)
    # unsorted emission order is the reference's too
    _, ids_ou, flat_ou = rs._C.intersect_tile(*a[:9], False, False)
    _, ids_tu, flat_tu = ref.intersect_tile(*a[:9], False, False)
    assert torch.equal(ids_ou, ids_tu) and torch.equal(flat_ou, flat_tu)


def test_intersect_tile_repeated_calls_and_cached_offsets(rs, ref):
    """Repeated sorted intersect_tile() calls of one problem shape with different intersection counts, each followed by
    intersect_offset() on the returned ids (served from the offsets the binning produced on the way) and on a COPY of them
    (derived from the 64-bit ids): all against the reference's kernel + cub sort, bit for bit.
    (A variant that sized the outputs from the previous call's count and read the count back only after the sort was
    enqueued was measured and dropped: it moves the host stall behind the sort, where it exposes the enqueue time of the
    compositing call -- 0.56 -> 0.64 ms for the c2 frame through rasterization().)"""
    W, H, C = 640, 360, 2
    s = synthetic_scene(33, 40_000, s_max=0.05, spread=1.5)
    vm, Ks = pinhole_cameras(C, W, H)
    ours, _ = project_both(rs, ref, s, vm, Ks, W, H)
    radii, means2d, depths = ours[0], ours[1], ours[2]
    tw, th = (W + 15) // 16, (H + 15) // 16
    sizes = []
    for scale in (1, 1, 4, 1):
        r = torch.where(radii > 0, radii * scale, radii)
        a = (means2d, r, depths, None, None, C, 16, tw, th, True, False)
        tpg_o, ids_o, flat_o = rs._C.intersect_tile(*a)
        tpg_t, ids_t, flat_t = ref.intersect_tile(*a)
        assert torch.equal(tpg_o, tpg_t) and torch.equal(ids_o, ids_t) and torch.equal(flat_o, flat_t), scale
        want = ref.intersect_offset(ids_t, C, tw, th)
        assert torch.equal(rs._C.intersect_offset(ids_o, C, tw, th), want)          # cached
        assert torch.equal(rs._C.intersect_offset(ids_o.clone(), C, tw, th), want)  # recomputed from the ids
        sizes.append(ids_o.numel())
    assert sizes[2] > 2 * sizes[1] and sizes[3] == sizes[0]


def _raster_inputs(rs, ref, seed, N, W, H, C, D):
    s = synthetic_scene(seed, N, s_max=0.08)
    vm, Ks = pinhole_cameras(C, W, H)
    _, theirs = project_both(rs, ref, s, vm, Ks, W, H)
    radii, means2d, depths, conics = theirs[:4]
    vis = (radii > 0).all(-1)
    # the reference leaves culled rows uninitialised; give both sides the same defined values
    means2d = torch.where(vis[..., None], means2d, torch.zeros_like(means2d))
    conics = torch.where(vis[..., None], conics, torch.zeros_like(conics))
    depths = torch.where(vis, depths, torch.zeros_like(depths))
    tw, th = (W + 15) // 16, (H + 15) // 16
    _, ids, flat = ref.intersect_tile(means2d, radii, depths, None, None, C, 16, tw, th, True, False)
    off = ref.intersect_offset(ids, C, tw, th)
    g = torch.Generator(device=DEV).manual_seed(seed)
    colors = torch.rand(C, N, D, device=DEV, generator=g)
    opac = T(s["opacities"])[None].expand(C, N).contiguous()
    return means2d, conics, colors, opac, off, flat


# channel counts the reference instantiates (gsplat/cuda/csrc/RasterizeToPixels3DGSFwd.cu:204-226)
@pytest.mark.parametrize("D", [1, 3, 4, 16, 32])
def test_raster_fwd_matches_reference_cuda(rs, ref, D):
    W, H, C = 300, 200, 2
    means2d, conics, colors, opac, off, flat = _raster_inputs(rs, ref, 3, 30_000, W, H, C, D)
    g = torch.Generator(device=DEV).manual_seed(1)
    bg = torch.rand(C, D, device=DEV, generator=g)
    a = (means2d, conics, colors, opac, bg, None, W, H, 16, off, flat)
    rc_o, ra_o, li_o = rs._C.rasterize_to_pixels_3dgs_fwd(*a)
    rc_t, ra_t, li_t = ref.rasterize_to_pixels_3dgs_fwd(*a)
    assert float(ra_t.mean()) > 0.05
    assert float((rc_o - rc_t).abs().max()) <= 1e-4
    assert float((ra_o - ra_t).abs().max()) <= 1e-4
    assert torch.equal(li_o, li_t)
    mse = float(((rc_o - rc_t).double() ** 2).mean())
    assert mse == 0.0 or 10 * np.log10(1.0 / mse) >= 60.0
    # stronger, informational-turned-assert: bit identical
    assert torch.equal(ra_o, ra_t)
    assert torch.equal(rc_o, rc_t)


def _worse_than_reference(ours, theirs, exact, slack=2.0):
    """Gradient bar: our kernel's error against the float64-accumulating CPU oracle must not exceed `slack` x the
    reference kernel's own error against it (both in max-abs and in L2), plus one float32 ulp of the tensor's scale --
    i.e. we may differ from the reference only by as much as the reference differs from the exact sum (float atomics in
    another order), instead of round 1's "2e-3 of the tensor max", which hid errors on small entries."""
    exact = exact.to(ours.device, torch.float64)
    e_o, e_t = (ours.double() - exact).abs(), (theirs.double() - exact).abs()
    ulp = float(exact.abs().max()) * 2.0 ** -23
    bad_max = float(e_o.max()) > slack * float(e_t.max()) + ulp
    bad_l2 = float(e_o.norm()) > slack * float(e_t.norm()) + ulp
    return bad_max or bad_l2, dict(ours_max=float(e_o.max()), ref_max=float(e_t.max()), ours_l2=float(e_o.norm()),
                                   ref_l2=float(e_t.norm()), scale=float(exact.abs().max()))


@pytest.mark.parametrize("D,absgrad", [(3, True), (16, False), (20, False)])
def test_raster_bwd_ring_kernel_equals_barrier_kernel(rs, ref, D, absgrad):
    """The two compositing-backward kernels of the library (batch ring with a producer warp; one barrier per batch for
    callers without record scratch) compute the same per-warp sums; only the order of the float atomics between warps
    differs."""
    W, H, C = 200, 160, 2
    means2d, conics, colors, opac, off, flat = _raster_inputs(rs, ref, 8, 20_000, W, H, C, D)
    a = (means2d, conics, colors, opac, None, None, W, H, 16, off, flat)
    rc, ra, li = rs._C.rasterize_to_pixels_3dgs_fwd(*a)
    g = torch.Generator(device=DEV).manual_seed(4)
    v_rc = torch.randn(rc.shape, device=DEV, generator=g)
    v_ra = torch.randn(ra.shape, device=DEV, generator=g)
    ring = rs._C.rasterize_to_pixels_3dgs_bwd(*a, ra, li, v_rc, v_ra, absgrad)
    plain = rs._C.rasterize_to_pixels_3dgs_bwd(*a, ra, li, v_rc, v_ra, absgrad, _ring=False)
    for x, y in zip(ring, plain):
        if x is None:
            assert y is None
            continue
        scale = float(y.abs().max())
        assert scale > 0 and float((x - y).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("D,absgrad", [(3, True), (16, False)])
def test_raster_bwd_matches_reference_cuda(rs, ref, orc, D, absgrad):
    W, H, C = 200, 160, 2
    means2d, conics, colors, opac, off, flat = _raster_inputs(rs, ref, 8, 20_000, W, H, C, D)
    a = (means2d, conics, colors, opac, None, None, W, H, 16, off, flat)
    rc, ra, li = ref.rasterize_to_pixels_3dgs_fwd(*a)
    g = torch.Generator(device=DEV).manual_seed(2)
    v_rc = torch.randn(rc.shape, device=DEV, generator=g)
    v_ra = torch.randn(ra.shape, device=DEV, generator=g)
    o = rs._C.rasterize_to_pixels_3dgs_bwd(*a, ra, li, v_rc, v_ra, absgrad)
    t = ref.rasterize_to_pixels_3dgs_bwd(*a, ra, li, v_rc, v_ra, absgrad)
    cpu = lambda x: x.detach().cpu().numpy()
    exact = orc.rasterize_bwd(cpu(means2d).reshape(-1, 2), cpu(conics).reshape(-1, 3), cpu(colors).reshape(-1, D),
                              cpu(opac).reshape(-1), W, H, 16, cpu(off), cpu(flat), cpu(ra), cpu(li), cpu(v_rc), cpu(v_ra),
                              absgrad=absgrad)
    names = ("v_means2d_abs", "v_means2d", "v_conics", "v_colors", "v_opacities")
    for k in range(0 if absgrad else 1, 5):
        want = torch.from_numpy(exact[names[k]]).reshape(t[k].shape)
        bad, info = _worse_than_reference(o[k], t[k], want)
        assert not bad, (names[k], info)
        assert rel_err(o[k], t[k]) < 2e-4, names[k]
    if not absgrad:
        assert o[0] is None


def _torch_rigid(means, quats, ids, bq, bt, bc):
    """apply_transform semantics (main.py:183-228) in torch, float32, for the reference arm of the end-to-end test."""
    q = bq / bq.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
    sel = ids >= 0
    k = ids.clamp_min(0).long()
    m2 = torch.einsum("nij,nj->ni", R[k], means - bc[k]) + bc[k] + bt[k]
    w1, x1, y1, z1 = q[k].unbind(-1)
    w2, x2, y2, z2 = quats.unbind(-1)
    q2 = torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                      w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)
    return torch.where(sel[:, None], m2, means), torch.where(sel[:, None], q2, quats)


# Whole-frame bound against the reference's apply_transform chain at c2 / c3 (1 M Gaussians, 1080p): the measured counts are
# printed by the tests; these are the asserted ceilings (an order of magnitude below round 1's "1e-4 of all pixels").
C2_MAX_RADII_FLIPS = 8       # Gaussians (of 1 M) whose integer radii differ
C2_MAX_PIXEL_FLIPS = 20      # pixels (of 2 073 600) whose RGB differs by more than 1e-4


def _frame_diff_stats(m, img, alpha, meta_r, rc_r, ra_r):
    """Counted differences between a fused frame (FrameRenderer.meta() + images) and the reference chain's outputs."""
    radii_r = meta_r["radii"].reshape(m["radii"].shape)
    tpg_r = meta_r["tiles_per_gauss"].reshape(m["tiles_per_gauss"].shape)
    err = (rc_r - img).abs()
    mse = float((err.double() ** 2).mean())
    peak = max(float(rc_r.abs().max()), 1e-12)
    return {
        "radii_differ": int((m["radii"] != radii_r).any(-1).sum()),
        "tile_counts_differ": int((m["tiles_per_gauss"] != tpg_r).sum()),
        "n_isects_delta": int(m["n_isects"]) - int(meta_r["flatten_ids"].numel()),
        "pixels_gt_1e-4": int((err > 1e-4).any(-1).sum()),
        "max_abs_rgb": float(err.max()),
        "max_abs_alpha": float((ra_r - alpha).abs().max()),
        "psnr_db": 999.0 if mse == 0.0 else float(10 * np.log10(peak * peak / mse)),
    }


def _reference_frame(ref, means, quats, scales, opac, colors, vm, Ks, W, H):
    C, N = vm.shape[0], means.shape[0]
    radii, means2d, depths, conics, _ = ref.projection_ewa_3dgs_fused_fwd(
        means, None, quats, scales, opac, vm, Ks, W, H, 0.3, 0.01, 1e10, 0.0, False, ref.PINHOLE)
    tw, th = (W + 15) // 16, (H + 15) // 16
    tpg, ids, flat = ref.intersect_tile(means2d, radii, depths, None, None, C, 16, tw, th, True, False)
    off = ref.intersect_offset(ids, C, tw, th)
    rc, ra, li = ref.rasterize_to_pixels_3dgs_fwd(
        means2d, conics, colors[None].expand(C, N, -1).contiguous(), opac[None].expand(C, N).contiguous(), None, None, W,
        H, 16, off, flat)
    return rc, ra, dict(radii=radii, tiles_per_gauss=tpg, isect_ids=ids, flatten_ids=flat, isect_offsets=off)


def test_c1_frame_vs_reference_cuda(rs, ref):
    """c1 (10 k Gaussians, 2 rigid bodies, 256x256): animate + render through our public API vs the reference kernels."""
    W = H = 256
    s = synthetic_scene(42, 10_000, K=2)
    vm, Ks = pinhole_cameras(1, W, H)
    t = {k: T(v) for k, v in s.items()}
    m_t, q_t = _torch_rigid(t["means"], t["quats"], t["cluster_ids"], t["body_quats"], t["body_trans"], t["body_centers"])
    rc_t, ra_t, meta_t = _reference_frame(ref, m_t, q_t, t["scales"], t["opacities"], t["colors"], T(vm), T(Ks), W, H)
    rc_o, ra_o, meta_o = rs.rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], T(vm), T(Ks), W,
                                          H, packed=False, cluster_ids=t["cluster_ids"], body_quats=t["body_quats"],
                                          body_trans=t["body_trans"], body_centers=t["body_centers"])
    assert float((rc_o - rc_t).abs().max()) <= 1e-4
    assert float((ra_o - ra_t).abs().max()) <= 1e-4
    mse = float(((rc_o - rc_t).double() ** 2).mean())
    assert mse == 0.0 or 10 * np.log10(1.0 / mse) >= 60.0
    # the fused rigid transform differs from the torch one by float rounding, so counts may differ on a handful of
    # Gaussians whose radius sits on a ceil() boundary
    d = (meta_o["tiles_per_gauss"] != meta_t["tiles_per_gauss"]).float().mean()
    assert float(d) < 1e-3


def test_c2_full_size_frame_vs_reference_cuda(rs, ref, refpy):
    """c2 size (1 M Gaussians, 20 bodies, 1080p): FrameRenderer (one C-ABI call) vs the reference kernels chained."""
    import bench

    W, H = 1920, 1080
    sc = bench.make_domino_scene(1_000_000, 20, device=DEV)
    bq, bt = bench.domino_poses(20, frame=100, device=DEV, centers=sc["body_centers"])
    fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                          cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"])
    img, alpha = fr.render(sc["viewmats"], sc["Ks"], bq, bt)
    torch.cuda.synchronize()
    assert not fr.overflowed()
    m = fr.meta()
    # (1) stage-wise bit-exactness at full size: the reference's isect + cub sort on OUR projected splats
    tpg_t, ids_t, flat_t = ref.intersect_tile(m["means2d"], m["radii"], m["depths"], None, None, 1, 16, m["tile_width"],
                                              m["tile_height"], True, False)
    assert ids_t.numel() == m["n_isects"] > 1_000_000
    assert torch.equal(m["tiles_per_gauss"], tpg_t)
    assert torch.equal(m["isect_ids"], ids_t)
    assert torch.equal(m["flatten_ids"], flat_t)
    off_t = ref.intersect_offset(ids_t, 1, m["tile_width"], m["tile_height"])
    assert torch.equal(m["isect_offsets"], off_t)
    # (2) compositing on identical sorted lists: bit-identical image
    rc_t, ra_t, li_t = ref.rasterize_to_pixels_3dgs_fwd(m["means2d"], m["conics"], sc["colors"][None].contiguous(),
                                                       sc["opacities"][None].contiguous(), None, None, W, H, 16, off_t,
                                                       flat_t)
    assert torch.equal(li_t, m["last_ids"])
    assert float((rc_t - img).abs().max()) <= 1e-4 and float((ra_t - alpha).abs().max()) <= 1e-4
    # (3) whole frame against the REFERENCE chain: main.py's own apply_transform() once per body (extracted from the
    #     installed reference copy) + the reference kernels.  Counted, not hidden: Gaussians whose radii / tile counts
    #     differ, the intersection-count delta, pixels beyond 1e-4 and the max-abs error.
    m_t, q_t, centers = apply_transform_per_body(refpy, sc, sc["cluster_ids"], bq, bt)
    # the fused path takes the pivots apply_transform() derives itself (means.mean(dim=0) per body, main.py:210)
    fr_c = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                            cluster_ids=sc["cluster_ids"], body_centers=centers)
    img_c, alpha_c = fr_c.render(sc["viewmats"], sc["Ks"], bq, bt)
    torch.cuda.synchronize()
    mc = fr_c.meta()
    rc_r, ra_r, meta_r = _reference_frame(ref, m_t, q_t, sc["scales"], sc["opacities"], sc["colors"], sc["viewmats"],
                                          sc["Ks"], W, H)
    stats = _frame_diff_stats(mc, img_c, alpha_c, meta_r, rc_r, ra_r)
    print("c2 whole frame vs apply_transform chain:", stats)
    assert stats["psnr_db"] >= 60.0
    assert stats["radii_differ"] <= C2_MAX_RADII_FLIPS, stats
    assert stats["pixels_gt_1e-4"] <= C2_MAX_PIXEL_FLIPS, stats
    if stats["radii_differ"] == 0:
        assert stats["n_isects_delta"] == 0 and stats["max_abs_rgb"] <= 1e-4, stats


def test_cpu_oracle_pinned_to_reference_cuda(rs, ref, orc):
    """The CPU restatement (oracle/oracle.c) against the reference CUDA path: this is what pins the oracle."""
    W, H = 256, 256
    s = synthetic_scene(42, 10_000)
    vm, Ks = pinhole_cameras(1, W, H)
    t = {k: T(v) for k, v in s.items()}
    rc_t, ra_t, meta_t = _reference_frame(ref, t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], T(vm),
                                          T(Ks), W, H)
    o = orc.render(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], vm, Ks, W, H)
    clear = o["ambiguous"] == 0
    assert np.array_equal(o["radii"][clear], meta_t["radii"].cpu().numpy()[clear])
    assert (~clear).mean() < 5e-3
    if np.array_equal(o["radii"], meta_t["radii"].cpu().numpy()):
        assert np.array_equal(o["tiles_per_gauss"], meta_t["tiles_per_gauss"].cpu().numpy())
        assert np.array_equal(o["isect_offsets"], meta_t["isect_offsets"].cpu().numpy())
    ok = o["margin"] > 1e-4
    assert np.abs(o["render_colors"] - rc_t.cpu().numpy())[ok].max() <= 1e-4
    assert np.abs(o["render_alphas"] - ra_t.cpu().numpy())[ok].max() <= 1e-4


def test_c4_scale_6m_gaussians_4k_two_cameras(rs, ref):
    """c4 scale on one GPU (6 M Gaussians, 500 bodies, 2 of the 8 ring cameras at 3840x2160 -> 15 tile bits + 2 image bits):
    the depth-ordered binning against the reference's 64-bit cub sort, bit-exact, plus image parity on identical lists."""
    import bench

    W, H, C, N, K = 3840, 2160, 2, 6_000_000, 500
    sc = bench.make_domino_scene(N, K, device=DEV, width=W, height=H, n_cameras=C)
    bq, bt = bench.domino_poses(K, frame=150, device=DEV, centers=sc["body_centers"])
    fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                          cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], n_cameras=C,
                          max_isects=120_000_000)
    img, alpha = fr.render(sc["viewmats"], sc["Ks"], bq, bt)
    torch.cuda.synchronize()
    assert not fr.overflowed()
    m = fr.meta()
    n = m["n_isects"]
    assert n > 5_000_000
    tpg_t, ids_t, flat_t = ref.intersect_tile(m["means2d"], m["radii"], m["depths"], None, None, C, 16, m["tile_width"],
                                              m["tile_height"], True, False)
    assert ids_t.numel() == n
    assert torch.equal(m["tiles_per_gauss"], tpg_t)
    assert torch.equal(m["isect_ids"], ids_t)
    assert torch.equal(m["flatten_ids"], flat_t)
    off_t = ref.intersect_offset(ids_t, C, m["tile_width"], m["tile_height"])
    assert torch.equal(m["isect_offsets"], off_t)
    del tpg_t, ids_t
    rc_t, ra_t, li_t = ref.rasterize_to_pixels_3dgs_fwd(
        m["means2d"], m["conics"], sc["colors"][None].expand(C, N, 3).contiguous(),
        sc["opacities"][None].expand(C, N).contiguous(), None, None, W, H, 16, off_t, flat_t)
    assert torch.equal(li_t, m["last_ids"])
    assert torch.equal(rc_t, img) and torch.equal(ra_t, alpha)
    assert float(alpha.mean()) > 0.005


@pytest.mark.parametrize("deg,K", [(0, 1), (1, 4), (2, 16), (3, 16), (4, 25)])
def test_spherical_harmonics_fwd_bwd_vs_reference_cuda(rs, ref, deg, K):
    """rs_sh_fwd / rs_sh_bwd against the reference's spherical_harmonics_fwd / _bwd and against float64 torch autograd."""
    from oracle import sh_torch

    n = 20_000
    g = torch.Generator(device=DEV).manual_seed(deg)
    dirs = torch.randn(n, 3, device=DEV, generator=g) * 3.0
    coeffs = torch.randn(n, K, 3, device=DEV, generator=g)
    masks = torch.rand(n, device=DEV, generator=g) > 0.2
    v_colors = torch.randn(n, 3, device=DEV, generator=g)
    got = rs._C.spherical_harmonics_fwd(deg, dirs, coeffs, masks)
    want = ref.spherical_harmonics_fwd(deg, dirs, coeffs, masks)
    assert float((got - want)[masks].abs().max()) <= 2e-5
    assert not got[~masks].any()
    vco, vd = rs._C.spherical_harmonics_bwd(K, deg, dirs, coeffs, masks, v_colors, True)
    vco_t, vd_t = ref.spherical_harmonics_bwd(K, deg, dirs, coeffs, masks, v_colors, True)
    assert float((vco - vco_t)[masks].abs().max()) <= 2e-5 * max(1.0, float(vco_t[masks].abs().max()))
    assert float((vd - vd_t)[masks].abs().max()) <= 1e-4 * max(1.0, float(vd_t[masks].abs().max()))
    assert not vco[~masks].any() and not vd[~masks].any()
    # float64 autograd of the torch restatement
    d64 = dirs.double().requires_grad_(True)
    c64 = coeffs.double().requires_grad_(True)
    out = sh_torch.spherical_harmonics_torch(deg, d64, c64, masks)
    assert float((out.detach() - got.double()).abs().max()) <= 2e-5
    out.backward(v_colors.double())
    assert float((c64.grad - vco.double()).abs().max()) <= 2e-5 * max(1.0, float(c64.grad.abs().max()))
    d_grad = d64.grad if d64.grad is not None else torch.zeros_like(d64)  # degree 0 does not depend on the direction
    assert float((d_grad - vd.double()).abs().max()) <= 1e-4 * max(1.0, float(d_grad.abs().max()))
    # the public autograd wrapper
    d32 = dirs.clone().requires_grad_(True)
    c32 = coeffs.clone().requires_grad_(True)
    rs.spherical_harmonics(deg, d32, c32, masks=masks).backward(v_colors)
    assert torch.equal(c32.grad, vco) and torch.equal(d32.grad, vd)


def test_c3_full_size_identity_feature_step_vs_reference_cuda(rs, ref, refpy):
    """c3 (BASELINE configs[2]): 1 M Gaussians, 16-dim identity features, fwd + bwd at 1080p through rasterization() with
    the rigid poses fused in; loss = sum(render * w).  Checked against the reference kernels at full size:
      (1) whole chain (torch rigid transform -> reference projection / intersect_tile / rasterize): image parity;
      (2) the reference's compositing fwd + bwd on OUR projected splats and sorted lists (identical inputs, so no
          threshold can flip): image bit-identical, v_means2d / v_conics / v_features / v_opacities <= 2e-3;
      (3) the reference's projection backward fed with OUR screen-space gradients, chained through the torch rigid
          transform: v_means / v_quats / v_scales <= 2e-3."""
    import bench

    W, H, N, K, D = 1920, 1080, 1_000_000, 20, 16
    sc = bench.make_domino_scene(N, K, device=DEV)
    bq, bt = bench.domino_poses(K, frame=100, device=DEV, centers=sc["body_centers"])
    g = torch.Generator(device=DEV).manual_seed(42)
    feats = torch.randn(N, D, device=DEV, generator=g)
    w = torch.rand(1, H, W, D, device=DEV, generator=g)
    leaves = [sc[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities")] + [feats.clone().requires_grad_()]
    vm, Ks = sc["viewmats"], sc["Ks"]
    m_t, q_t, centers = apply_transform_per_body(refpy, sc, sc["cluster_ids"], bq, bt)  # main.py's own function, per body
    img, alpha, meta = rs.rasterization(*leaves, vm, Ks, W, H, packed=False, cluster_ids=sc["cluster_ids"],
                                        body_quats=bq, body_trans=bt, body_centers=centers)
    meta["means2d"].retain_grad()
    meta["conics"].retain_grad()
    (img * w).sum().backward()
    tw, th = (W + 15) // 16, (H + 15) // 16

    # (1) whole reference chain: apply_transform per body -> reference projection / intersect_tile / rasterize
    rc_r, ra_r, meta_r = _reference_frame(ref, m_t, q_t, sc["scales"], sc["opacities"], feats, vm, Ks, W, H)
    m_ours = dict(radii=meta["radii"], tiles_per_gauss=meta["tiles_per_gauss"], n_isects=meta["flatten_ids"].numel())
    stats = _frame_diff_stats(m_ours, img.detach(), alpha.detach(), meta_r, rc_r, ra_r)
    print("c3 whole frame vs apply_transform chain:", stats)
    assert stats["psnr_db"] >= 60.0
    assert stats["radii_differ"] <= C2_MAX_RADII_FLIPS, stats
    assert stats["pixels_gt_1e-4"] <= C2_MAX_PIXEL_FLIPS, stats
    if stats["radii_differ"] == 0:
        assert stats["n_isects_delta"] == 0 and stats["max_abs_rgb"] <= 1e-4 * max(1.0, float(rc_r.abs().max())), stats

    # (2) reference compositing on our splats / lists
    means2d, conics = meta["means2d"].detach().contiguous(), meta["conics"].detach().contiguous()
    _, ids, flat = ref.intersect_tile(means2d, meta["radii"], meta["depths"].detach(), None, None, 1, 16, tw, th, True, False)
    assert torch.equal(flat, meta["flatten_ids"]) and torch.equal(ids, meta["isect_ids"])
    off = ref.intersect_offset(ids, 1, tw, th)
    a = (means2d, conics, feats[None].contiguous(), sc["opacities"][None].contiguous(), None, None, W, H, 16, off, flat)
    rc, ra, li = ref.rasterize_to_pixels_3dgs_fwd(*a)
    assert torch.equal(rc, img.detach()) and torch.equal(ra, alpha.detach())
    _, v_m2, v_con, v_col, v_op = ref.rasterize_to_pixels_3dgs_bwd(*a, ra, li, w.contiguous(), torch.zeros_like(ra), False)
    assert rel_err(meta["means2d"].grad, v_m2) < 2e-3
    assert rel_err(meta["conics"].grad, v_con) < 2e-3
    assert rel_err(leaves[4].grad, v_col[0]) < 2e-3
    assert rel_err(leaves[3].grad, v_op[0]) < 2e-3

    # (3) reference projection backward on our screen-space gradients, chained back through the rigid transform
    v_means_t, _, v_quats_t, v_scales_t, _ = ref.projection_ewa_3dgs_fused_bwd(
        m_t, None, q_t, sc["scales"], vm, Ks, W, H, 0.3, ref.PINHOLE, meta["radii"], conics, None,
        meta["means2d"].grad.contiguous(), torch.zeros_like(meta["depths"]), meta["conics"].grad.contiguous(), None, False)
    assert rel_err(leaves[2].grad, v_scales_t) < 2e-3
    m_leaf = sc["means"].clone().requires_grad_()
    q_leaf = sc["quats"].clone().requires_grad_()
    m2, q2 = _torch_rigid(m_leaf, q_leaf, sc["cluster_ids"], bq, bt, centers)
    torch.autograd.backward([m2, q2], [v_means_t, v_quats_t])
    assert rel_err(leaves[0].grad, m_leaf.grad) < 2e-3
    assert rel_err(leaves[1].grad, q_leaf.grad) < 2e-3
