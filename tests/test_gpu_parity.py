"""GPU parity tests: the CUDA path (called through the C ABI via the `_C` shim / public API) against the CPU oracle on
identical seeded inputs.  Bars (SURVEY.md section 8c / BASELINE.md section 4):
  * integers (tile counts, keys, sorted order, offsets, flatten ids): bit-exact, stage-wise on identical inputs;
  * projected floats: <= 1e-5 relative (approximate GPU intrinsics vs libm), radii exact wherever the oracle says the
    ceil()/cull decision is not within float noise of flipping;
  * images / alphas: max-abs <= 1e-4 on every pixel whose threshold decisions have a relative margin > 1e-4 in the
    oracle (alpha vs 1/255, T vs 1e-4), PSNR >= 60 dB over ALL pixels;
  * gradients: <= 2e-3 relative to the tensor's max magnitude (float atomics in arbitrary order vs float64 oracle).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, pinhole_cameras, synthetic_scene

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def rel_err(got, ref):
    scale = max(float(np.abs(ref).max()), 1e-12)
    return float(np.abs(got - ref).max()) / scale


def check_projection(out, ref, min_visible=1):
    radii, means2d, depths, conics = (o.cpu().numpy() for o in out[:4])
    clear = ref["ambiguous"] == 0
    assert np.array_equal(radii[clear], ref["radii"][clear])
    assert (~clear).mean() < 5e-3
    both = (radii > 0).all(-1) & (ref["radii"] > 0).all(-1)
    assert both.sum() >= min_visible
    np.testing.assert_allclose(means2d[both], ref["means2d"][both], rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(depths[both], ref["depths"][both], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(conics[both], ref["conics"][both], rtol=2e-4, atol=1e-6)
    culled = ~(radii > 0).all(-1)
    assert not means2d[culled].any() and not conics[culled].any() and not depths[culled].any()
    assert (radii[culled] == 0).all()


# ---------------------------------------------------------------------------------------------------------------------
# projection (+ rigid)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rigid", [False, True])
@pytest.mark.parametrize("C", [1, 3])
def test_project_fwd_vs_oracle(rs, orc, rigid, C):
    W, H = 256, 256
    s = synthetic_scene(42, 10_000, K=2 if rigid else 0)
    vm, Ks = pinhole_cameras(C, W, H)
    m, q = s["means"], s["quats"]
    rp = None
    if rigid:
        m, q = orc.rigid_transform(m, q, s["cluster_ids"], s["body_quats"], s["body_trans"], s["body_centers"])
        rp = rs.RigidPoses(T(s["cluster_ids"]), T(s["body_quats"]), T(s["body_trans"]), T(s["body_centers"]))
    ref = orc.project(m, q, s["scales"], s["opacities"], vm, Ks, W, H)
    out = rs._C.projection_ewa_3dgs_fused_fwd(T(s["means"]), None, T(s["quats"]), T(s["scales"]), T(s["opacities"]),
                                               T(vm), T(Ks), W, H, 0.3, 0.01, 1e10, 0.0, False, rs._C.PINHOLE, rp)
    check_projection(out, ref, min_visible=9000)
    assert out[4] is None


def test_project_fwd_golden_c1(rs):
    """CUDA rigid+projection vs the reference's own apply_transform -> _fully_fused_projection outputs (c1)."""
    g = load_golden("c1_rigid_project_isect.npz")
    W, H = int(g["width"]), int(g["height"])
    rp = rs.RigidPoses(T(g["cluster_ids"]), T(g["body_quats"]), T(g["body_trans"]), T(g["body_centers"]))
    radii, means2d, depths, conics, comp = rs._C.projection_ewa_3dgs_fused_fwd(
        T(g["means"]), None, T(g["quats"]), T(g["scales"]), None, T(g["viewmats"]), T(g["Ks"]), W, H, 0.3, 0.01, 1e10,
        0.0, True, rs._C.PINHOLE, rp)
    radii, means2d, depths, conics, comp = (x.cpu().numpy() for x in (radii, means2d, depths, conics, comp))
    both = (radii > 0).all(-1) & (g["ref_radii"] > 0).all(-1)
    assert both.sum() > 9000
    assert (radii != g["ref_radii"]).any(-1).mean() < 2e-3  # ceil() flips only
    np.testing.assert_allclose(means2d[both], g["ref_means2d"][both], rtol=2e-5, atol=5e-4)
    np.testing.assert_allclose(depths[both], g["ref_depths"][both], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(conics[both], g["ref_conics"][both], rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(comp[both], g["ref_compensations"][both], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("model", ["pinhole", "ortho", "fisheye"])
def test_project_camera_models_batched_fwd_bwd(rs, orc, model):
    """Batch dims + every camera model: forward vs oracle, backward vs the reference's torch autograd (golden)."""
    g = load_golden("projection_fwd_bwd.npz")
    W, H = int(g["width"]), int(g["height"])
    cm = {"pinhole": 0, "ortho": 1, "fisheye": 2}[model]
    p = lambda k: g[f"{model}_{k}"]
    ref = orc.project(p("means"), p("quats"), p("scales"), None, p("viewmats"), p("Ks"), W, H,
                      calc_compensations=True, camera_model=cm)
    means, quats, scales = T(p("means")), T(p("quats")), T(p("scales"))
    out = rs._C.projection_ewa_3dgs_fused_fwd(means, None, quats, scales, None, T(p("viewmats")), T(p("Ks")), W, H, 0.3,
                                               0.01, 1e10, 0.0, True, cm)
    check_projection(out, ref, min_visible=100)
    # backward: cotangents masked to the reference's visible set (as the golden loss was)
    sel = (p("radii") > 0).all(-1)
    radii = T(np.where(sel[..., None], np.maximum(p("radii"), 1), 0).astype(np.int32))
    v_means, _, v_quats, v_scales, _ = rs._C.projection_ewa_3dgs_fused_bwd(
        means, None, quats, scales, T(p("viewmats")), T(p("Ks")), W, H, 0.3, cm, radii, T(p("conics")),
        T(p("compensations")), T(p("v_means2d")), T(p("v_depths")), T(p("v_conics")), T(p("v_compensations")), False)
    for got, key in ((v_means, "g_means"), (v_quats, "g_quats"), (v_scales, "g_scales")):
        ref_g = p(key)
        assert rel_err(got.cpu().numpy(), ref_g) < 2e-3, key


def test_project_bwd_rigid_chain(rs):
    """Gradients through the fused rigid transform == torch autograd through apply_rigid_torch + un-fused projection."""
    import importlib

    tr = importlib.import_module("3dgs_rigidbody_b200.torch_ref")
    W, H = 128, 96
    s = synthetic_scene(5, 3000, K=3)
    vm, Ks = pinhole_cameras(2, W, H)
    rp = rs.RigidPoses(T(s["cluster_ids"]), T(s["body_quats"]), T(s["body_trans"]), T(s["body_centers"]))
    means = T(s["means"]).requires_grad_(True)
    quats = T(s["quats"]).requires_grad_(True)
    scales = T(s["scales"]).requires_grad_(True)
    gen = torch.Generator(device="cpu").manual_seed(1)

    def run(fused):
        for t in (means, quats, scales):
            t.grad = None
        if fused:
            r, m2, d, c, _ = rs.fully_fused_projection(means, None, quats, scales, T(vm), T(Ks), W, H,
                                                       opacities=T(s["opacities"]), rigid=rp)
        else:
            m_t, q_t = tr.apply_rigid_torch(means, quats, rp)
            r, m2, d, c, _ = rs.fully_fused_projection(m_t, None, q_t, scales, T(vm), T(Ks), W, H,
                                                       opacities=T(s["opacities"]))
        gen.manual_seed(1)
        w1 = torch.randn(m2.shape, generator=gen).to(DEV)
        w2 = torch.randn(d.shape, generator=gen).to(DEV)
        w3 = torch.randn(c.shape, generator=gen).to(DEV) * 0.1
        ((m2 * w1).sum() + (d * w2).sum() + (c * w3).sum()).backward()
        return r, [t.grad.clone().cpu().numpy() for t in (means, quats, scales)]

    r1, g1 = run(True)
    r2, g2 = run(False)
    assert (r1 != r2).any(-1).float().mean().item() < 2e-3
    same = ((r1 == r2).all(-1).all(0)).cpu().numpy()  # Gaussians whose visibility agrees in every camera
    for a, b in zip(g1, g2):
        assert rel_err(a[same], b[same]) < 2e-3


# ---------------------------------------------------------------------------------------------------------------------
# tile intersection, sort, offsets: bit-exact on identical inputs
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H,C", [(256, 256, 1), (200, 120, 3), (1920, 1080, 1)])
def test_isect_bit_exact(rs, orc, W, H, C):
    s = synthetic_scene(11, 20_000, spread=1.5)
    vm, Ks = pinhole_cameras(C, W, H)
    pr = orc.project(s["means"], s["quats"], s["scales"], s["opacities"], vm, Ks, W, H)
    tw, th = -(-W // 16), -(-H // 16)
    tpg_r, ids_r, flat_r = orc.isect_tiles(pr["means2d"], pr["radii"], pr["depths"], 16, tw, th)
    off_r = orc.isect_offset_encode(ids_r, C, tw, th)
    tpg, ids, flat = rs.isect_tiles(T(pr["means2d"]), T(pr["radii"]), T(pr["depths"]), 16, tw, th)
    assert np.array_equal(tpg.cpu().numpy(), tpg_r)
    assert np.array_equal(ids.cpu().numpy(), ids_r)
    assert np.array_equal(flat.cpu().numpy(), flat_r)
    off = rs.isect_offset_encode(ids, C, tw, th)
    assert np.array_equal(off.cpu().numpy(), off_r)
    # unsorted emission order too (ascending element, tiles y-outer x-inner)
    _, ids_u, flat_u = rs.isect_tiles(T(pr["means2d"]), T(pr["radii"]), T(pr["depths"]), 16, tw, th, sort=False)
    _, ids_ur, flat_ur = orc.isect_tiles(pr["means2d"], pr["radii"], pr["depths"], 16, tw, th, sort=False)
    assert np.array_equal(ids_u.cpu().numpy(), ids_ur) and np.array_equal(flat_u.cpu().numpy(), flat_ur)
    # segmented=True gives the same order (images are already grouped)
    _, ids_s, flat_s = rs.isect_tiles(T(pr["means2d"]), T(pr["radii"]), T(pr["depths"]), 16, tw, th, segmented=True)
    assert np.array_equal(ids_s.cpu().numpy(), ids_r) and np.array_equal(flat_s.cpu().numpy(), flat_r)


def test_isect_depth_ties_keep_index_order(rs, orc):
    """Equal (image, tile, depth) keys must stay in ascending flatten-id order (stable sort of ascending emission)."""
    N = 3000
    rng = np.random.default_rng(0)
    means2d = (rng.random((1, N, 2)) * 64).astype(np.float32)
    radii = rng.integers(1, 20, size=(1, N, 2)).astype(np.int32)
    depths = rng.integers(1, 4, size=(1, N)).astype(np.float32)  # only 3 distinct depths -> masses of ties
    _, ids_r, flat_r = orc.isect_tiles(means2d, radii, depths, 16, 4, 4)
    _, ids, flat = rs.isect_tiles(T(means2d), T(radii), T(depths), 16, 4, 4)
    assert np.array_equal(ids.cpu().numpy(), ids_r) and np.array_equal(flat.cpu().numpy(), flat_r)


def test_isect_packed_and_edge_cases(rs, orc):
    # packed rows with explicit image ids
    rng = np.random.default_rng(2)
    nnz, I = 5000, 3
    means2d = (rng.random((nnz, 2)) * np.array([96, 64])).astype(np.float32)
    radii = rng.integers(0, 12, size=(nnz, 2)).astype(np.int32)
    depths = (rng.random(nnz) * 10 + 0.1).astype(np.float32)
    image_ids = np.sort(rng.integers(0, I, size=nnz)).astype(np.int64)
    gaussian_ids = rng.integers(0, 1000, size=nnz).astype(np.int64)
    tpg_r, ids_r, flat_r = orc.isect_tiles(means2d, radii, depths, 16, 6, 4, n_images=I, image_ids=image_ids)
    tpg, ids, flat = rs.isect_tiles(T(means2d), T(radii), T(depths), 16, 6, 4, packed=True, n_images=I,
                                    image_ids=T(image_ids), gaussian_ids=T(gaussian_ids))
    assert np.array_equal(tpg.cpu().numpy(), tpg_r) and np.array_equal(ids.cpu().numpy(), ids_r)
    assert np.array_equal(flat.cpu().numpy(), flat_r)
    assert np.array_equal(rs.isect_offset_encode(ids, I, 6, 4).cpu().numpy(), orc.isect_offset_encode(ids_r, I, 6, 4))
    # nothing visible -> empty outputs, offsets all zero
    z = torch.zeros(1, 100, 2, device=DEV)
    tpg, ids, flat = rs.isect_tiles(z, torch.zeros(1, 100, 2, dtype=torch.int32, device=DEV),
                                    torch.ones(1, 100, device=DEV), 16, 4, 4)
    assert ids.numel() == 0 and flat.numel() == 0 and not tpg.any()
    assert not rs.isect_offset_encode(ids, 1, 4, 4).any()
    # zero Gaussians
    tpg, ids, flat = rs.isect_tiles(torch.zeros(1, 0, 2, device=DEV), torch.zeros(1, 0, 2, dtype=torch.int32, device=DEV),
                                    torch.zeros(1, 0, device=DEV), 16, 4, 4)
    assert tpg.shape == (1, 0) and ids.numel() == 0
    # one splat covering every tile, partly off-screen (negative tile bounds saturate to 0)
    means2d = np.array([[[-5.0, 30.0]]], np.float32)
    radii = np.array([[[4000, 4000]]], np.int32)
    tpg, ids, flat = rs.isect_tiles(T(means2d), T(radii), T(np.array([[2.0]], np.float32)), 16, 7, 5)
    assert int(tpg.item()) == 35 and np.array_equal(ids.cpu().numpy() >> 32, np.arange(35))


@pytest.mark.parametrize("n", [1, 2, 4095, 4096, 4097, 100_003, 1_000_000])
def test_radix_sort_pairs(rs, n):
    """The hand-written LSD sort == numpy stable argsort on the masked key bits (any bit range, bits above ignored)."""
    import ctypes
    import importlib

    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    lib = _lib.load()
    rng = np.random.default_rng(n)
    for end_bit, spread in ((46, 46), (13, 60), (64, 63), (33, 8)):
        keys = rng.integers(0, 1 << spread, size=n, dtype=np.int64)
        vals = np.arange(n, dtype=np.int32)
        ka, va = T(keys), T(vals)
        kb, vb = torch.empty_like(ka), torch.empty_like(va)
        ws_bytes = lib.rs_radix_sort_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
        in_b = ctypes.c_int32(0)
        sa = _lib.rs_sort_args()
        sa.n, sa.n_dev, sa.begin_bit, sa.end_bit = n, None, 0, end_bit
        sa.keys_a, sa.keys_b, sa.vals_a, sa.vals_b = ka.data_ptr(), kb.data_ptr(), va.data_ptr(), vb.data_ptr()
        sa.workspace, sa.workspace_bytes = ws.data_ptr(), ws_bytes
        sa.result_in_b = ctypes.addressof(in_b)
        _lib.check(lib.rs_radix_sort_pairs(ctypes.byref(sa), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ko, vo = (kb, vb) if in_b.value else (ka, va)
        mask = (1 << end_bit) - 1 if end_bit < 64 else -1
        order = np.argsort((keys & mask).astype(np.uint64), kind="stable")
        assert np.array_equal(ko.cpu().numpy(), keys[order]), (n, end_bit)
        assert np.array_equal(vo.cpu().numpy(), vals[order]), (n, end_bit)


# ---------------------------------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------------------------------
def raster_inputs(orc, seed, N, W, H, C=1, D=3, s_max=0.08, spread=1.2):
    s = synthetic_scene(seed, N, s_max=s_max, spread=spread)
    vm, Ks = pinhole_cameras(C, W, H)
    pr = orc.project(s["means"], s["quats"], s["scales"], s["opacities"], vm, Ks, W, H)
    tw, th = -(-W // 16), -(-H // 16)
    _, ids, flat = orc.isect_tiles(pr["means2d"], pr["radii"], pr["depths"], 16, tw, th)
    off = orc.isect_offset_encode(ids, C, tw, th)
    rng = np.random.default_rng(seed + 1)
    colors = rng.random((C, N, D)).astype(np.float32)
    opac = np.broadcast_to(s["opacities"][None], (C, N)).copy()
    return pr, flat, off, colors, opac


def check_images(rc, ra, li, ref, atol=1e-4):
    rc_r, ra_r, li_r, mg = ref
    clear = mg > 1e-4
    assert clear.mean() > 0.995
    assert np.abs(rc - rc_r)[clear].max() <= atol
    assert np.abs(ra - ra_r)[clear].max() <= atol
    assert np.array_equal(li[clear], li_r[clear])
    assert psnr(rc, rc_r) >= 60.0


@pytest.mark.parametrize("D", [1, 3, 4, 5, 7, 16, 17, 32, 40])
def test_raster_fwd_channels(rs, orc, D):
    W, H = 160, 112
    pr, flat, off, colors, opac = raster_inputs(orc, 3, 6000, W, H, C=2, D=D)
    bg = np.random.default_rng(9).random((2, D)).astype(np.float32) if D % 2 else None
    ref = orc.rasterize_fwd(pr["means2d"].reshape(-1, 2), pr["conics"].reshape(-1, 3), colors.reshape(-1, D),
                            opac.reshape(-1), W, H, 16, off, flat, backgrounds=bg)
    rc, ra, li = rs._C.rasterize_to_pixels_3dgs_fwd(T(pr["means2d"]), T(pr["conics"]), T(colors), T(opac),
                                                    None if bg is None else T(bg), None, W, H, 16, T(off), T(flat))
    check_images(rc.cpu().numpy(), ra.cpu().numpy(), li.cpu().numpy(), ref)


def test_raster_fwd_ragged_masks_and_empty(rs, orc):
    W, H = 250, 130  # not multiples of 16
    pr, flat, off, colors, opac = raster_inputs(orc, 4, 8000, W, H, C=1, D=3)
    th, tw = off.shape[-2:]
    masks = np.random.default_rng(0).random((1, th, tw)) > 0.3
    bg = np.array([[0.2, 0.4, 0.6]], np.float32)
    args = (pr["means2d"].reshape(-1, 2), pr["conics"].reshape(-1, 3), colors.reshape(-1, 3), opac.reshape(-1))
    ref = orc.rasterize_fwd(*args, W, H, 16, off, flat, backgrounds=bg, masks=masks)
    rc, ra, li = rs._C.rasterize_to_pixels_3dgs_fwd(T(pr["means2d"]), T(pr["conics"]), T(colors), T(opac), T(bg),
                                                    T(masks), W, H, 16, T(off), T(flat))
    # masked tiles: colours = background, alphas / last_ids untouched (compare only unmasked pixels for those)
    pix_mask = np.kron(masks[0], np.ones((16, 16), bool))[:H, :W][None]
    rc_n, ra_n, li_n = rc.cpu().numpy(), ra.cpu().numpy(), li.cpu().numpy()
    assert np.allclose(rc_n[~pix_mask], np.broadcast_to(bg[0], rc_n.shape)[~pix_mask])
    clear = (ref[3] > 1e-4) & pix_mask
    assert np.abs(rc_n - ref[0])[clear].max() <= 1e-4
    assert np.abs(ra_n - ref[1])[clear].max() <= 1e-4
    assert np.array_equal(li_n[clear], ref[2][clear])
    # no intersections at all: background only, alpha 0, last_ids 0
    z = torch.zeros(0, dtype=torch.int32, device=DEV)
    offs0 = torch.zeros(1, th, tw, dtype=torch.int32, device=DEV)
    rc, ra, li = rs._C.rasterize_to_pixels_3dgs_fwd(T(pr["means2d"]), T(pr["conics"]), T(colors), T(opac), T(bg), None,
                                                    W, H, 16, offs0, z)
    assert torch.allclose(rc, T(bg).view(1, 1, 1, 3).expand_as(rc)) and not ra.any() and not li.any()


@pytest.mark.parametrize("D,absgrad", [(3, False), (3, True), (16, True), (17, False), (40, False)])
def test_raster_bwd_vs_oracle(rs, orc, D, absgrad):
    W, H = 128, 96
    pr, flat, off, colors, opac = raster_inputs(orc, 6, 4000, W, H, C=2, D=D)
    bg = np.random.default_rng(1).random((2, D)).astype(np.float32)
    a = (pr["means2d"].reshape(-1, 2), pr["conics"].reshape(-1, 3), colors.reshape(-1, D), opac.reshape(-1))
    rc, ra, li = rs._C.rasterize_to_pixels_3dgs_fwd(T(pr["means2d"]), T(pr["conics"]), T(colors), T(opac), T(bg), None,
                                                    W, H, 16, T(off), T(flat))
    rng = np.random.default_rng(2)
    v_rc = rng.normal(size=tuple(rc.shape)).astype(np.float32)
    v_ra = rng.normal(size=tuple(ra.shape)).astype(np.float32)
    # the oracle walks the GPU's own forward state (alphas, last_ids) so both sides differentiate the same function
    ref = orc.rasterize_bwd(*a, W, H, 16, off, flat, ra.cpu().numpy(), li.cpu().numpy(), v_rc, v_ra, backgrounds=bg,
                            absgrad=absgrad)
    vabs, vm, vc, vcol, vo = rs._C.rasterize_to_pixels_3dgs_bwd(
        T(pr["means2d"]), T(pr["conics"]), T(colors), T(opac), T(bg), None, W, H, 16, T(off), T(flat), ra, li, T(v_rc),
        T(v_ra), absgrad)
    assert rel_err(vm.cpu().numpy().reshape(-1, 2), ref["v_means2d"]) < 2e-3
    assert rel_err(vc.cpu().numpy().reshape(-1, 3), ref["v_conics"]) < 2e-3
    assert rel_err(vcol.cpu().numpy().reshape(-1, D), ref["v_colors"]) < 2e-3
    assert rel_err(vo.cpu().numpy().reshape(-1), ref["v_opacities"]) < 2e-3
    if absgrad:
        assert rel_err(vabs.cpu().numpy().reshape(-1, 2), ref["v_means2d_abs"]) < 2e-3
    else:
        assert vabs is None


# ---------------------------------------------------------------------------------------------------------------------
# public API end to end
# ---------------------------------------------------------------------------------------------------------------------
def test_rasterization_c1_vs_oracle(rs, orc):
    """Config c1: 10k Gaussians, 2 rigid clusters, one 256x256 frame, whole path vs the oracle's whole path."""
    g = load_golden("c1_rigid_project_isect.npz")
    W, H = int(g["width"]), int(g["height"])
    ref = orc.render(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], g["viewmats"], g["Ks"], W, H,
                     cluster_ids=g["cluster_ids"], body_quats=g["body_quats"], body_trans=g["body_trans"],
                     body_centers=g["body_centers"])
    img, alpha, meta = rs.rasterization(
        T(g["means"]), T(g["quats"]), T(g["scales"]), T(g["opacities"]), T(g["colors"]), T(g["viewmats"]), T(g["Ks"]),
        W, H, packed=False, cluster_ids=T(g["cluster_ids"]), body_quats=T(g["body_quats"]),
        body_trans=T(g["body_trans"]), body_centers=T(g["body_centers"]))
    radii = meta["radii"].cpu().numpy()
    mism = (radii != ref["radii"]).any(-1)
    assert mism.mean() < 2e-3
    if not mism.any():  # identical tile lists -> identical keys
        assert np.array_equal(meta["tiles_per_gauss"].cpu().numpy(), ref["tiles_per_gauss"])
        assert np.array_equal(meta["flatten_ids"].cpu().numpy(), ref["flatten_ids"])
        assert np.array_equal(meta["isect_offsets"].cpu().numpy(), ref["isect_offsets"])
    img_n, alpha_n = img.cpu().numpy(), alpha.cpu().numpy()
    assert psnr(img_n, ref["render_colors"]) >= 60.0
    clear = ref["margin"] > 1e-3
    frac_bad = (np.abs(img_n - ref["render_colors"]).max(-1)[clear] > 1e-4).mean()
    assert frac_bad < 1e-3  # end-to-end: upstream 1-ulp differences may flip isolated threshold decisions
    assert set(meta) >= {"batch_ids", "camera_ids", "gaussian_ids", "radii", "means2d", "depths", "conics", "opacities",
                         "tile_width", "tile_height", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets",
                         "width", "height", "tile_size", "n_batches", "n_cameras"}
    assert img.shape == (1, H, W, 3) and alpha.shape == (1, H, W, 1)


def test_rasterization_modes_packed_and_grads(rs):
    """packed == unpacked images; RGB+ED / D modes; backgrounds; means2d keeps a grad (+absgrad) for densification."""
    W, H = 176, 100
    s = synthetic_scene(8, 5000, s_max=0.08)
    vm, Ks = pinhole_cameras(2, W, H)
    args = [T(s[k]) for k in ("means", "quats", "scales", "opacities", "colors")] + [T(vm), T(Ks), W, H]
    img_u, a_u, meta_u = rs.rasterization(*args, packed=False)
    img_p, a_p, meta_p = rs.rasterization(*args, packed=True)
    assert torch.equal(img_u, img_p) and torch.equal(a_u, a_p)
    nnz = int((meta_u["radii"] > 0).all(-1).sum())
    assert meta_p["means2d"].shape == (nnz, 2) and meta_p["camera_ids"].shape == (nnz,)
    img_ed, a_ed, _ = rs.rasterization(*args, packed=False, render_mode="RGB+ED")
    assert img_ed.shape[-1] == 4 and torch.equal(img_ed[..., :3], img_u)
    img_d, _, _ = rs.rasterization(*args, packed=False, render_mode="D")
    ed = img_d / a_u.clamp(min=1e-10)
    assert torch.allclose(img_ed[..., 3:], ed, rtol=1e-5, atol=1e-6)
    bg = torch.rand(2, 3, device=DEV)
    img_bg, _, _ = rs.rasterization(*args, packed=False, backgrounds=bg)
    assert torch.allclose(img_bg, img_u + (1 - a_u) * bg[:, None, None, :], atol=1e-6)
    # training-style call: grads reach every parameter and meta["means2d"]
    params = [T(s[k]).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")]
    img, alpha, meta = rs.rasterization(*params, T(vm), T(Ks), W, H, packed=False, absgrad=True)
    meta["means2d"].retain_grad()
    (img.sum() + alpha.sum()).backward()
    for p_ in params:
        assert p_.grad is not None and torch.isfinite(p_.grad).all() and p_.grad.abs().sum() > 0
    assert meta["means2d"].grad.shape == meta["means2d"].shape
    assert meta["means2d"].absgrad.shape == meta["means2d"].shape
    assert (meta["means2d"].absgrad >= meta["means2d"].grad.abs() - 1e-6).all()


def test_rasterization_sh_and_antialiased(rs):
    W, H = 128, 80
    s = synthetic_scene(9, 3000, s_max=0.08)
    vm, Ks = pinhole_cameras(1, W, H)
    sh = torch.randn(3000, 16, 3, device=DEV) * 0.2
    img, alpha, meta = rs.rasterization(T(s["means"]), T(s["quats"]), T(s["scales"]), T(s["opacities"]), sh, T(vm),
                                        T(Ks), W, H, sh_degree=3, packed=False, render_mode="RGB+ED",
                                        rasterize_mode="antialiased")
    assert img.shape == (1, H, W, 4) and torch.isfinite(img).all() and (alpha >= 0).all() and (alpha <= 1).all()


def test_frame_renderer_matches_compat_path(rs, orc):
    """The sync-free fused frame path == rasterization(packed=False) bit for bit, and both match the oracle."""
    W, H = 320, 176
    s = synthetic_scene(12, 30_000, K=5, s_max=0.06, spread=1.5)
    vm, Ks = pinhole_cameras(2, W, H)
    fr = rs.FrameRenderer(T(s["means"]), T(s["quats"]), T(s["scales"]), T(s["opacities"]), T(s["colors"]), W, H,
                          cluster_ids=T(s["cluster_ids"]), body_centers=T(s["body_centers"]), n_cameras=2)
    img_f, a_f = fr.render(T(vm), T(Ks), T(s["body_quats"]), T(s["body_trans"]))
    img_f, a_f = img_f.clone(), a_f.clone()
    img_c, a_c, meta = rs.rasterization(
        T(s["means"]), T(s["quats"]), T(s["scales"]), T(s["opacities"]), T(s["colors"]), T(vm), T(Ks), W, H,
        packed=False, cluster_ids=T(s["cluster_ids"]), body_quats=T(s["body_quats"]), body_trans=T(s["body_trans"]),
        body_centers=T(s["body_centers"]))
    assert torch.equal(img_f, img_c) and torch.equal(a_f, a_c)
    m = fr.meta()
    assert m["n_isects"] == meta["flatten_ids"].numel() and not fr.overflowed()
    assert torch.equal(m["flatten_ids"], meta["flatten_ids"]) and torch.equal(m["isect_ids"], meta["isect_ids"])
    assert torch.equal(m["isect_offsets"], meta["isect_offsets"]) and torch.equal(m["radii"], meta["radii"])
    # idempotence: a second render of the same frame is bit-identical
    img_2, _ = fr.render(T(vm), T(Ks), T(s["body_quats"]), T(s["body_trans"]))
    assert torch.equal(img_2, img_f)
    # overflow is reported, not silently ignored
    small = rs.FrameRenderer(T(s["means"]), T(s["quats"]), T(s["scales"]), T(s["opacities"]), T(s["colors"]), W, H,
                             cluster_ids=T(s["cluster_ids"]), body_centers=T(s["body_centers"]), n_cameras=2,
                             max_isects=1000)
    small.render(T(vm), T(Ks), T(s["body_quats"]), T(s["body_trans"]))
    assert small.overflowed() and small.n_isects() == m["n_isects"]
    assert small.ensure_capacity() and not small.overflowed() is None
    img_3, _ = small.render(T(vm), T(Ks), T(s["body_quats"]), T(s["body_trans"]))
    assert torch.equal(img_3, img_f)


@pytest.mark.parametrize("case", ["small_splats", "large_anisotropic", "low_opacity", "general_kernel", "huge_splats",
                                  "c2_full_size"])
def test_tight_tile_lists_give_the_same_image(rs, case):
    """FrameRenderer(tight_tiles=True) lists a (tile, splat) pair only where the splat can reach alpha >= 1/255 inside the
    tile; every dropped pair is one the compositing of RasterizeToPixels3DGSFwd.cu:148-149 skips at every pixel, so the float
    image, the alphas and the 8-bit frame must be BIT-identical to the frame rendered from the reference's lists, and the
    tight lists must be an order-preserving subset of them."""
    import bench

    if case == "c2_full_size":
        sc = bench.make_domino_scene(1_000_000, 20, device=DEV)
        W, H, C = 1920, 1080, 1
        vm, Ks = sc["viewmats"], sc["Ks"]
        bq, bt = bench.domino_poses(20, frame=77, device=DEV, centers=sc["body_centers"])
        scene = (sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"])
        ids, centers = sc["cluster_ids"], sc["body_centers"]
    else:
        W, H, C = 400, 300, 2
        kw = dict(small_splats=dict(s_max=0.03), large_anisotropic=dict(s_max=0.5, spread=1.5),
                  low_opacity=dict(s_max=0.15), general_kernel=dict(s_max=0.08), huge_splats=dict(s_max=3.0))[case]
        # huge_splats: every rectangle has far more than 64 tiles (kept whole) and a CTA of the emission meets more such
        # elements than its whole-CTA list holds (256), so the per-thread rectangle walk of the masked kernel runs too
        s = synthetic_scene(21, 6_000 if case == "huge_splats" else 40_000, K=4, **kw)
        if case == "large_anisotropic":  # needles: one long axis
            s["scales"][:, 1:] *= 0.04
        if case == "low_opacity":  # many splats near the 1/255 threshold
            s["opacities"] = (s["opacities"] * 0.03).astype(np.float32)
        vmn, Ksn = pinhole_cameras(C, W, H)
        vm, Ks = T(vmn), T(Ksn)
        scene = tuple(T(s[k]) for k in ("means", "quats", "scales", "opacities", "colors"))
        if case == "general_kernel":  # means not 16-byte aligned: the projection kernel without TMA staging, which
            # computes the footprints per lane instead of per warp
            pad = torch.cat([torch.zeros(1, 3, device=DEV), scene[0]])
            scene = (pad[1:],) + scene[1:]
            assert scene[0].data_ptr() % 16 != 0 and scene[0].is_contiguous()
        ids, centers, bq, bt = T(s["cluster_ids"]), T(s["body_centers"]), T(s["body_quats"]), T(s["body_trans"])
    out = {}
    for tight in (False, True):
        fr = rs.FrameRenderer(*scene, W, H, cluster_ids=ids, body_centers=centers, n_cameras=C, rgb8=True, tight_tiles=tight,
                              max_isects=(8 << 20) if case == "huge_splats" else None)
        img, alpha = fr.render(vm, Ks, bq, bt)
        torch.cuda.synchronize()
        assert not fr.overflowed()
        m = fr.meta()
        out[tight] = (img.clone(), alpha.clone(), fr.render_rgb8.clone(), m["n_isects"], m["flatten_ids"].clone(),
                      m["isect_offsets"].clone(), m["isect_ids"].clone())
        del fr
    ref_l, tight_l = out[False], out[True]
    assert torch.equal(ref_l[0], tight_l[0]) and torch.equal(ref_l[1], tight_l[1]) and torch.equal(ref_l[2], tight_l[2])
    assert float(ref_l[1].mean()) > 0.01
    assert tight_l[3] <= ref_l[3]
    if case not in ("low_opacity", "huge_splats"):
        assert tight_l[3] < ref_l[3], "the test scene does not exercise the culling"
    # subset in the same order: the tight list is the reference list with some entries removed
    keys_ref = ref_l[6]
    keys_tight = tight_l[6]
    assert bool((keys_tight[1:] >= keys_tight[:-1]).all())
    pair = lambda keys, ids: (keys >> 32) * (1 << 31) + ids.long()  # (image|tile, flatten id) identifies an entry
    pr, pt = pair(keys_ref, ref_l[4]), pair(keys_tight, tight_l[4])
    pos = torch.searchsorted(pr.sort().values, pt)
    assert bool((pr.sort().values[pos.clamp(max=pr.numel() - 1)] == pt).all())
    print(f"{case}: {ref_l[3]} -> {tight_l[3]} intersections ({tight_l[3] / max(ref_l[3], 1):.3f})")


def test_tight_tile_lists_drop_only_pairs_no_pixel_can_use(rs):
    """The direct statement of the culling rule, independent of occlusion: every (tile, splat) pair of the reference's list
    that the tight list leaves out has alpha < 1/255 at ALL 256 pixel centres of the tile (evaluated here in float64), i.e.
    RasterizeToPixels3DGSFwd.cu:148-149 would skip it at every pixel."""
    W, H, C, N = 400, 300, 2, 40_000
    s = synthetic_scene(33, N, K=4, s_max=0.25, spread=1.3)
    s["scales"][::2, 1:] *= 0.1  # half of the splats needle-shaped
    vmn, Ksn = pinhole_cameras(C, W, H)
    scene = tuple(T(s[k]) for k in ("means", "quats", "scales", "opacities", "colors"))
    kw = dict(cluster_ids=T(s["cluster_ids"]), body_centers=T(s["body_centers"]), n_cameras=C)
    lists = {}
    for tight in (False, True):
        fr = rs.FrameRenderer(*scene, W, H, tight_tiles=tight, **kw)
        fr.render(T(vmn), T(Ksn), T(s["body_quats"]), T(s["body_trans"]))
        torch.cuda.synchronize()
        m = fr.meta()
        lists[tight] = ((m["isect_ids"] >> 32).clone(), m["flatten_ids"].long().clone())
        if not tight:
            means2d, conics = m["means2d"].reshape(-1, 2).double().clone(), m["conics"].reshape(-1, 3).double().clone()
            tw, th = m["tile_width"], m["tile_height"]
        del fr
    pair = lambda kt: kt[0] * (1 << 31) + kt[1]
    pr, pt = pair(lists[False]), pair(lists[True])
    dropped = ~torch.isin(pr, pt)
    assert int(dropped.sum()) > 1000 and bool(torch.isin(pt, pr).all())
    tile_bits = int(tw * th).bit_length()  # IntersectTile.cu: floor(log2(n_tiles)) + 1
    key, row = lists[False][0][dropped], lists[False][1][dropped]
    tile = key & ((1 << tile_bits) - 1)
    tx, ty = (tile % tw).double(), torch.div(tile, tw, rounding_mode="floor").double()
    off = torch.arange(16, device=DEV, dtype=torch.float64) + 0.5
    px = (tx * 16)[:, None, None] + off[None, None, :]  # [P,1,16]
    py = (ty * 16)[:, None, None] + off[None, :, None]  # [P,16,1]
    mu, cn = means2d[row], conics[row]
    op = scene[3].double()[row % N]
    dx, dy = mu[:, 0, None, None] - px, mu[:, 1, None, None] - py
    sigma = 0.5 * (cn[:, 0, None, None] * dx * dx + cn[:, 2, None, None] * dy * dy) + cn[:, 1, None, None] * dx * dy
    alpha = torch.minimum(torch.full_like(sigma, 0.999), op[:, None, None] * torch.exp(-sigma))
    alpha = torch.where(sigma < 0, torch.zeros_like(alpha), alpha)
    worst = float(alpha.reshape(alpha.shape[0], -1).max(dim=1).values.max())
    assert worst < (1.0 / 255.0) * (1.0 - 5e-4), worst


def test_full_size_properties_1m_1080p(rs):
    """BASELINE c2 size (1 M Gaussians, 20 bodies, 1080p): size-independent properties instead of a CPU comparison."""
    import bench

    sc = bench.make_domino_scene(1_000_000, 20, device=DEV)
    W, H = 1920, 1080
    fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], W, H,
                          cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"])
    bq, bt = bench.domino_poses(20, frame=60, device=DEV)
    img, alpha = fr.render(sc["viewmats"], sc["Ks"], bq, bt)
    torch.cuda.synchronize()
    assert not fr.overflowed()
    m = fr.meta()
    n = m["n_isects"]
    assert n > 1_000_000
    assert int(m["tiles_per_gauss"].sum()) == n  # checksum of the tile counts
    keys = m["isect_ids"]
    assert bool((keys[1:] >= keys[:-1]).all())  # sortedness
    tile_of_key = (keys >> 32) & ((1 << 13) - 1)
    counts = torch.bincount(tile_of_key, minlength=m["tile_width"] * m["tile_height"])
    off = m["isect_offsets"].reshape(-1).long()
    assert torch.equal(torch.cat([off[1:], torch.tensor([n], device=DEV)]) - off, counts)  # offsets <-> keys
    # every isect's flatten id is a visible Gaussian and its depth bits match the key
    flat = m["flatten_ids"].long()
    assert bool((m["radii"].reshape(-1, 2)[flat] > 0).all())
    assert torch.equal(m["depths"].reshape(-1)[flat].view(torch.int32).long() & 0xFFFFFFFF, keys & 0xFFFFFFFF)
    assert bool((alpha >= 0).all()) and bool((alpha <= 1).all()) and bool(torch.isfinite(img).all())
    assert float(alpha.mean()) > 0.01
    img2, _ = fr.render(sc["viewmats"], sc["Ks"], bq, bt)
    assert torch.equal(img2, img)  # idempotence / determinism


def test_frame_pipeline_matches_sequential(rs):
    """Several frames in flight (one stream + workspace each) give exactly the frames a single renderer gives."""
    W, H = 320, 192
    s = synthetic_scene(9, 30_000, K=4)
    vm, Ks = pinhole_cameras(1, W, H)
    t = {k: T(v) for k, v in s.items()}
    common = dict(cluster_ids=t["cluster_ids"], body_centers=t["body_centers"], max_isects=1 << 21)
    args = (t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], W, H)
    fr = rs.FrameRenderer(*args, **common)
    pipe = rs.FramePipeline(3, *args, **common)
    rng = np.random.default_rng(0)
    poses = [(T(rng.normal(size=(4, 4)).astype(np.float32)), T((rng.normal(size=(4, 3)) * 0.3).astype(np.float32)))
             for _ in range(7)]
    want = []
    for q, tr in poses:
        img, alpha = fr.render(T(vm), T(Ks), q, tr)
        want.append((img.clone(), alpha.clone()))
    got = []
    for q, tr in poses:
        img, alpha, done = pipe.submit(T(vm), T(Ks), q, tr)
        done.synchronize()  # consume before the slot is reused
        got.append((img.clone(), alpha.clone()))
    pipe.join()
    # and without consuming in between: only the last `depth` frames remain in the slots
    outs = [pipe.submit(T(vm), T(Ks), q, tr) for q, tr in poses]
    pipe.join()
    torch.cuda.synchronize()
    for (wi, wa), (gi, ga) in zip(want, got):
        assert torch.equal(wi, gi) and torch.equal(wa, ga)
    for k in range(len(poses) - 3, len(poses)):
        assert torch.equal(outs[k][0], want[k][0])
    assert not pipe.overflowed()


def test_frame_renderer_sh_matches_rasterization(rs):
    """SH colours evaluated inside the projection kernel (frame path) vs rasterization(sh_degree=3) (operator path:
    rs_sh_fwd on explicit view directions), rigid poses on."""
    W, H, N, K = 288, 160, 20_000, 3
    s = synthetic_scene(13, N, K=K)
    vm, Ks = pinhole_cameras(2, W, H)
    t = {k: T(v) for k, v in s.items()}
    g = torch.Generator(device=DEV).manual_seed(4)
    coeffs = torch.randn(N, 16, 3, device=DEV, generator=g) * 0.3
    rigid = dict(cluster_ids=t["cluster_ids"], body_quats=t["body_quats"], body_trans=t["body_trans"],
                 body_centers=t["body_centers"])
    want, want_a, _ = rs.rasterization(t["means"], t["quats"], t["scales"], t["opacities"], coeffs, T(vm), T(Ks), W, H,
                                       sh_degree=3, packed=False, **rigid)
    fr = rs.FrameRenderer(t["means"], t["quats"], t["scales"], t["opacities"], coeffs, W, H, cluster_ids=t["cluster_ids"],
                          body_centers=t["body_centers"], n_cameras=2, max_isects=1 << 21, sh_degree=3)
    img, alpha = fr.render(T(vm), T(Ks), t["body_quats"], t["body_trans"])
    assert float(want.abs().max()) > 0.1
    assert float((img - want).abs().max()) <= 1e-4
    assert float((alpha - want_a).abs().max()) <= 1e-4
    # lower degree uses only the first rows
    fr1 = rs.FrameRenderer(t["means"], t["quats"], t["scales"], t["opacities"], coeffs, W, H, cluster_ids=t["cluster_ids"],
                           body_centers=t["body_centers"], n_cameras=2, max_isects=1 << 21, sh_degree=1)
    img1, _ = fr1.render(T(vm), T(Ks), t["body_quats"], t["body_trans"])
    want1, _, _ = rs.rasterization(t["means"], t["quats"], t["scales"], t["opacities"], coeffs, T(vm), T(Ks), W, H,
                                   sh_degree=1, packed=False, **rigid)
    assert float((img1 - want1).abs().max()) <= 1e-4


@pytest.mark.parametrize("packed", [False, True])
def test_rasterization_fused_sh_matches_differentiable_sh_path(rs, packed):
    """rasterization(sh_degree=3, cluster_ids=...) -- main.py:328-339's call plus the rigid poses.  Without gradients the
    colours come out of the projection kernel; with gradients the differentiable spherical_harmonics() operator on explicit
    view directions runs (rendering.py:491-525).  Same image, same meta."""
    W, H, N, K = 288, 160, 20_000, 3
    s = synthetic_scene(13, N, K=K)
    vm, Ks = pinhole_cameras(2, W, H)
    t = {k: T(v) for k, v in s.items()}
    g = torch.Generator(device=DEV).manual_seed(4)
    coeffs = torch.randn(N, 16, 3, device=DEV, generator=g) * 0.3
    bg = torch.rand(2, 3, device=DEV, generator=g)
    kw = dict(sh_degree=3, packed=packed, render_mode="RGB+ED", backgrounds=bg, cluster_ids=t["cluster_ids"],
              body_quats=t["body_quats"], body_trans=t["body_trans"], body_centers=t["body_centers"])
    args = (t["means"], t["quats"], t["scales"], t["opacities"])
    with torch.no_grad():
        img_f, alpha_f, meta_f = rs.rasterization(*args, coeffs, T(vm), T(Ks), W, H, **kw)
    img_d, alpha_d, meta_d = rs.rasterization(*args, coeffs.clone().requires_grad_(True), T(vm), T(Ks), W, H, **kw)
    assert img_f.shape == (2, H, W, 4) and float(img_d[..., :3].abs().max()) > 0.1
    assert torch.equal(meta_f["flatten_ids"], meta_d["flatten_ids"]) and torch.equal(meta_f["radii"], meta_d["radii"])
    assert float((img_f - img_d.detach()).abs().max()) <= 1e-4
    assert torch.equal(alpha_f, alpha_d.detach())


def test_frame_renderer_rgb8_is_the_quantised_float_frame(rs):
    """The 8-bit frame written by the compositing epilogue equals torchvision save_image's quantisation of the float frame
    (what main.py:140-171 save_rendered_image stores): x * 255 + 0.5, clamp to [0, 255], truncate -- with a background."""
    W, H, N = 333, 201, 20_000
    s = synthetic_scene(31, N, K=3, s_max=0.15)
    vm, Ks = pinhole_cameras(2, W, H)
    t = {k: torch.from_numpy(v).to(DEV) for k, v in s.items()}
    bg = torch.tensor([[0.2, 0.9, 1.3], [0.0, 0.5, 0.25]], device=DEV)
    fr = rs.FrameRenderer(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"] * 1.4, W, H,
                          cluster_ids=t["cluster_ids"], body_centers=t["body_centers"], n_cameras=2, backgrounds=bg,
                          rgb8=True)
    img, _ = fr.render(torch.from_numpy(vm).to(DEV), torch.from_numpy(Ks).to(DEV), t["body_quats"], t["body_trans"])
    want = img.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)
    assert fr.render_rgb8.dtype == torch.uint8 and fr.render_rgb8.shape == (2, H, W, 3)
    assert torch.equal(fr.render_rgb8, want)
    assert int(want.max()) == 255 and int(want.min()) < 64  # saturated and dark pixels are both present
