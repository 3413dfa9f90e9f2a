"""GPU parity of the packed (COO) projection: `projection_ewa_3dgs_packed_fwd/bwd`.

Checked three ways:
  * against this package's dense projection on the same inputs: the packed rows must be exactly the visible rows of the
    dense result, bit for bit, in row-major (batch, camera, gaussian) order, with matching `indptr`;
  * against the reference's own CUDA packed operators (oracle/_ref/gsplat_ref_cuda.so) on identical tensors: same rows
    (up to the counted ceil()/threshold noise the dense test also allows), floats <= 1e-5 relative, gradients <= 2e-3;
  * through rasterization(packed=True) against packed=False: identical images and gradients.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, pinhole_cameras, synthetic_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REF_SO = os.path.join(ROOT, "oracle", "_ref", "gsplat_ref_cuda.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/gsplat_ref_cuda.so not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("gsplat_ref_cuda", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rel_err(got, want):
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-12)


def scene_args(N, C, W, H, seed=3, batch=(), covars=False, misalign=False, comp=False, radius_clip=0.0, K=0):
    s = synthetic_scene(seed, N * max(int(np.prod(batch)), 1), s_max=0.08, K=0)
    shp = tuple(batch)
    means = T(s["means"]).reshape(shp + (N, 3))
    quats = T(s["quats"]).reshape(shp + (N, 4))
    scales = T(s["scales"]).reshape(shp + (N, 3))
    opac = T(s["opacities"]).reshape(shp + (N,))
    if misalign:  # contiguous tensors whose base pointers are not 16-byte aligned -> the non-TMA kernel
        means = torch.cat([means.new_zeros(1, 3), means])[1:]
        scales = torch.cat([scales.new_zeros(1, 3), scales])[1:]
        assert means.is_contiguous() and means.data_ptr() % 16 != 0
    vm, Ks = pinhole_cameras(C, W, H)
    vm, Ks = T(vm).expand(shp + (C, 4, 4)).contiguous(), T(Ks).expand(shp + (C, 3, 3)).contiguous()
    cov = None
    if covars:
        from importlib import import_module

        tr = import_module("3dgs_rigidbody_b200.torch_ref")
        M = tr.normalized_quat_to_rotmat(torch.nn.functional.normalize(quats, dim=-1)) * scales[..., None, :]
        full = M @ M.transpose(-1, -2)
        cov = full[..., (0, 0, 0, 1, 1, 2), (0, 1, 2, 1, 2, 2)].contiguous()
        quats = scales = None
    return (means, cov, quats, scales, opac, vm, Ks, W, H, 0.3, 0.01, 1e10, radius_clip, comp)


CASES = [
    dict(N=50_000, C=1),
    dict(N=50_001, C=3, comp=True),  # ragged last chunk, several cameras
    dict(N=20_000, C=2, batch=(2,)),  # batch dims
    dict(N=30_000, C=2, covars=True),  # covariance input -> general kernel
    dict(N=30_011, C=2, misalign=True),  # unaligned tensors -> general kernel
    dict(N=100, C=1),  # fewer pairs than one chunk
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_packed_fwd_is_the_visible_rows_of_the_dense_projection(rs, case):
    case = dict(case)
    N, C = case.pop("N"), case.pop("C")
    args = scene_args(N, C, 320, 200, **case)
    dense = rs._C.projection_ewa_3dgs_fused_fwd(*args, rs._C.PINHOLE)
    packed = rs._C.projection_ewa_3dgs_packed_fwd(*args, rs._C.PINHOLE)
    indptr, b_ids, c_ids, g_ids, radii, means2d, depths, conics, comps = packed
    B = int(np.prod(case.get("batch", ()))) if case.get("batch") else 1
    sel = (dense[0] > 0).all(-1).reshape(B, C, N)
    b_w, c_w, g_w = torch.nonzero(sel, as_tuple=True)
    assert b_ids.dtype == torch.int64 and radii.dtype == torch.int32 and indptr.dtype == torch.int32
    assert torch.equal(b_ids, b_w) and torch.equal(c_ids, c_w) and torch.equal(g_ids, g_w)
    flat = (b_w * C + c_w) * N + g_w
    assert torch.equal(radii, dense[0].reshape(-1, 2)[flat])
    assert torch.equal(means2d, dense[1].reshape(-1, 2)[flat])
    assert torch.equal(depths, dense[2].reshape(-1)[flat])
    assert torch.equal(conics, dense[3].reshape(-1, 3)[flat])
    if case.get("comp"):
        assert torch.equal(comps, dense[4].reshape(-1)[flat])
    else:
        assert comps is None
    want_indptr = torch.zeros(B * C + 1, dtype=torch.int64, device=DEV)
    want_indptr[1:] = sel.reshape(B * C, N).sum(-1).cumsum(0)
    assert torch.equal(indptr.long(), want_indptr)
    assert 0 < b_ids.numel() <= B * C * N


def test_packed_fwd_regrows_from_a_small_capacity_hint(rs):
    args = scene_args(40_000, 2, 320, 200)
    full = rs._C.projection_ewa_3dgs_packed_fwd(*args, rs._C.PINHOLE)
    small = rs._C.projection_ewa_3dgs_packed_fwd(*args, rs._C.PINHOLE, None, 1000)
    assert full[1].numel() > 1000
    for a, b in zip(full, small):
        assert (a is None and b is None) or torch.equal(a, b)


def test_packed_fwd_empty_inputs(rs):
    args = list(scene_args(64, 1, 64, 64))
    args[0], args[2], args[3], args[4] = args[0][:0], args[2][:0], args[3][:0], args[4][:0]
    out = rs._C.projection_ewa_3dgs_packed_fwd(*args, rs._C.PINHOLE)
    assert out[1].numel() == 0 and out[4].shape == (0, 2) and torch.equal(out[0], torch.zeros(2, dtype=torch.int32, device=DEV))
    # everything behind the camera: rows exist but none is visible
    args = list(scene_args(5000, 2, 64, 64))
    args[0] = args[0] - torch.tensor([0.0, 0.0, 100.0], device=DEV)
    out = rs._C.projection_ewa_3dgs_packed_fwd(*args, rs._C.PINHOLE)
    assert out[1].numel() == 0 and torch.equal(out[0], torch.zeros(3, dtype=torch.int32, device=DEV))


def test_packed_fwd_with_rigid_poses_matches_dense(rs):
    N, C, K = 30_000, 2, 7
    s = synthetic_scene(9, N, K=K, s_max=0.08)
    vm, Ks = pinhole_cameras(C, 320, 200)
    rigid = rs._C.RigidPoses(T(s["cluster_ids"]), T(s["body_quats"]), T(s["body_trans"]), T(s["body_centers"]))
    args = (T(s["means"]), None, T(s["quats"]), T(s["scales"]), T(s["opacities"]), T(vm), T(Ks), 320, 200, 0.3, 0.01,
            1e10, 0.0, False, rs._C.PINHOLE, rigid)
    dense = rs._C.projection_ewa_3dgs_fused_fwd(*args)
    packed = rs._C.projection_ewa_3dgs_packed_fwd(*args)
    c_w, g_w = torch.nonzero((dense[0] > 0).all(-1), as_tuple=True)
    assert torch.equal(packed[2], c_w) and torch.equal(packed[3], g_w)
    assert torch.equal(packed[5], dense[1][c_w, g_w]) and torch.equal(packed[7], dense[3][c_w, g_w])


@pytest.mark.parametrize("C,comp", [(1, False), (3, True)])
def test_packed_fwd_matches_reference_cuda(rs, ref, C, comp):
    N = 50_000
    args = scene_args(N, C, 320, 200, seed=11, comp=comp)
    o = rs._C.projection_ewa_3dgs_packed_fwd(*args, rs._C.PINHOLE)
    t = ref.projection_ewa_3dgs_packed_fwd(*args, ref.PINHOLE)
    key_o = o[2] * N + o[3]
    key_t = t[2] * N + t[3]
    # the same rows, except where a cull threshold sits within float noise (counted, bounded)
    common = torch.isin(key_o, key_t)
    assert float((~common).float().mean()) < 2e-4 and abs(key_o.numel() - key_t.numel()) <= 2e-4 * key_t.numel() + 2
    sel_o = common
    sel_t = torch.isin(key_t, key_o)
    assert torch.equal(key_o[sel_o], key_t[sel_t])  # identical order
    same_radii = (o[4][sel_o] == t[4][sel_t]).all(-1)
    assert float((~same_radii).float().mean()) < 2e-4
    for k, tol in ((5, 1e-5), (6, 1e-6), (7, 2e-5)):
        a, b = o[k][sel_o], t[k][sel_t]
        assert float(((a - b).abs() / b.abs().clamp_min(1.0)).max()) <= tol, k
    if comp:
        assert float((o[8][sel_o] - t[8][sel_t]).abs().max()) <= 1e-5
    if key_o.numel() == key_t.numel() and bool(common.all()):
        assert torch.equal(o[0], t[0])  # indptr


@pytest.mark.parametrize("sparse_grad", [False, True])
def test_packed_bwd_matches_reference_cuda_and_dense(rs, ref, sparse_grad):
    N, C, W, H = 20_000, 2 if not sparse_grad else 1, 256, 192
    args = scene_args(N, C, W, H, seed=5, comp=True)
    means, _, quats, scales, _opac, vm, Ks = args[:7]
    t_f = ref.projection_ewa_3dgs_packed_fwd(*args, ref.PINHOLE)
    _indptr, b_ids, c_ids, g_ids, _radii, _m2, _d, conics, comps = t_f
    nnz = g_ids.numel()
    g = torch.Generator(device=DEV).manual_seed(0)
    v_m2 = torch.randn(nnz, 2, device=DEV, generator=g)
    v_d = torch.randn(nnz, device=DEV, generator=g)
    v_con = torch.randn(nnz, 3, device=DEV, generator=g) * 0.1
    v_comp = torch.randn(nnz, device=DEV, generator=g)
    head = (means, None, quats, scales, vm, Ks, W, H, 0.3)
    tail = (b_ids, c_ids, g_ids, conics, comps, v_m2, v_d, v_con, v_comp, True, sparse_grad)
    o = rs._C.projection_ewa_3dgs_packed_bwd(*head, rs._C.PINHOLE, *tail)
    t = ref.projection_ewa_3dgs_packed_bwd(*head, ref.PINHOLE, *tail)
    for k, name in ((0, "v_means"), (2, "v_quats"), (3, "v_scales"), (4, "v_viewmats")):
        assert o[k].shape == t[k].shape, name
        assert rel_err(o[k], t[k]) < 2e-3, name
    if sparse_grad:
        assert o[0].shape == (nnz, 3) and o[2].shape == (nnz, 4)
    # and against our dense backward fed with the scattered cotangents
    dense_f = rs._C.projection_ewa_3dgs_fused_fwd(*args, rs._C.PINHOLE)

    def scatter(v, tail_shape):
        out = torch.zeros((C, N) + tail_shape, device=DEV)
        out[c_ids, g_ids] = v
        return out

    radii_d = torch.zeros(C, N, 2, dtype=torch.int32, device=DEV)
    radii_d[c_ids, g_ids] = 1
    d = rs._C.projection_ewa_3dgs_fused_bwd(
        *head, rs._C.PINHOLE, radii_d, scatter(conics, (3,)), scatter(comps, ()), scatter(v_m2, (2,)), scatter(v_d, ()),
        scatter(v_con, (3,)), scatter(v_comp, ()), True)
    assert dense_f[0].shape == (C, N, 2)
    for k in (0, 2, 3):
        got = o[k]
        if sparse_grad:
            got = torch.zeros_like(d[k]).index_add_(0, g_ids, o[k])
        assert rel_err(got, d[k]) < 1e-5, k
    assert rel_err(o[4], d[4]) < 1e-4


@pytest.mark.parametrize("sparse_grad", [False, True])
def test_rasterization_packed_equals_unpacked(rs, sparse_grad):
    N, C, W, H = 30_000, 2 if not sparse_grad else 1, 200, 120
    s = synthetic_scene(17, N, s_max=0.1, K=4)
    vm, Ks = pinhole_cameras(C, W, H)
    results = []
    for packed in (False, True):
        leaves = [T(s[k]).requires_grad_() for k in ("means", "quats", "scales", "opacities", "colors")]
        img, alpha, meta = rs.rasterization(
            *leaves, T(vm), T(Ks), W, H, packed=packed, sparse_grad=sparse_grad and packed, rasterize_mode="antialiased",
            cluster_ids=T(s["cluster_ids"]), body_quats=T(s["body_quats"]), body_trans=T(s["body_trans"]),
            body_centers=T(s["body_centers"]))
        w = torch.linspace(0.5, 1.5, img.numel(), device=DEV).reshape(img.shape)
        ((img * w).sum() + alpha.sum()).backward()
        grads = [(x.grad.to_dense() if x.grad.is_sparse else x.grad) for x in leaves]
        results.append((img.detach(), alpha.detach(), grads, meta))
    (i0, a0, g0, m0), (i1, a1, g1, m1) = results
    assert m1["gaussian_ids"].dtype == torch.int64 and m1["means2d"].dim() == 2
    assert torch.equal(i0, i1) and torch.equal(a0, a1)
    for x, y in zip(g0, g1):
        assert rel_err(y, x) < 2e-3


@pytest.mark.parametrize("packed", [False, True])
def test_rasterization_with_no_visible_gaussian(rs, packed):
    """A camera that sees nothing (every Gaussian behind it): zero packed rows / zero intersections all the way through
    projection, binning and compositing, forward and backward -> background-only image, zero alpha, zero gradients."""
    N, W, H = 5_000, 96, 64
    s = synthetic_scene(2, N, s_max=0.1)
    vm, Ks = pinhole_cameras(2, W, H)
    means = (T(s["means"]) - torch.tensor([0.0, 0.0, 100.0], device=DEV)).requires_grad_()
    colors = T(s["colors"]).requires_grad_()
    bg = torch.tensor([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]], device=DEV)
    img, alpha, meta = rs.rasterization(means, T(s["quats"]), T(s["scales"]), T(s["opacities"]), colors, T(vm), T(Ks), W, H,
                                        packed=packed, backgrounds=bg)
    assert meta["flatten_ids"].numel() == 0 and int(meta["isect_offsets"].abs().max()) == 0
    assert float(alpha.abs().max()) == 0.0
    assert torch.equal(img, bg[:, None, None, :].expand(2, H, W, 3))
    if packed:
        assert meta["gaussian_ids"].numel() == 0 and meta["means2d"].shape == (0, 2)
    img.sum().backward()
    assert float(means.grad.abs().max()) == 0.0 and float(colors.grad.abs().max()) == 0.0
