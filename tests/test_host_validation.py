"""Host-side behaviour of the API mirror that needs no GPU: argument validation mirrors the reference's asserts
(gsplat/rendering.py:252-330, gsplat/cuda/_wrapper.py:349-392, 585-607), and CPU tensors are rejected loudly -- there is
no CPU or eager fallback behind `rasterization()`."""
import numpy as np
import pytest
import torch

from conftest import pinhole_cameras, synthetic_scene


def _inputs(N=64, C=2, W=32, H=24):
    s = synthetic_scene(0, N)
    vm, Ks = pinhole_cameras(C, W, H)
    t = {k: torch.from_numpy(v) for k, v in s.items()}
    return t, torch.from_numpy(vm), torch.from_numpy(Ks), W, H


def test_cpu_tensors_are_rejected_not_rendered(rs):
    t, vm, Ks, W, H = _inputs()
    with pytest.raises(RuntimeError, match="CUDA"):
        rs.rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], vm, Ks, W, H, packed=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        rs._C.projection_ewa_3dgs_packed_fwd(t["means"], None, t["quats"], t["scales"], t["opacities"], vm, Ks, W, H, 0.3,
                                             0.01, 1e10, 0.0, False, rs._C.PINHOLE)
    with pytest.raises(RuntimeError, match="CUDA"):
        rs.cgc_contrastive_clustering_loss(torch.randn(8, 8, 4), torch.ones(8, 8, dtype=torch.long))


@pytest.mark.parametrize("bad", ["quats", "opacities", "viewmats", "Ks", "colors"])
def test_shape_errors_match_the_reference_asserts(rs, bad):
    t, vm, Ks, W, H = _inputs()
    args = dict(means=t["means"], quats=t["quats"], scales=t["scales"], opacities=t["opacities"], colors=t["colors"],
                viewmats=vm, Ks=Ks)
    # a row short (per-Gaussian tensors, where any channel count would be legal) or a column short (the rest)
    args[bad] = args[bad][:-1] if bad in ("opacities", "colors") else args[bad][..., :-1]
    with pytest.raises(AssertionError):
        rs.rasterization(args["means"], args["quats"], args["scales"], args["opacities"], args["colors"], args["viewmats"],
                         args["Ks"], W, H)


def test_unsupported_options_are_refused_like_the_reference(rs):
    t, vm, Ks, W, H = _inputs()
    base = (t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], vm, Ks, W, H)
    with pytest.raises(NotImplementedError):
        rs.rasterization(*base, with_ut=True)
    with pytest.raises(AssertionError, match="only supported with"):
        rs.rasterization(*base, radial_coeffs=torch.zeros(2, 6))
    with pytest.raises(AssertionError):
        rs.rasterization(*base, render_mode="RGBD")
    with pytest.raises(AssertionError, match="tile_size"):
        rs.rasterization(*base, tile_size=8)
    with pytest.raises(AssertionError, match="body poses"):
        rs.rasterization(*base, body_quats=torch.zeros(2, 4), body_trans=torch.zeros(2, 3))
    with pytest.raises(AssertionError, match="sparse_grad"):
        rs.fully_fused_projection(t["means"], None, t["quats"], t["scales"], vm, Ks, W, H, packed=False, sparse_grad=True)


def test_c4_camera_assignment_is_a_balanced_partition():
    """tools/bench_c4.py: camera c of frame f -> rank (c + f) % world.  Every (frame, camera) pair is rendered exactly once
    and every rank meets every viewpoint equally often over `world` consecutive frames."""
    C, world, F = 8, 8, 16
    seen = np.zeros((F, C), int)
    per_rank_views = np.zeros((world, C), int)
    for rank in range(world):
        for f in range(F):
            for c in [c for c in range(C) if (c + f) % world == rank]:
                seen[f, c] += 1
                per_rank_views[rank, c] += 1
    assert (seen == 1).all() and (per_rank_views == F // world).all()
