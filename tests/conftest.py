import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def rs():
    """The product package (its directory name starts with a digit, hence importlib)."""
    import __graft_entry__ as ge

    ge.build()
    return importlib.import_module("3dgs_rigidbody_b200")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle


REF_SO = os.path.join(ROOT, "oracle", "_ref", "gsplat_ref_cuda.so")


@pytest.fixture(scope="session")
def ref():
    """The reference's own compiled CUDA extension (oracle/build_ref.py): the GPU-side checker."""
    import importlib.util

    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/gsplat_ref_cuda.so not built (needs /root/reference at build time)")
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location("gsplat_ref_cuda", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def refpy(rs):
    """baseline/install_ref.py: importer of the UNMODIFIED reference python (gsplat package, main.py functions) from
    baseline/_ref -- git-ignored, installed where /root/reference exists, travels to the GPU box."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("install_ref", os.path.join(ROOT, "baseline", "install_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not mod.available():
        try:
            mod.install()
        except FileNotFoundError:
            pytest.skip("baseline/_ref not installed (needs /root/reference at build time)")
    # make `gsplat` importable right away (our operator module stands in for its extension; tests that want the
    # reference's own kernels call set_backend(ref)): reference_functions() needs gsplat.utils whatever test runs first
    mod.load_reference(rs._C)
    return mod


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def synthetic_scene(seed, N, K=0, s_max=0.05, z_off=8.0, spread=1.0):
    """Seeded synthetic Gaussians (value distributions of gsplat/_helper.py:49-53, examples/image_fitting.py:35-68)."""
    rng = np.random.default_rng(seed)
    means = (rng.normal(size=(N, 3)) * spread).astype(np.float32)
    means[:, 2] += z_off
    quats = rng.normal(size=(N, 4)).astype(np.float32)
    scales = (rng.random((N, 3)) * s_max).astype(np.float32)
    opacities = rng.random(N).astype(np.float32)
    colors = rng.random((N, 3)).astype(np.float32)
    out = dict(means=means, quats=quats, scales=scales, opacities=opacities, colors=colors)
    if K > 0:
        out["cluster_ids"] = rng.integers(-1, K, size=N).astype(np.int32)
        out["body_quats"] = rng.normal(size=(K, 4)).astype(np.float32)
        out["body_trans"] = (rng.normal(size=(K, 3)) * 0.3).astype(np.float32)
        out["body_centers"] = (rng.normal(size=(K, 3)) * 0.5 + np.array([0, 0, z_off])).astype(np.float32)
    return out


def pinhole_cameras(C, W, H, f_scale=0.8):
    viewmats = np.tile(np.eye(4, dtype=np.float32), (C, 1, 1))
    for c in range(C):
        ang = 0.2 * c
        viewmats[c, 0, 0] = np.cos(ang)
        viewmats[c, 0, 2] = np.sin(ang)
        viewmats[c, 2, 0] = -np.sin(ang)
        viewmats[c, 2, 2] = np.cos(ang)
        viewmats[c, 0, 3] = 0.3 * c
    Ks = np.tile(np.array([[f_scale * W, 0, W / 2], [0, f_scale * W, H / 2], [0, 0, 1]], np.float32), (C, 1, 1))
    return viewmats, Ks


def apply_transform_per_body(refpy, splats, cluster_ids, body_quats, body_trans):
    """The reference's animation step: apply_transform() once per body on that body's Gaussians (main.py:366-400, 183-228).
    Returns the moved (means, quats) and the per-body pivots it used (means.mean(dim=0) of the body, main.py:210)."""
    import torch

    fns = refpy.reference_functions("main.py", ["apply_transform", "quat_multiply"])
    means, quats = splats["means"].clone(), splats["quats"].clone()
    K = body_quats.shape[0]
    centers = torch.zeros(K, 3, device=means.device)
    for k in range(K):
        idx = torch.nonzero(cluster_ids == k).squeeze(-1)
        if idx.numel() == 0:
            continue
        part = {"means": splats["means"][idx], "quats": splats["quats"][idx]}
        centers[k] = part["means"].mean(dim=0)
        moved = fns["apply_transform"](part, body_trans[k], body_quats[k])
        means[idx] = moved["means"]
        quats[idx] = moved["quats"]
    return means, quats, centers
