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
