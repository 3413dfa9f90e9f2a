"""Host-side data contract around the renderer: cluster_groups archive -> dense cluster ids, per-body rigid
parameters, physics pose stream -> pose tables.  CPU only."""
import importlib

import numpy as np
import torch

from conftest import synthetic_scene


def _mods():
    return importlib.import_module("3dgs_rigidbody_b200.rigid"), importlib.import_module("3dgs_rigidbody_b200.torch_ref")


def test_cluster_groups_archive_roundtrip(rs, tmp_path):
    rigid, _ = _mods()
    N = 1000
    rng = np.random.default_rng(0)
    labels = rng.integers(-1, 4, size=N)
    # the producer's format: np.savez_compressed(**{str(object id): indices, "background": indices})
    groups = {str(10 * (k + 1)): np.where(labels == k)[0] for k in range(4)}
    groups["background"] = np.where(labels == -1)[0]
    path = str(tmp_path / "cluster_groups.npy.npz")
    np.savez_compressed(path, **groups)
    ids, names = rigid.load_cluster_groups(path, N)
    assert ids.dtype == torch.int32 and ids.shape == (N,)
    assert names == {0: "10", 1: "20", 2: "30", 3: "40"}
    assert np.array_equal(ids.numpy(), labels)
    # Gaussians listed nowhere are static too
    del groups["background"]
    ids2, _ = rigid.cluster_ids_from_groups(groups, N)
    assert torch.equal(ids, ids2)
    assert rigid.instance_mask_path("/d", "images/frame_0007.JPG") == "/d/masks/instance_ids_npy/frame_0007_instance_id.npy"


def test_body_properties_of_a_known_body(rs):
    rigid, _ = _mods()
    # body 0: eight equal point masses on the corners of a 2 x 4 x 6 box centred at (1, 2, 3); body 1: empty
    corners = np.array([[sx, 2 * sy, 3 * sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) + [1, 2, 3]
    means = torch.from_numpy(np.concatenate([corners, [[9, 9, 9]]]).astype(np.float32))
    scales = torch.full((9, 3), 0.5)
    opac = torch.full((9,), 0.8)
    ids = torch.tensor([0] * 8 + [-1], dtype=torch.int32)
    mass, com, inertia = rigid.body_properties(means, scales, opac, ids, 2)
    m = 0.8 * 0.125
    assert torch.allclose(mass, torch.tensor([8 * m, 0.0]))
    assert torch.allclose(com[0], torch.tensor([1.0, 2.0, 3.0]))
    want = 8 * m * torch.diag(torch.tensor([4.0 + 9.0, 1.0 + 9.0, 1.0 + 4.0]))
    assert torch.allclose(inertia[0], want, atol=1e-5) and float(inertia[1].abs().max()) == 0.0
    assert torch.allclose(com[0], rigid.body_centers(means, ids, 2)[0])


def test_pose_stream_reproduces_apply_transform_about_the_body_centre(rs, tmp_path):
    rigid, tr = _mods()
    s = synthetic_scene(4, 500, K=3)
    means, quats = torch.from_numpy(s["means"]), torch.from_numpy(s["quats"])
    ids = torch.from_numpy(s["cluster_ids"])
    centers = rigid.body_centers(means, ids, 3)
    F = 5
    g = torch.Generator().manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(F, 3, 4, generator=g), dim=-1)
    pos = centers[None] + 0.2 * torch.randn(F, 3, 3, generator=g)  # where the physics engine puts each body origin
    stream = rigid.PoseStream(torch.cat([q, pos], -1), centers)
    path = str(tmp_path / "poses.npz")
    stream.save(path)
    again = rigid.PoseStream.load(path)
    assert len(again) == F and again.num_bodies == 3
    bq, bt = again.frame(2)
    poses = rigid.make_rigid(ids, bq, bt, again.centers)
    moved, _ = tr.apply_rigid_torch(means, quats, poses)
    # main.py:210-222 for one body k: (x - c) R^T + c + t, with t = p - c
    for k in range(3):
        R = tr.normalized_quat_to_rotmat(q[2, k])
        sel = ids == k
        want = (means[sel] - centers[k]) @ R.T + pos[2, k]
        assert torch.allclose(moved[sel], want, atol=1e-5)
    assert torch.equal(moved[ids < 0], means[ids < 0])
