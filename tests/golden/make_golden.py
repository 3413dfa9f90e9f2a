"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own python code in this container.

The reference repository has no tests or golden vectors of its own, so the fixtures are outputs of
the reference itself on seeded synthetic inputs:
  * main.py `quat_multiply` / `apply_transform` (extracted by AST -- main.py cannot be imported: it pulls in a COLMAP
    trainer and third-party packages that are absent here), one call per body, exactly as its animation loop does;
  * gsplat/cuda/_torch_impl.py `_quat_scale_to_covar_preci`, `_fully_fused_projection` (+ autograd gradients),
    `_isect_tiles`, `_isect_offset_encode`, `_spherical_harmonics`.
What the reference cannot run on a CPU (`_rasterize_to_pixels` needs its CUDA extension and nerfacc) is not in here;
compositing parity is pinned on the GPU box against the reference CUDA extension (oracle/_ref/).

Run (only where /root/reference exists):   python tests/golden/make_golden.py
The fixtures travel with the repo; /root/reference is never read at test time.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.modules.setdefault("plyfile", types.SimpleNamespace(PlyData=None, PlyElement=None))
    sys.path.insert(0, REF)
    import gsplat  # noqa: F401
    from gsplat.cuda import _torch_impl
    from gsplat.utils import normalized_quat_to_rotmat

    return _torch_impl, normalized_quat_to_rotmat


def extract_main_functions(normalized_quat_to_rotmat):
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("quat_multiply", "apply_transform")]
    mod = ast.Module(body=wanted, type_ignores=[])
    from typing import Dict

    ns = {"torch": torch, "Tensor": torch.Tensor, "Dict": Dict, "normalized_quat_to_rotmat": normalized_quat_to_rotmat}
    exec(compile(mod, "main.py", "exec"), ns)
    return ns["quat_multiply"], ns["apply_transform"]


def scene(seed, N, K, s_max=0.05):
    g = torch.Generator().manual_seed(seed)
    means = torch.randn(N, 3, generator=g)
    means[:, 2] += 8.0
    quats = torch.randn(N, 4, generator=g)
    scales = torch.rand(N, 3, generator=g) * s_max
    opacities = torch.rand(N, generator=g)
    colors = torch.rand(N, 3, generator=g)
    cluster_ids = torch.randint(0, K, (N,), generator=g, dtype=torch.int32)
    body_quats = torch.randn(K, 4, generator=g)
    body_trans = torch.randn(K, 3, generator=g) * 0.3
    return means, quats, scales, opacities, colors, cluster_ids, body_quats, body_trans


def cameras(C, W, H):
    viewmats = torch.eye(4).repeat(C, 1, 1)
    for c in range(C):
        ang = 0.15 * c
        viewmats[c, 0, 0] = np.cos(ang)
        viewmats[c, 0, 2] = np.sin(ang)
        viewmats[c, 2, 0] = -np.sin(ang)
        viewmats[c, 2, 2] = np.cos(ang)
        viewmats[c, 0, 3] = 0.2 * c
    Ks = torch.tensor([[0.8 * W, 0.0, W / 2], [0.0, 0.8 * W, H / 2], [0.0, 0.0, 1.0]]).repeat(C, 1, 1)
    return viewmats, Ks


def main():
    ti, nq2r = import_reference()
    quat_multiply, apply_transform = extract_main_functions(nq2r)
    torch.manual_seed(42)

    # ---- c1: 10k Gaussians, 2 rigid clusters, one 256x256 frame ------------------------------------------------------
    N, K, W, H = 10000, 2, 256, 256
    means, quats, scales, opacities, colors, cluster_ids, body_quats, body_trans = scene(42, N, K)
    viewmats, Ks = cameras(1, W, H)
    # reference animation step: one apply_transform() per body on that body's Gaussians (main.py:280-297, 366-400)
    t_means, t_quats = means.clone(), quats.clone()
    centers = torch.zeros(K, 3)
    for k in range(K):
        sel = cluster_ids == k
        splats = {"means": means[sel], "quats": quats[sel]}
        centers[k] = splats["means"].mean(dim=0)
        out = apply_transform(splats, body_trans[k], body_quats[k])
        t_means[sel], t_quats[sel] = out["means"], out["quats"]
    covars, _ = ti._quat_scale_to_covar_preci(t_quats, scales, compute_covar=True, compute_preci=False)
    radii, means2d, depths, conics, comps = ti._fully_fused_projection(
        t_means, covars, viewmats, Ks, W, H, calc_compensations=True)
    tw, th = W // 16, H // 16
    tpg, isect_ids, flatten_ids = ti._isect_tiles(means2d, radii, depths, 16, tw, th)
    offsets = ti._isect_offset_encode(isect_ids, 1, tw, th)
    np.savez_compressed(
        os.path.join(OUT, "c1_rigid_project_isect.npz"),
        means=means.numpy(), quats=quats.numpy(), scales=scales.numpy(), opacities=opacities.numpy(),
        colors=colors.numpy(), cluster_ids=cluster_ids.numpy(), body_quats=body_quats.numpy(),
        body_trans=body_trans.numpy(), body_centers=centers.numpy(), viewmats=viewmats.numpy(), Ks=Ks.numpy(),
        width=W, height=H,
        ref_means=t_means.numpy(), ref_quats=t_quats.numpy(), ref_covars=covars.numpy(),
        ref_radii=radii.numpy(), ref_means2d=means2d.numpy(), ref_depths=depths.numpy(), ref_conics=conics.numpy(),
        ref_compensations=comps.numpy(), ref_tiles_per_gauss=tpg.numpy(), ref_isect_ids=isect_ids.numpy(),
        ref_flatten_ids=flatten_ids.numpy(), ref_offsets=offsets.numpy(),
    )

    # ---- projection forward + autograd gradients: 3 camera models, 2 cameras, batch of 2 -----------------------------
    N, W, H = 1500, 200, 120
    fix = {}
    for model in ("pinhole", "ortho", "fisheye"):
        m, q, s, o, c, _, _, _ = scene(7, 2 * N, 2, s_max=0.1)
        m, q, s = m.reshape(2, N, 3), q.reshape(2, N, 4), s.reshape(2, N, 3)
        if model == "ortho":
            # keep the projected footprint comparable to eps2d: the reference CUDA backward divides by
            # (compensation + 1e-6) (Utils.cuh:406) which departs from autograd when compensation -> 0
            s = s * 0.5
        vm, K3 = cameras(2, W, H)
        vm, K3 = vm[None].repeat(2, 1, 1, 1), K3[None].repeat(2, 1, 1, 1)
        if model == "ortho":
            K3 = K3.clone()
            K3[..., 0, 0] = 20.0
            K3[..., 1, 1] = 20.0
        m.requires_grad_(True)
        q.requires_grad_(True)
        s.requires_grad_(True)
        cov, _ = ti._quat_scale_to_covar_preci(q, s, compute_covar=True, compute_preci=False)
        r, m2, d, con, comp = ti._fully_fused_projection(m, cov, vm, K3, W, H, calc_compensations=True,
                                                         camera_model=model)
        g = torch.Generator().manual_seed(3)
        v_m2, v_d, v_con = torch.randn(m2.shape, generator=g), torch.randn(d.shape, generator=g), \
            torch.randn(con.shape, generator=g) * 0.1
        v_comp = torch.randn(comp.shape, generator=g)
        sel = (r > 0).all(-1)
        loss = (m2 * v_m2 * sel[..., None]).sum() + (d * v_d * sel).sum() + (con * v_con * sel[..., None]).sum() + \
            (comp * v_comp * sel).sum()
        loss.backward()
        for k_, v_ in dict(means=m, quats=q, scales=s, viewmats=vm, Ks=K3, radii=r, means2d=m2, depths=d, conics=con,
                           compensations=comp, v_means2d=v_m2, v_depths=v_d, v_conics=v_con, v_compensations=v_comp,
                           g_means=m.grad, g_quats=q.grad, g_scales=s.grad).items():
            fix[f"{model}_{k_}"] = v_.detach().numpy()
    fix["width"], fix["height"] = W, H
    np.savez_compressed(os.path.join(OUT, "projection_fwd_bwd.npz"), **fix)

    # ---- spherical harmonics -------------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    dirs = torch.randn(500, 3, generator=g)
    coeffs = torch.randn(500, 25, 3, generator=g)
    sh = {"dirs": dirs.numpy(), "coeffs": coeffs.numpy()}
    for deg in range(5):
        sh[f"deg{deg}"] = ti._spherical_harmonics(deg, dirs, coeffs).numpy()
    np.savez_compressed(os.path.join(OUT, "spherical_harmonics.npz"), **sh)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
