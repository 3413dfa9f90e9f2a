"""Torch restatement of the real spherical harmonics up to degree 4 (directions normalised inside), differentiable:
the float64 checker of csrc/sh.cu.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/, never by the product package.
Follows the reference's `_eval_sh_bases_fast` / `_spherical_harmonics` (/root/reference/gsplat/cuda/_torch_impl.py:720-802,
804-830) and `spherical_harmonics` (gsplat/cuda/_wrapper.py:151-181); pinned against the reference's own output by
tests/golden/spherical_harmonics.npz (tests/golden/make_golden.py): tests/test_oracle_golden.py::test_sh_matches_reference.
"""
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)
_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
       -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def sh_bases(degree: int, dirs: Tensor) -> Tensor:
    """Real SH basis values [..., (degree+1)^2] at unit directions `dirs` [..., 3]."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, _C0)]
    if degree >= 1:
        out += [-_C1 * y, _C1 * z, -_C1 * x]
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
    if degree >= 3:
        out += [
            _C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
            _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
            _C3[6] * x * (xx - 3 * yy),
        ]
    if degree >= 4:
        out += [
            _C4[0] * xy * (xx - yy), _C4[1] * yz * (3 * xx - yy), _C4[2] * xy * (7 * zz - 1),
            _C4[3] * yz * (7 * zz - 3), _C4[4] * (zz * (35 * zz - 30) + 3), _C4[5] * xz * (7 * zz - 3),
            _C4[6] * (xx - yy) * (7 * zz - 1), _C4[7] * xz * (xx - 3 * yy),
            _C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)),
        ]
    return torch.stack(out, dim=-1)


def spherical_harmonics_torch(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    assert 0 <= degrees_to_use <= 4, degrees_to_use
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    batch_dims = dirs.shape[:-1]
    assert dirs.shape == batch_dims + (3,), dirs.shape
    assert coeffs.dim() == len(batch_dims) + 2 and coeffs.shape[:-2] == batch_dims and coeffs.shape[-1] == 3, coeffs.shape
    nb = (degrees_to_use + 1) ** 2
    bases = sh_bases(degrees_to_use, F.normalize(dirs, p=2, dim=-1))  # [..., nb]
    colors = (bases[..., None] * coeffs[..., :nb, :]).sum(dim=-2)
    if masks is not None:
        assert masks.shape == batch_dims, masks.shape
        colors = torch.where(masks[..., None], colors, torch.zeros_like(colors))
    return colors

