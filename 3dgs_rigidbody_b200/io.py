"""Reading the reference's splat files: `load_ply` mirrors gsplat/utils.py:259-347, the loader main.py feeds
its renderer from.  Host-side I/O only; nothing here is on the per-frame path.

File format (written by the reference's `save_ply`): an ASCII header listing float32 vertex properties, then `N` packed
little-endian float32 records.  Semantics reproduced from the reference, including one quirk: the `scale_*`, `rot_*`,
`f_dc_*` and `f_rest_*` columns are taken in LEXICOGRAPHIC order of their names (`sorted()` on strings, so `f_rest_10` comes
before `f_rest_2`), not in numeric order, before being reshaped to [N, 3, K] and transposed to [N, K, 3].
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch


def load_ply(path: str, device: str = "cuda") -> Dict[str, torch.Tensor]:
    """-> {"means" [N,3], "opacities" [N], "scales" [N,3], "quats" [N,4], "sh0" [N,1,3], "shN" [N,K-1,3]} float32 on
    `device` (raw file values: no activation applied, as in the reference)."""
    with open(path, "rb") as fh:
        blob = fh.read()
    end = blob.index(b"end_header")
    body_at = blob.index(b"\n", end) + 1
    names: List[str] = []
    n_points = 0
    for line in blob[:end].decode("utf-8").splitlines():
        line = line.strip()
        if line.startswith("element vertex"):
            n_points = int(line.split()[-1])
        elif line.startswith("property"):
            names.append(line.split()[-1])  # every property of these files is float32 (utils.py:289-294)
    table = np.frombuffer(blob, dtype="<f4", count=n_points * len(names), offset=body_at).reshape(n_points, len(names))
    col = {n: i for i, n in enumerate(names)}

    def columns(prefix: str) -> np.ndarray:
        picked = sorted(n for n in names if n.startswith(prefix))  # lexicographic, on purpose
        return table[:, [col[n] for n in picked]] if picked else np.empty((n_points, 0), np.float32)

    def to_k3(flat: np.ndarray) -> np.ndarray:  # [N, 3*K] channel-major -> [N, K, 3]
        return flat.reshape(n_points, 3, -1).transpose(0, 2, 1)

    rest = columns("f_rest_")
    out = {
        "means": table[:, [col["x"], col["y"], col["z"]]],
        "opacities": table[:, col["opacity"]],
        "scales": columns("scale_"),
        "quats": columns("rot_"),
        "sh0": to_k3(columns("f_dc_")),
        "shN": to_k3(rest) if rest.shape[1] > 0 else np.zeros((n_points, 0, 3), np.float32),
    }
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device) for k, v in out.items()}
