"""Writes tests/golden/splats_small.ply and the reference's reading of it (tests/golden/splats_small_ref.npz) with the
reference's own `load_ply` (gsplat/utils.py:259-347, extracted by AST: the module imports `plyfile`, which is absent here,
but the function itself only needs numpy).  Run in the build container only."""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
fn = [n for n in ast.parse(open("/root/reference/gsplat/utils.py").read()).body
      if isinstance(n, ast.FunctionDef) and n.name == "load_ply"][0]
ns = {"np": np, "torch": torch}
exec(compile(ast.Module([fn], []), "reference_utils", "exec"), ns)

N, K = 37, 16  # sh degree 3: 45 f_rest columns, so the lexicographic column order differs from the numeric one
props = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(3 * (K - 1))] + \
        ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
rng = np.random.default_rng(7)
table = rng.normal(size=(N, len(props))).astype("<f4")
path = os.path.join(HERE, "splats_small.ply")
with open(path, "wb") as f:
    f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % N).encode())
    for p in props:
        f.write(f"property float {p}\n".encode())
    f.write(b"end_header\n")
    f.write(table.tobytes())
ref = ns["load_ply"](path, device="cpu")
np.savez_compressed(os.path.join(HERE, "splats_small_ref.npz"), **{k: v.numpy() for k, v in ref.items()})
print({k: tuple(v.shape) for k, v in ref.items()})
