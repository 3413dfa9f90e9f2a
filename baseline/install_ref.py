"""Installs the UNMODIFIED reference python (gsplat package + main.py) into baseline/_ref/ -- TEST / BENCH INFRASTRUCTURE.

The reference tree has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` cannot work
("neither setup.py nor pyproject.toml found"); this script does what that install would do for a pure-python package:
it copies the package's .py files (not its CUDA sources: the compiled reference extension is oracle/_ref/gsplat_ref_cuda.so,
built by oracle/build_ref.py) and main.py, byte for byte, into baseline/_ref/, which is git-ignored (never part of the repo's
history) but travels to the GPU box with the gpurun snapshot.  Used by

  * tests/test_gpu_dropin.py   the reference's OWN `gsplat.rendering.rasterization()` / `_wrapper.py` running on top of
                               `3dgs_rigidbody_b200._C` (the drop-in claim of INTEGRATION.md), compared with the same
                               python on the reference's own compiled extension;
  * tests (parity)             the reference's `apply_transform` / `quat_multiply` (main.py:173-228), extracted by AST,
                               as the comparison arm of the whole-frame tests;
  * bench.py                   the `ref_cuda` leg: the reference's python + kernels timed on the same B200 and frames.

Only runs where /root/reference exists (the build container).  `load_reference()` below is the importer the tests use.
"""
import ast
import importlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def install(force: bool = False) -> str:
    marker = os.path.join(DEST, "gsplat", "rendering.py")
    if os.path.exists(marker) and not force:
        return DEST
    if not os.path.isdir(SRC):
        raise FileNotFoundError(f"{SRC} not present and {DEST} not installed")
    for dirpath, dirnames, files in os.walk(os.path.join(SRC, "gsplat")):
        dirnames[:] = [d for d in dirnames if d not in ("csrc", "include", "__pycache__")]
        rel = os.path.relpath(dirpath, SRC)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(DEST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(dirpath, f), os.path.join(DEST, rel, f))
    os.makedirs(os.path.join(DEST, "_scripts"), exist_ok=True)
    shutil.copyfile(os.path.join(SRC, "main.py"), os.path.join(DEST, "_scripts", "main.py"))
    for rel in ("examples/load_identity_encodings.py", "examples/utils.py"):
        if os.path.exists(os.path.join(SRC, rel)):
            shutil.copyfile(os.path.join(SRC, rel), os.path.join(DEST, "_scripts", os.path.basename(rel)))
    return DEST


def available() -> bool:
    return os.path.exists(os.path.join(DEST, "gsplat", "rendering.py"))


def load_reference(backend):
    """Imports the reference's `gsplat` package from baseline/_ref with `backend` standing in for its compiled extension
    (`from gsplat import csrc as _C`, gsplat/cuda/_backend.py:170).  Returns the `gsplat` module; the operator module can be
    swapped later with `set_backend()` because the reference resolves `_C` lazily on every call (_wrapper.py:12-19)."""
    if not available():
        raise FileNotFoundError(f"{DEST} not installed (run baseline/install_ref.py where /root/reference exists)")
    if "plyfile" not in sys.modules:
        try:
            importlib.import_module("plyfile")
        except ImportError:  # the only import blocker of the package (gsplat/utils.py:6); never called on this path
            sys.modules["plyfile"] = types.SimpleNamespace(PlyData=None, PlyElement=None)
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    sys.modules["gsplat.csrc"] = backend
    gsplat = importlib.import_module("gsplat")
    set_backend(backend)
    return gsplat


def set_backend(backend) -> None:
    be = importlib.import_module("gsplat.cuda._backend")
    be._C = backend
    sys.modules["gsplat.csrc"] = backend


def reference_functions(script: str, names):
    """Top-level functions of a reference script (e.g. main.py's apply_transform / quat_multiply), extracted by AST so that
    the script's argparse / file-loading body never runs.  Returns {name: function}."""
    path = os.path.join(DEST, "_scripts", script)
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    import torch
    from typing import Dict, List, Optional, Tuple

    ns = {"torch": torch, "Tensor": torch.Tensor, "Dict": Dict, "List": List, "Optional": Optional, "Tuple": Tuple,
          "F": torch.nn.functional}
    if "normalized_quat_to_rotmat" not in names:
        ns["normalized_quat_to_rotmat"] = importlib.import_module("gsplat.utils").normalized_quat_to_rotmat
    exec(compile(mod, path, "exec"), ns)
    return {n: ns[n] for n in names}


def reference_statements(script: str, want):
    """Statements of a reference script selected by `want(node) -> bool` over every statement of the module (including the
    ones nested in functions and `if __name__` blocks), in source order, compiled into ONE code object.  For code the
    reference keeps inline instead of in a function -- e.g. the writer of the `cluster_groups` archive
    (examples/load_identity_encodings.py:478-486, 566-568).  Returns (code, [source line numbers])."""
    path = os.path.join(DEST, "_scripts", script)
    tree = ast.parse(open(path).read())
    picked = []

    def visit(body):
        for node in body:
            if want(node):
                picked.append(node)
                continue
            for field in ("body", "orelse", "finalbody"):
                sub = getattr(node, field, None)
                if isinstance(sub, list) and sub and isinstance(sub[0], ast.stmt):
                    visit(sub)

    visit(tree.body)
    mod = ast.Module(body=picked, type_ignores=[])
    return compile(mod, path, "exec"), [n.lineno for n in picked]


if __name__ == "__main__":
    print(install(force="-f" in sys.argv))
