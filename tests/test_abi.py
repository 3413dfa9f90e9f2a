"""The C-ABI library loads and exports every symbol include/rigidsplat.h declares (no compute: runs without a GPU)."""
import ctypes
import importlib
import os
import re
import subprocess


def test_header_symbols_exported(rs):
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    declared = _lib.declared_functions()
    assert len(declared) >= 18, declared
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rigidsplat.h but not exported"
    # every bound prototype is a declared function and vice versa
    assert sorted(_lib.EXPORTS) == declared


def test_struct_layouts_match(rs):
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    lib = _lib.load()
    order = [_lib.rs_project_fwd_args, _lib.rs_project_bwd_args, _lib.rs_isect_args, _lib.rs_sort_args,
             _lib.rs_raster_fwd_args, _lib.rs_raster_bwd_args, _lib.rs_frame_args, _lib.rs_rigid_t,
             _lib.rs_isect_sorted_args, _lib.rs_sh_args, _lib.rs_project_packed_fwd_args, _lib.rs_exchange_args, _lib.rs_cgc_args,
             _lib.rs_seghead_args, _lib.rs_exchange_grad_args]
    for which, st in enumerate(order):
        assert lib.rs_sizeof_args(which) == ctypes.sizeof(st), st.__name__
    assert lib.rs_sizeof_args(99) == 0
    header_version = int(re.search(r"#define\s+RS_ABI_VERSION\s+(\d+)", open(_lib.HEADER).read()).group(1))
    assert lib.rs_abi_version() == header_version >= 18


def test_host_only_entry_points(rs):
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    lib = _lib.load()
    assert lib.rs_isect_num_blocks(0) == 0
    assert lib.rs_isect_num_blocks(1) == 1
    assert lib.rs_isect_num_blocks(1024) == 1
    assert lib.rs_isect_num_blocks(1025) == 2
    assert lib.rs_radix_sort_workspace_bytes(1) > 0
    assert lib.rs_radix_sort_workspace_bytes(10_000_000) >= 256 * 4 * ((10_000_000 + 4095) // 4096)
    n = lib.rs_frame_workspace_bytes(1, 1_000_000, 1920, 1080, 16, 3, 16_000_000)
    assert n > 16_000_000 * 24
    assert lib.rs_frame_workspace_bytes(0, 10, 16, 16, 16, 3, 10) == 0


def test_argument_errors_are_reported(rs):
    """Bad arguments fail loudly with a message (mirrors TORCH_CHECK -> RuntimeError); still no GPU work."""
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    lib = _lib.load()
    a = _lib.rs_raster_fwd_args()
    a.tile_size, a.channels, a.I, a.image_width, a.image_height, a.tile_width, a.tile_height = 8, 3, 1, 16, 16, 2, 2
    assert lib.rs_raster_fwd(ctypes.byref(a), None) != 0
    assert b"tile_size" in lib.rs_last_error()
    a.tile_size, a.channels, a.tile_width, a.tile_height = 16, 0, 1, 1
    assert lib.rs_raster_fwd(ctypes.byref(a), None) != 0
    assert b"Unsupported number of color channels" in lib.rs_last_error()
    p = _lib.rs_project_fwd_args()
    p.B, p.C, p.N, p.camera_model = 1, 1, 4, 3  # ftheta
    assert lib.rs_project_fwd(ctypes.byref(p), None) != 0
    assert b"camera model" in lib.rs_last_error()


def test_library_is_sm100a_native(rs):
    """The shipped .so holds sm_100a SASS only (no multi-arch dispatch, no PTX-JIT fallback for other GPUs)."""
    _lib = importlib.import_module("3dgs_rigidbody_b200._lib")
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        return  # cuobjdump unavailable: nothing to check
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_in_product():
    """The product package never imports, links or executes anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "3dgs_rigidbody_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/ ", ""), f"{f} mentions the oracle"
