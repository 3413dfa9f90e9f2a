#!/usr/bin/env python
"""Head-to-head of the hand-written one-sweep radix sort (librigidsplat.so: rs_radix_sort_pairs / rs_radix_sort_pairs32)
against cub::DeviceRadixSort::SortPairs -- the call the reference makes (gsplat/cuda/csrc/IntersectTile.cu:296-339) -- on
IDENTICAL keys, on the problem shapes of the path:
    1 M / 5 M / 20 M pairs, 32-bit keys (depth bits; tile keys over 14 / 17 bits)
    5 M / 20 M pairs, 64-bit keys over 46 bits (1080p, 1 image) and 51 bits (4K, 8 images): the reference's own sort
Both results are checked against torch.sort(stable=True).  Prints a table and one JSON line (kept under profiles/).

    python tools/sort_bench.py [path/to/librigidsplat_variant.so ...]
The cub arm is built by tools/build_cub_ref.sh (test-side, never linked into the product)."""
import ctypes
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = importlib.import_module("3dgs_rigidbody_b200._lib")
CUB = os.path.join(ROOT, "tools", "_tmp", "libcubsort.so")
DEV = "cuda:0"


def make_keys(n, bits, kind, wide):
    g = torch.Generator(device=DEV).manual_seed(0)
    if kind == "random":
        lo = torch.randint(0, 1 << min(bits, 31), (n,), dtype=torch.int64, device=DEV, generator=g)
        if bits > 31:
            lo = lo | (torch.randint(0, 1 << (bits - 31), (n,), dtype=torch.int64, device=DEV, generator=g) << 31)
        keys = lo
    else:  # tile-like: runs of 5 consecutive keys (a splat's tile row), as the emission produces them
        base = torch.randint(0, (1 << bits) - 8, (n // 5 + 1,), dtype=torch.int64, device=DEV, generator=g)
        keys = (base[:, None] + torch.arange(5, dtype=torch.int64, device=DEV)[None]).reshape(-1)[:n].contiguous()
    return keys if wide else keys.to(torch.int32)


def time_it(fn, reset, reps=20):
    ts = []
    for r in range(reps + 3):
        reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run_ours(lib, keys, bits, wide):
    n = keys.numel()
    vals = torch.arange(n, dtype=torch.int32, device=DEV)
    ka, kb, va, vb = torch.empty_like(keys), torch.empty_like(keys), torch.empty_like(vals), torch.empty_like(vals)
    ws_bytes = lib.rs_radix_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    a = _lib.rs_sort_args()
    a.n, a.n_dev, a.begin_bit, a.end_bit = n, None, 0, bits
    a.keys_a, a.keys_b, a.vals_a, a.vals_b = ka.data_ptr(), kb.data_ptr(), va.data_ptr(), vb.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    res = ctypes.c_int32(0)
    a.result_in_b = ctypes.addressof(res)
    fn_sort = lib.rs_radix_sort_pairs if wide else lib.rs_radix_sort_pairs32
    stream = torch.cuda.current_stream().cuda_stream

    def reset():
        ka.copy_(keys)
        va.copy_(vals)

    def fn():
        st = fn_sort(ctypes.byref(a), stream)
        assert st == 0, lib.rs_last_error()

    ms = time_it(fn, reset)
    return ms, (vb if res.value else va)


def run_cub(cub, keys, bits, wide):
    n = keys.numel()
    vals = torch.arange(n, dtype=torch.int32, device=DEV)
    ka, kb, va, vb = torch.empty_like(keys), torch.empty_like(keys), torch.empty_like(vals), torch.empty_like(vals)
    fn_sort = cub.cub_sort_pairs_u64 if wide else cub.cub_sort_pairs_u32
    fn_sort.restype = ctypes.c_int
    fn_sort.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)] + [ctypes.c_void_p] * 4 + [
        ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    tb = ctypes.c_size_t(0)
    sel = ctypes.c_int(0)
    stream = torch.cuda.current_stream().cuda_stream
    assert fn_sort(None, ctypes.byref(tb), ka.data_ptr(), kb.data_ptr(), va.data_ptr(), vb.data_ptr(), n, 0, bits, stream,
                   ctypes.byref(sel)) == 0
    temp = torch.empty(max(tb.value, 16), dtype=torch.uint8, device=DEV)

    def reset():
        ka.copy_(keys)
        va.copy_(vals)

    def fn():
        assert fn_sort(temp.data_ptr(), ctypes.byref(tb), ka.data_ptr(), kb.data_ptr(), va.data_ptr(), vb.data_ptr(), n, 0,
                       bits, stream, ctypes.byref(sel)) == 0

    ms = time_it(fn, reset)
    return ms, (vb if sel.value else va)


def main():
    libs = [_lib.LIB_PATH] + sys.argv[1:]
    cub = ctypes.CDLL(CUB) if os.path.exists(CUB) else None
    cases = [  # (n, key bits sorted, key kind, 64-bit keys)
        (1_000_000, 32, "random", False), (5_000_000, 14, "tile", False), (5_000_000, 14, "random", False),
        (20_000_000, 17, "tile", False), (20_000_000, 32, "random", False),
        (5_000_000, 46, "random", True), (20_000_000, 46, "random", True), (20_000_000, 51, "random", True),
    ]
    rows = []
    for n, bits, kind, wide in cases:
        keys = make_keys(n, bits, kind, wide)
        mask = (1 << bits) - 1
        _, want = torch.sort((keys.long() & mask), stable=True)
        kb_bytes = 8 if wide else 4
        passes = (bits + 7) // 8
        row = {"n": n, "bits": bits, "keys": kind, "key_bytes": kb_bytes, "passes_8bit": passes}
        for path in libs:
            lib = ctypes.CDLL(path)
            for name, (res, args) in _lib._PROTOS.items():
                if hasattr(lib, name):
                    getattr(lib, name).restype = res
                    getattr(lib, name).argtypes = args
            ms, out = run_ours(lib, keys, bits, wide)
            tag = "ours" if path == _lib.LIB_PATH else os.path.basename(path)
            row[tag + "_us"] = round(ms * 1e3, 1)
            row[tag + "_correct"] = bool(torch.equal(out.long(), want))
            row[tag + "_GBps_moved"] = round(n * (kb_bytes + passes * 2 * (kb_bytes + 4)) / (ms * 1e-3) / 1e9, 1)
        if cub is not None:
            ms, out = run_cub(cub, keys, bits, wide)
            row["cub_us"] = round(ms * 1e3, 1)
            row["cub_correct"] = bool(torch.equal(out.long(), want))
            row["ours_over_cub"] = round(row["ours_us"] / row["cub_us"], 3)
        rows.append(row)
        print(" ".join(f"{k}={v}" for k, v in row.items()), flush=True)
    print(json.dumps({"tool": "sort_bench", "gpu": torch.cuda.get_device_name(0), "rows": rows}))


if __name__ == "__main__":
    main()
