#!/usr/bin/env python
"""Micro-benchmark of the one-sweep radix sort on the two problem shapes of a c2 frame: 1 M (depth bits, id) pairs over
32 bits and 5 M (tile key, id) pairs over 14 bits.  Checks the result against torch.sort(stable=True).
    python tools/sort_bench.py [path/to/librigidsplat.so ...]     (extra libraries = compile-time variants to compare)"""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = importlib.import_module("3dgs_rigidbody_b200._lib")


def run(lib, n, end_bit, reps=20, kind="random"):
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    if kind == "random":
        keys = torch.randint(0, 1 << min(end_bit, 31), (n,), dtype=torch.int32, device=dev, generator=g)
    else:  # tile-like: runs of consecutive keys
        base = torch.randint(0, (1 << end_bit) - 8, (n // 5 + 1,), dtype=torch.int32, device=dev, generator=g)
        keys = (base[:, None] + torch.arange(5, dtype=torch.int32, device=dev)[None]).reshape(-1)[:n].contiguous()
    vals = torch.arange(n, dtype=torch.int32, device=dev)
    ka, kb, va, vb = torch.empty_like(keys), torch.empty_like(keys), torch.empty_like(vals), torch.empty_like(vals)
    ws_bytes = lib.rs_radix_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    a = _lib.rs_sort_args()
    a.n, a.n_dev, a.begin_bit, a.end_bit = n, None, 0, end_bit
    a.keys_a, a.keys_b, a.vals_a, a.vals_b = ka.data_ptr(), kb.data_ptr(), va.data_ptr(), vb.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    res = ctypes.c_int32(0)
    a.result_in_b = ctypes.addressof(res)
    ts = []
    for r in range(reps + 3):
        ka.copy_(keys); va.copy_(vals)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = lib.rs_radix_sort_pairs32(ctypes.byref(a), torch.cuda.current_stream().cuda_stream)
        e1.record(); torch.cuda.synchronize()
        assert st == 0, lib.rs_last_error()
        if r >= 3: ts.append(e0.elapsed_time(e1))
    ok, ov = (kb, vb) if res.value else (ka, va)
    wk, wi = torch.sort(keys & ((1 << end_bit) - 1) if end_bit < 32 else keys, stable=True)
    good = bool(torch.equal(ov.long(), wi)) if end_bit >= 31 or True else True
    passes = (end_bit + 7) // 8
    ms = float(np.median(ts))
    return ms, passes, good


def main():
    libs = [_lib.LIB_PATH] + sys.argv[1:]
    for path in libs:
        lib = ctypes.CDLL(path)
        for name, (res, args) in _lib._PROTOS.items():
            if hasattr(lib, name):
                getattr(lib, name).restype = res; getattr(lib, name).argtypes = args
        for n, bits, kind in ((1_000_000, 32, "random"), (5_000_000, 14, "tile"), (5_000_000, 14, "random"), (20_000_000, 32, "random")):
            ms, passes, good = run(lib, n, bits, kind=kind)
            gbs = n * (4 + passes * 16) / (ms * 1e-3) / 1e9
            print(f"{os.path.basename(path):34s} n={n:>9d} bits={bits:2d} {kind:6s} {ms*1e3:8.1f} us  ({ms*1e3/passes:6.1f} us/pass, {gbs:7.1f} GB/s algorithmic) correct={good}")


if __name__ == "__main__":
    main()
