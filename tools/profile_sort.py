#!/usr/bin/env python
"""One radix sort of N (key, value) pairs by the hand-written sort and one by cub, for `ncu --set full` side by side:
    ncu --set full --clock-control none -k regex:"sort_pass|sort_hist|Onesweep|Histogram" -c 12 -o gpurun_out/prof_sort \
        python tools/profile_sort.py --n 20000000 --bits 32"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sort_bench as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20_000_000)
    ap.add_argument("--bits", type=int, default=32)
    args = ap.parse_args()
    keys = sb.make_keys(args.n, args.bits, "random", False)
    sb.time_it.__defaults__ = (1,)  # one timed repetition after the three warm-ups
    lib = sb._lib.load()
    print("ours", sb.run_ours(lib, keys, args.bits, False)[0])
    cub = ctypes.CDLL(sb.CUB)
    print("cub", sb.run_cub(cub, keys, args.bits, False)[0])


if __name__ == "__main__":
    main()
