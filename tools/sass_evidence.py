#!/usr/bin/env python
"""Occurrences per kernel of the sm_100a instructions the design relies on, from `cuobjdump -sass` of the built library:
UBLKCP (cp.async.bulk = TMA bulk copy), SYNCS (mbarrier), LDGSTS (cp.async global->shared), MATCH (match.any ranking),
REDUX (warp reduce, tile-footprint owner lookup), REDG / ATOMG (global reductions), ATOMS (shared-memory atomics), MUFU.EX2,
system-scope accesses (flags of the peer exchange).

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3dgs_rigidbody_b200", "lib", "librigidsplat.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS(mbarrier)", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("MATCH", r"\bMATCH"),
            ("REDUX", r"\bREDUX"), ("REDG (global reduction)", r"\bREDG?\."), ("ATOMG", r"\bATOMG"), ("ATOMS", r"\bATOMS"), ("MUFU.EX2", r"MUFU\.EX2"),
            ("SYS-scope", r"\.SYS\b"), ("UTMA/UTC (tensor)", r"\bUTMA|\bUTC")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, it = {}, [], None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", next(it))
            counts[cur] = collections.OrderedDict()
            order.append(cur)
            continue
        if cur is None:
            continue
        for label, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][label] = counts[cur].get(label, 0) + 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# cuobjdump -sass 3dgs_rigidbody_b200/lib/librigidsplat.so   arch: {', '.join(arch)}   (tools/sass_evidence.py)")
    print("# occurrences per kernel: " + ", ".join(l for l, _ in PATTERNS))
    for k in sorted(order):
        if counts[k]:
            print(f"{k[:78]:80s} {dict(counts[k])}")
    tot = collections.Counter()
    for k in order:
        tot.update(counts[k])
    print("# total:", dict(tot))


if __name__ == "__main__":
    main()
