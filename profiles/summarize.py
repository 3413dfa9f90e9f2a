"""Condense ncu exports into the text summaries committed under profiles/.
    python profiles/summarize.py launches gpurun_out/launches.csv FRAMES > profiles/rNN_launches.txt
    python profiles/summarize.py raw gpurun_out/prof_raw.csv > profiles/rNN_kernels.txt
(raw csv = `ncu -i X.ncu-rep --page raw --csv`)"""
import collections
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "inst"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
]


def launches(path, frames):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0][:64], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {frames} frames; times are cold-cache/serialised: compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:66s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.2f}us per_frame={sum(v)/frames/1e3:9.2f}us share={sum(v)/tot:6.3f}")
    print(f"sum per frame: {tot/frames/1e3:.1f} us")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    seen = collections.OrderedDict()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0][:48]
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[len(rs) // 2]
        out = [f"{name:48s} launches={len(rs)}"]
        for key, short in KEYS:
            if key in hdr:
                i = hdr.index(key)
                try:
                    v = float(r[i].replace(",", ""))
                    out.append(f"{short}={v:.4g}{units[i] if short in ('time','dram_rd','dram_wr') else ''}")
                except ValueError:
                    out.append(f"{short}={r[i]}")
        print(" ".join(out))


def traffic(paths, source):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel (median launch) -> JSON for bench.py's
    `roofline.traffic` (profiles/r02_kernel_traffic.json)."""
    import json
    import re

    out = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                 "gpu__time_duration.sum"))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per = collections.OrderedDict()
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[ki]).split("<")[0].split("(")[0]
            b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
            per.setdefault(name, []).append((b, float(r[ti].replace(",", "")), units[ti]))
        for name, v in per.items():
            v.sort()
            b, t, tu = v[len(v) // 2]
            out[name] = {"dram_bytes_per_launch": int(b), "launches_captured": len(v), "time": f"{t:g} {tu}"}
    print(json.dumps({"source": source, "kernels": out}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[3:], sys.argv[2])
    else:
        raw(sys.argv[2])
