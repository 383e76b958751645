"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel count /
total / share (the SHARE of the step is what must agree with bench.py's live CUDA-event timing;
ncu's per-launch times are cold-cache and serialised).

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    for r in rd:
        if r[mi] != "gpu__time_duration.sum":
            continue
        rows.append((r[ki], float(r[vi].replace(",", "")), r[gi], r[bi]))
    agg = collections.OrderedDict()
    for name, ns, grid, block in rows:
        short = re.sub(r"\(.*$", "", name)
        short = short if len(short) < 110 else short[:107] + "..."
        a = agg.setdefault(short, [0, 0.0, 1e30, grid, block])
        a[0] += 1
        a[1] += ns
        a[2] = min(a[2], ns)
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms of GPU time under ncu (serialised, cold cache)")
    print(f"{'share':>7} {'count':>6} {'total_us':>12} {'avg_us':>10} {'min_us':>10}  grid block  kernel")
    for name, (n, ns, mn, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * ns / total:6.2f}% {n:6d} {ns / 1e3:12.1f} {ns / 1e3 / n:10.2f} {mn / 1e3:10.2f}  {grid} {block}  {name}")


if __name__ == "__main__":
    main(sys.argv[1])
