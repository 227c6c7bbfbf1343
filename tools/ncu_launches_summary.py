"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count / time / share of the
LAST complete step (a step starts at the `albert_embed_kernel` launch).

    python tools/ncu_launches_summary.py gpurun_out/launches.csv "<command line>" > profiles/x_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        rows.append((r[ki], ns))
    starts = [i for i, (k, _) in enumerate(rows) if "albert_embed" in k]
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]
    else:
        a, b = 0, len(rows)
    step = rows[a:b]
    agg = OrderedDict()
    for k, ns in step:
        k = re.sub(r"^void\s+", "", k)
        k = re.sub(r"\(.*$", "", k).replace("kkx::", "").replace("(anonymous namespace)::", "")
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(cmd)
    print(f"{len(rows)} launches captured; last complete step = launches {a}..{b - 1}: {len(step)} launches, "
          f"total {tot / 1e6:.2f} ms (cold caches, serialised under ncu: compare shares, not absolutes)")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} n={n:4d} {ns / 1e6:9.3f} ms {100 * ns / tot:5.1f}%")


if __name__ == "__main__":
    main()
