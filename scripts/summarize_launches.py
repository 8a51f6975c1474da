#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share.

    python scripts/summarize_launches.py gpurun_out/launches.csv [steps_captured] > profiles/rNN_launches.md
    python scripts/summarize_launches.py gpurun_out/launches.csv --steps-between adam_kernel 2
        keeps exactly the launches after the 1st occurrence of that kernel up to and including the 3rd (= 2 whole steps)
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    marker = None
    if len(sys.argv) > 3 and sys.argv[2] == "--steps-between":
        marker, steps = sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    else:
        steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if marker:
        hits = [i for i, r in enumerate(rows) if marker in r["Kernel Name"]]
        if len(hits) < int(steps) + 1:
            sys.exit(f"only {len(hits)} launches of {marker} in {path}")
        rows = rows[hits[0] + 1:hits[int(steps)] + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", re.sub(r"[<(].*", "", r["Kernel Name"]))
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r["Metric Unit"], 1e-6)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"source: `{path}` ({sum(v[0] for v in agg.values())} launches, {steps:g} step(s) captured; "
          "ncu times are cold-cache and serialised: compare SHARES)\n")
    print("| kernel | launches/step | ms/step | share |")
    print("|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0005:
            continue
        print(f"| `{k}` | {v[0] / steps:.1f} | {v[1] / steps:.2f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | | {tot / steps:.2f} | |")


if __name__ == "__main__":
    main()
