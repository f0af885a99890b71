#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table:
per-kernel-class totals over the whole run and the launch list of ONE training step (the launches
between the last two L2-flush fills).  Usage: summarize_launches.py launches.csv title > out.md"""
import collections
import csv
import sys


def main():
    path, title = sys.argv[1], sys.argv[2]
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    ks = [(r["Kernel Name"], float(r["Metric Value"]) / 1000.0, r["Grid Size"], r["Block Size"]) for r in rows]
    flush = [i for i, k in enumerate(ks) if "FillFunctor<unsigned char>" in k[0]]
    a, b = flush[-2] + 1, flush[-1]
    step = ks[a:b]
    print(f"# {title}\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    tot = sum(k[1] for k in step)
    print(f"One captured step = {len(step)} kernel launches, {tot:.1f} us summed.\n")
    agg = collections.OrderedDict()
    for n, t, g, blk in step:
        key = n.split("(")[0][:90]
        c = agg.setdefault(key, [0, 0.0])
        c[0] += 1
        c[1] += t
    print("| kernel | launches | us (sum) | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    print("\nFull per-launch list of that step:\n\n| # | us | grid | block | kernel |\n|---:|---:|---|---|---|")
    for i, (n, t, g, blk) in enumerate(step):
        print(f"| {i} | {t:.2f} | {g} | {blk} | `{n[:110]}` |")


if __name__ == "__main__":
    main()
