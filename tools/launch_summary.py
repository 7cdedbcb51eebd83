"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (and grid)."""
import collections
import csv
import re
import sys


def short(n):
    m = re.search(r'(\w+_kernel\w*|\w+Kernel\w*)', n)
    base = m.group(1) if m else n[:40]
    t = re.search(r'_kernel<([^>]*)>', n)
    return base + ('<' + t.group(1).replace('(int)', '') + '>' if t else '')


def main(path, steps=3.0, detail=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = short(r['Kernel Name'])
        agg[k][0] += 1
        agg[k][1] += float(r['Metric Value']) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)}  total {tot / 1e3:.3f} ms  (~{tot / 1e3 / steps:.3f} ms/step over {steps} steps)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"  {k[:58]:58s} n/step={v[0] / steps:6.1f} us/step={v[1] / steps:8.1f} share={v[1] / tot:.3f} avg_us={v[1] / v[0]:.1f}")
    if detail:
        g = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            if detail in r['Kernel Name']:
                key = (short(r['Kernel Name']), r['Grid Size'])
                g[key][0] += 1
                g[key][1] += float(r['Metric Value']) / 1e3
        for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:30]:
            print(f"    {k[0][:34]:34s} grid={k[1]:16s} n/step={v[0] / steps:5.1f} us/step={v[1] / steps:8.1f} avg_us={v[1] / v[0]:7.1f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 3.0, sys.argv[3] if len(sys.argv) > 3 else None)
