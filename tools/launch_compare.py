"""Per-launch comparison of two ncu launch lists (gpu__time_duration.sum, --csv) of tools/step_once.py: aligns the LAST step of
each list and prints per-kernel-class totals plus every launch whose kernel/grid differs between the two.
    python tools/launch_compare.py a.csv b.csv"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"us": 1e3, "ms": 1e6, "ns": 1, "s": 1e9}.get(r["Metric Unit"], 1)
        n = r["Kernel Name"]
        m = re.search(r"(\w+_kernel)\s*<([^>]*)>", n) or re.search(r"(\w+_kernel)", n)
        if m is None:
            name = n[:40]
        else:
            name = m.group(1) + ("<" + m.group(2).replace("(int)", "").replace("(bool)", "") + ">" if m.lastindex and m.lastindex > 1 else "")
        rows.append((name, r["Grid Size"].replace(" ", ""), v / 1e3))
    nt = [i for i, r in enumerate(rows) if "next_timestep" in r[0]]
    return rows[nt[-1]:] if nt else rows


def main(a, b):
    A, B = load(a), load(b)
    for tag, L in (("A " + a, A), ("B " + b, B)):
        agg = collections.defaultdict(lambda: [0, 0.0])
        for n, g, t in L:
            agg[n][0] += 1
            agg[n][1] += t
        tot = sum(v[1] for v in agg.values())
        print(f"== {tag}: {len(L)} launches, {tot:.1f} us")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"   {k:52s} n={v[0]:4d} us={v[1]:9.1f} avg={v[1] / v[0]:7.2f}")
    print("== launches in order (A | B), '*' where kernel or grid differ")
    ia = ib = 0
    while ia < len(A) or ib < len(B):
        ra = A[ia] if ia < len(A) else ("-", "", 0.0)
        rb = B[ib] if ib < len(B) else ("-", "", 0.0)
        same = ra[0] == rb[0] and ra[1] == rb[1]
        if same:
            if abs(ra[2] - rb[2]) > 1.5:
                print(f"    {ra[0][:44]:44s} {ra[1]:14s} {ra[2]:7.1f} | {rb[2]:7.1f}  d={rb[2] - ra[2]:+.1f}")
            ia += 1
            ib += 1
        else:
            print(f"  * {ra[0][:44]:44s} {ra[1]:14s} {ra[2]:7.1f} | {rb[0][:44]:44s} {rb[1]:14s} {rb[2]:7.1f}")
            ia += 1
            ib += 1


if __name__ == "__main__":
    main(*sys.argv[1:3])
