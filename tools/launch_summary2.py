"""Summarise an ncu multi-metric launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum, --csv) of
bench.py into per-kernel and per-GEMM-grid tables over the last three sampling steps.
   python tools/launch_summary2.py gpurun_out/launches.csv [out.txt] [traffic.json]"""
import collections
import csv
import json
import re
import sys


def short(n):
    m = re.search(r'(\w+_kernel\w*)', n)
    base = m.group(1) if m else n[:40]
    t = re.search(r'_kernel<([^>]*)>', n)
    return base + ('<' + t.group(1).replace('(int)', '').replace('(bool)', '') + '>' if t else '')


def main(path, out_path=None, traffic_path=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    byid = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = byid.setdefault(r['ID'], {'name': r['Kernel Name'], 'grid': r['Grid Size']})
        v = float(r['Metric Value'].replace(',', ''))
        v *= {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1e3, 'ms': 1e6}.get(r['Metric Unit'], 1)
        d[r['Metric Name']] = v
    L = list(byid.values())
    nt = [i for i, d in enumerate(L) if 'next_timestep' in d['name']]
    ns = min(3, len(nt) - 1)                               # complete steady-state steps in the capture (three when available)
    L = L[nt[-1 - ns]:nt[-1]]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in L:
        a = agg[short(d['name'])]
        a[0] += 1; a[1] += d['gpu__time_duration.sum'] / 1e3; a[2] += d['dram__bytes_read.sum']; a[3] += d['dram__bytes_write.sum']
    tot = sum(a[1] for a in agg.values())
    out = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip S --launch-count C "
           "python tools/step_once.py --steps 6   (eager launches of the step program, no CUDA graph)",
           f"# {ns} steady-state sampling step(s) (config 2: SD1.5 64x64 latent, UNet batch 2, bf16); per-launch times are cold-cache and "
           "serialised (ncu), so SHARES are the comparable quantity",
           f"launches {len(L)}  total {tot / 1e3:.3f} ms  (~{tot / (ns * 1e3):.3f} ms/step under ncu)"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"  {k[:52]:52s} n/step={a[0] / ns:6.1f} us/step={a[1] / ns:8.1f} share={a[1] / tot:.3f} avg_us={a[1] / a[0]:6.1f}  "
                   f"dram MB/step rd={a[2] / (ns * 1e6):8.1f} wr={a[3] / (ns * 1e6):7.1f}")
    g = collections.defaultdict(lambda: [0, 0.0, 0.0])
    gem = [d for d in L if 'conv_gemm_tc' in d['name'] or 'splitk' in d['name']]
    for d in gem:
        a = g[(short(d['name']), d['grid'])]
        a[0] += 1; a[1] += d['gpu__time_duration.sum'] / 1e3; a[2] += d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']
    out.append("\n# tcgen05 GEMM launches by grid")
    for k, a in sorted(g.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append(f"    {k[0][:40]:40s} grid={k[1]:16s} n/step={a[0] / ns:5.1f} us/step={a[1] / ns:8.1f} avg_us={a[1] / a[0]:7.1f} "
                   f"dram MB/launch={a[2] / a[0] / 1e6:7.2f}")
    text = "\n".join(out) + "\n"
    print(text)
    if out_path:
        open(out_path, 'w').write(text)
    if traffic_path:
        rd = sum(d['dram__bytes_read.sum'] for d in gem) / ns
        wr = sum(d['dram__bytes_write.sum'] for d in gem) / ns
        t = sum(d['gpu__time_duration.sum'] for d in gem) / ns
        json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum summed over all tcgen05 GEMM + split-K reduce launches of ONE step "
                           f"({len(gem) // ns} launches), ncu --metrics, default cache control (L2 flushed before every kernel: cold-cache upper "
                           "bound; writes stay in L2 until evicted), SD1.5 64x64 UNet batch 2",
                   "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "sum_kernel_time_ns_under_ncu": int(t),
                   "algorithmic": "weights 1.72e9 B (bf16) streamed once + activations; SURVEY 8(d) whole-step figure 4.42e9 B",
                   "source": out_path or path}, open(traffic_path, 'w'), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
