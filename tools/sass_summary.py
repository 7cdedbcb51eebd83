"""Per-kernel SASS mnemonic counts of the shipped library: which kernels are Blackwell-native.

    python tools/sass_summary.py [path/to/libsdb200.so] > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (5th-gen tensor core, TMEM accumulator), LDTM / STTM = tcgen05.ld / .st (TMEM <-> registers),
UTMALDG / UTMASTG = TMA tensor load / store, UTMAREDG = TMA reduce, HMMA = legacy mma.sync, UBLKCP = bulk copy.
Runs on the build machine (cuobjdump only, no GPU).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "FFMA", "MUFU.EX2", "SYNCS"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stable-diffusion-pytorch_b200", "libsdb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
    print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   (sm_100a; counts of SASS instructions per kernel)")
    print(f"# {'kernel':<78} {'instr':>6} " + " ".join(f"{m:>8}" for m in MNEMONICS))
    for fn in order:
        c = counts[fn]
        name = demangle(fn).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"\([^()]*\)$", "", name).replace("(int)", "").replace("(bool)", "").replace("void ", "")
        print(f"  {name[:78]:<78} {c['_total']:>6} " + " ".join(f"{c[m]:>8}" for m in MNEMONICS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"# {'TOTAL':<78} {tot['_total']:>6} " + " ".join(f"{tot[m]:>8}" for m in MNEMONICS))


if __name__ == "__main__":
    main()
