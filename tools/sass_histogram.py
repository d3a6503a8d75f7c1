"""Per-kernel SASS opcode histogram of the in-tree library: proves which kernels carry tcgen05 / TMEM /
TMA instructions. Runs anywhere cuobjdump exists (no GPU needed):
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (bf16), LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit,
UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), UTMAPF = TMA L2 prefetch, UBLKCP = cp.async.bulk (1-D bulk
copy), SYNCS = mbarrier ops, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents, REDG / ATOMG = global
reductions / atomics."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tactile_gan_b200", "libtactile_gan_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "ACQBULK",
         "PREEXIT", "HMMA", "REDG", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "FFMA", "BAR")

if __name__ == "__main__":
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels; columns = instruction counts in the sm_100a SASS")
    print("# " + " ".join(f"{w:>8s}" for w in ("total",) + WATCH) + "  kernel")
    tot = collections.Counter()
    for (name, cnt), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r"\(.*", "", pretty)
        row = [sum(cnt.values())] + [cnt.get(w, 0) for w in WATCH]
        for w in WATCH:
            tot[w] += cnt.get(w, 0)
        print("  " + " ".join(f"{v:8d}" for v in row) + "  " + pretty)
    print("# totals: " + ", ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]))
