"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,sm__pipe_tensor_cycles_active...]` log per kernel:
    python tools/ncu_launch_shares.py launches.csv [--tensor] > profiles/rNN_ncu_launch_shares.txt
--tensor: time-weighted sm__pipe_tensor_cycles_active per kernel and over all implicit-GEMM launches."""
import collections
import csv
import re
import sys

if __name__ == "__main__":
    path = sys.argv[1]
    want_tensor = "--tensor" in sys.argv
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    h = rows[0]
    kn, mn, mv, mu, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        d = per.setdefault(r[idc], {"name": re.sub(r"\(.*", "", r[kn])})
        val = float(r[mv].replace(",", "")) if r[mv] not in ("", "n/a") else 0.0
        if r[mn].startswith("gpu__time_duration"):
            d["us"] = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[mu], 1e-3)
        elif "pipe_tensor" in r[mn]:
            d["tensor"] = val
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("us", 0.0)
        a[2] += d.get("us", 0.0) * d.get("tensor", 0.0)
    total = sum(a[1] for a in agg.values())
    print(f"# {len(per)} launches, {total / 1e3:.1f} ms total")
    for name, (n, us, tw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = f"  tensor pipe {tw / us:5.1f} %" if (want_tensor and us > 0) else ""
        print(f"{name[:76]:76s} {n:6d} {us / 1e3:9.2f} ms {100 * us / total:5.1f}%{extra}")
    if want_tensor:
        g = [(us, tw) for name, (n, us, tw) in agg.items() if re.search("igemm|wgrad", name)]
        tu, tt = sum(x[0] for x in g), sum(x[1] for x in g)
        c = [(us, tw) for name, (n, us, tw) in agg.items() if re.search("igemm", name)]
        cu, ct = sum(x[0] for x in c), sum(x[1] for x in c)
        if tu > 0:
            print(f"# all conv GEMMs (forward, input gradient, weight gradient): {tt / tu:.1f} %  over {tu / 1e3:.1f} ms")
        if cu > 0:
            print(f"# forward + input-gradient GEMMs only: {ct / cu:.1f} %  over {cu / 1e3:.1f} ms")
