"""Summarise .ncu-rep files (raw page): one line per captured launch with duration, DRAM bytes, L2 / tensor /
shared-memory utilisation.

    python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]                 # text table
    python tools/ncu_summary.py --json out.json a.ncu-rep [b.ncu-rep ...]  # + JSON (bench.py reads `roofline_launch`,
                                                                          #   the longest igemm_halo launch captured)
"""
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("sm__cycles_elapsed.avg", "cycles"),
    ("sm__cycles_elapsed.avg.per_second", "ghz"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6,
         "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    idx = {k: h.index(k) for k, _ in KEYS if k in h}
    kn = h.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"kernel": r[kn], "file": path}
        for k, short in KEYS:
            if k in idx:
                v = float(r[idx[k]].replace(",", "")) if r[idx[k]] not in ("", "n/a") else None
                if v is not None and short in ("dur", "dram_rd", "dram_wr"):
                    v *= SCALE.get(units[idx[k]], 1.0)          # durations in us, bytes in bytes
                d[short] = v
        res.append(d)
    return res


def main():
    args = sys.argv[1:]
    out_json = None
    if args and args[0] == "--json":
        out_json, args = args[1], args[2:]
    allrows = []
    for p in args:
        allrows += rows_of(p)
    for d in allrows:
        print("%-44s dur=%8.1fus dram=%8.1fMB(r)+%7.1fMB(w) dram%%=%5.1f l2%%=%5.1f l2hit=%5.1f tensor%%=%5.1f "
              "smem_tc%%=%5.1f sm%%=%5.1f ghz=%.2f grid=%d regs=%d" % (
                  d["kernel"][:44], d["dur"], d["dram_rd"] / 1e6, d["dram_wr"] / 1e6, d["dram%"], d["l2%"],
                  d.get("l2hit%") or 0, d["tensor%"], d.get("smem_tc%") or 0, d["sm%"], d["ghz"], d["grid"], d["regs"]))
    if out_json:
        # the dominant kernel: the row-resident conv (round 1, second half) if captured, else the halo kernel
        halo = [d for d in allrows if "igemm_rows" in d["kernel"]] or [d for d in allrows if "igemm_halo" in d["kernel"]]
        top = max(halo, key=lambda d: d["dur"]) if halo else None
        doc = {"source": "ncu --set full --clock-control none (cold-cache, serialised replays)", "launches": allrows}
        if top:
            doc["roofline_launch"] = {"kernel": top["kernel"], "duration_us": top["dur"],
                                      "dram_bytes": top["dram_rd"] + top["dram_wr"], "dram_read_bytes": top["dram_rd"],
                                      "dram_write_bytes": top["dram_wr"], "tensor_pipe_pct": top["tensor%"],
                                      "smem_tensor_read_pct": top.get("smem_tc%"), "grid": top["grid"]}
        json.dump(doc, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
