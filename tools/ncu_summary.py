"""Summarise an .ncu-rep (raw page) into one line per captured launch: duration, DRAM bytes, L2 / tensor / smem
utilisation. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/...txt]"""
import csv
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
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_elapsed", "tcinst%"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__cycles_elapsed.avg", "cycles"),
    ("sm__cycles_elapsed.avg.per_second", "ghz"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
idx = {k: h.index(k) for k, _ in KEYS if k in h}
kn = h.index("Kernel Name")
for r in rows[2:]:
    parts = [r[kn][:48].ljust(48)]
    for k, short in KEYS:
        if k in idx:
            parts.append("%s=%s%s" % (short, r[idx[k]], units[idx[k]] if short in ("dur", "dram_rd", "dram_wr") else ""))
    print("  ".join(parts))
