"""Per-kernel extract of an `ncu -i X.ncu-rep --page raw --csv` export: the handful of metrics profiles/ keeps.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv ; python tools/ncu_extract.py x_raw.csv
"""
import csv, re, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen = set()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
    key = (name, r[idx["Grid Size"]], r[idx["Block Size"]])
    if key in seen and "--all" not in sys.argv:
        continue
    seen.add(key)
    print(f"--- {name[:70]}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}")
    for w in WANT:
        if w in idx:
            print(f"   {w:78s} {r[idx[w]]} {units[idx[w]]}")
