"""Summarise an `ncu --page raw --csv` dump: one block per kernel launch with the metrics the
roofline needs.  Usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum"]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = re.sub(r"^void ", "", name)
    print("==", name[:110])
    rd = wr = None
    for w in want:
        if w in idx:
            print(f"   {w:62s} {r[idx[w]]:>16s} {units[idx[w]]}")
            if w == "dram__bytes_read.sum":
                rd = (float(r[idx[w]]), units[idx[w]])
            if w == "dram__bytes_write.sum":
                wr = (float(r[idx[w]]), units[idx[w]])
    if rd and wr:
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd[0] * mul[rd[1]] + wr[0] * mul[wr[1]]
        dur = float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[units[idx["gpu__time_duration.sum"]]]
        print(f"   {'dram traffic (read+write) per launch':62s} {tot:16.0f} byte   -> {tot / dur / 1e9:8.1f} GB/s under ncu")
