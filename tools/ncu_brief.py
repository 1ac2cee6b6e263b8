#!/usr/bin/env python
"""Brief of an `ncu --page raw --csv` export: duration, DRAM bytes, pipe / issue utilisation and the warp-stall breakdown per kernel.
usage: python tools/ncu_brief.py raw.csv [more.csv ...]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
STALL = "smsp__average_warps_issue_stalled_"


def brief(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s  [%s]" % (r[idx["Kernel Name"]][:70], path))
        for k, name in KEYS:
            if k in idx:
                print("   %-20s %s %s" % (name, r[idx[k]], units[idx[k]]))
        st = sorted(((float(r[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]) for h, i in idx.items()
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]), reverse=True)
        print("   stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))


for p in sys.argv[1:]:
    brief(p)
