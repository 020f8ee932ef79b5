#!/usr/bin/env python
"""Pull the metrics profiles/ keeps out of `ncu -i X.ncu-rep --page raw --csv`.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_extract.py raw.csv [row]      # JSON on stdout
"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum",
]


def main(path, row=0):
    rows = list(csv.reader(open(path)))
    head, units, vals = rows[0], rows[1], rows[2 + row]
    out = {"kernel": vals[head.index("Kernel Name")]}
    for name in KEEP:
        if name in head:
            i = head.index(name)
            out[name] = f"{vals[i]} {units[i]}".strip()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
