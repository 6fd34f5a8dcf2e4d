#!/usr/bin/env python
"""Extracts the metrics DESIGN.md / profiles/ncu_summary_*.md quote from .ncu-rep files (read here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/ncu_summary_rNN.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    return hdr, units, rd[2:]


def main():
    for rep in sys.argv[1:]:
        hdr, units, rows = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            name = r[idx["Kernel Name"]]
            print(f"## {name}\n\n(`{rep.split('/')[-1]}`, launch id {r[idx['ID']]}, grid {r[idx['Grid Size']]} block {r[idx['Block Size']]})\n")
            print("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in idx:
                    print(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
            if "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum" in idx:
                try:
                    s = float(r[idx["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]].replace(",", ""))
                    q = float(r[idx["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]].replace(",", ""))
                    if q > 0:
                        print(f"| global-load sectors per request | {s / q:.2f} | |")
                except ValueError:
                    pass
            print()


if __name__ == "__main__":
    main()
