#!/usr/bin/env python3
"""Summarise an `ncu --csv --page raw` capture: one markdown table per kernel launch with the metrics the profiles/ notes quote.
usage: ncu_summary.py capture.csv [more.csv ...]  (prints markdown)"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[h], rows[h + 1], rows[h + 2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, data = load(path)
        col = {n: i for i, n in enumerate(hdr)}
        for d in data:
            print(f"### {d[col['Kernel Name']]}  (`{path.split('/')[-1]}`, launch id {d[0]})\n")
            print("| metric | unit | value |\n|---|---|---|")
            for k in KEYS:
                if k in col and d[col[k]] != "":
                    print(f"| {k} | {units[col[k]]} | {d[col[k]]} |")
            st = []
            for n, i in col.items():
                if n.startswith(STALL) and not n.endswith("_not_issued"):
                    try:
                        st.append((float(d[i].replace(",", "")), n[len(STALL):]))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            tot = sum(v for v, _ in st) or 1.0
            print("\nStall reasons (share of sampled warp-cycles): " + ", ".join(f"{n} {100 * v / tot:.1f} %" for v, n in st[:8]) + "\n")


if __name__ == "__main__":
    main()
