"""Condense `.ncu-rep` captures (ncu --set full --clock-control none --import-source on) into the text kept
under profiles/:  python tools/ncu_summary.py "<how it was captured>" a.ncu-rep b.ncu-rep ... > profiles/x.txt
Reads the reports with `ncu -i ... --page raw --csv`; one section per profiled launch."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max",
]
STALL = "smsp__average_warps_issue_stalled_"


def sections(path, how):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(head)}
    for r in rows[2:]:
        print(f"== {r[col['Kernel Name']]}  ({how})")
        for k in KEEP:
            if k in col and r[col[k]] != "":
                print(f"{k} = {r[col[k]]} {units[col[k]]}")
        for name, i in sorted(col.items()):
            if name.startswith(STALL) and name.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.15:
                    print(f"  stall {name[len(STALL):-len('_per_issue_active.ratio')]} = {v:.2f} warps per issue")
        print()


if __name__ == "__main__":
    how = sys.argv[1]
    for path in sys.argv[2:]:
        sections(path, how)
