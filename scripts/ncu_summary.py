"""Summarise an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline discussion uses.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/prof_summary.csv
"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "smsp__inst_executed_op_global_atom.sum",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in WANT]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
