"""Condense `ncu -i report.ncu-rep --page raw --csv` into the handful of columns the profiles/ summaries keep
(one row per captured launch).  The .ncu-rep files themselves stay in gpurun_out/ (scratch, tens of MB).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/ncu_condense.py /tmp/x.csv > profiles/<name>.csv
"""
import csv
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor"]


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    hdr = rows[0]
    keep = [c for c in COLS if c in hdr]
    idx = [hdr.index(c) for c in keep]
    w = csv.writer(sys.stdout)
    for r in rows:
        w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
