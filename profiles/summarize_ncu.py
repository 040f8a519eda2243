"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary committed under
profiles/ (one row per captured launch, the metrics DESIGN.md quotes).

    python profiles/summarize_ncu.py gpurun_out/prof_raw.csv > profiles/ncu_full_rNN_summary.csv"""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
cols = [c for c in COLS if c in hdr]
w.writerow(cols)
w.writerow([units[hdr.index(c)] for c in cols])
for r in rows[2:]:
    if len(r) == len(hdr):
        w.writerow([r[hdr.index(c)] for c in cols])
