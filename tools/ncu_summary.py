"""Summarise an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv`."""
import csv
import sys

COLS = ["ID", "Kernel Name", "Grid Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
