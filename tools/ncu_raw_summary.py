#!/usr/bin/env python
"""Compact summary of `ncu --page raw --csv` exports (one row per captured kernel launch).
usage: ncu_raw_summary.py out.csv file1.raw.csv [file2.raw.csv ...]"""
import csv
import os
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("sm__cycles_elapsed.avg", "sm_cycles"), ("smsp__inst_executed.sum", "warp_insts"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts")]


def main():
    out, files = sys.argv[1], sys.argv[2:]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["capture", "kernel"] + [k[1] for k in KEYS])
        for f in files:
            rows = list(csv.reader(open(f)))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            for vals in rows[2:]:
                name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
                rec = []
                for key, _ in KEYS:
                    if key in hdr:
                        i = hdr.index(key)
                        rec.append(f"{vals[i]} {units[i]}".strip())
                    else:
                        rec.append("")
                w.writerow([os.path.basename(f).replace(".raw.csv", ""), name.split("(")[0][-40:]] + rec)


if __name__ == "__main__":
    main()
