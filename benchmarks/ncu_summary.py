"""Print the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` output (one column per captured launch)."""
import csv
import re
import subprocess
import sys

KEYS = [r"gpu__time_duration\.sum", r"dram__bytes_read\.sum$", r"dram__bytes_write\.sum$", r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed",
        r"gpu__dram_throughput", r"sm__throughput\.avg\.pct", r"sm__warps_active\.avg\.pct_of_peak_sustained_active",
        r"launch__registers_per_thread", r"launch__occupancy_limit", r"launch__waves_per_multiprocessor", r"launch__grid_size",
        r"launch__shared_mem_per_block", r"sm__inst_executed\.sum$", r"smsp__issue_active\.avg\.pct", r"smsp__inst_executed\.avg\.per_cycle_active",
        r"sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|xu|uniform)\.avg\.pct", r"smsp__average_warps_issue_stalled_.*_per_issue_active",
        r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"lts__t_sector_hit_rate\.pct", r"smsp__cycles_active\.avg$", r"sm__cycles_elapsed\.max$",
        r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$", r"smsp__inst_executed_op_shared", r"sm__pipe_tensor"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernels:", [r[hdr.index("Kernel Name")][:60] for r in data])
    for i, h in enumerate(hdr):
        name = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
        if any(re.search(k, h) for k in KEYS):
            print(f"{h[-95:]:95s} {units[i]:14s} {[r[i] for r in data]}")


if __name__ == "__main__":
    main(sys.argv[1])
