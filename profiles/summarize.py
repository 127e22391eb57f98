#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep "title" > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written (inside the launch; outputs mostly stay in L2)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 sectors written by SMs"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks/SM"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch values, cold-cache, serialised)\n")
    names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
    print("kernels: " + "; ".join(sorted(set(n.split("(")[0] for n in names))) + f"  ({len(names)} launches)\n")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    print("|---|---|" + "---|" * (len(rows) - 2))
    for key, label in KEYS:
        if key in hdr:
            i = hdr.index(key)
            vals = []
            for r in rows[2:]:
                try:
                    vals.append(f"{float(r[i].replace(',', '')):,.2f}")
                except ValueError:
                    vals.append(r[i])
            print(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
