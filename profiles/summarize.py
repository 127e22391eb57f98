#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep "title" > profiles/<name>.md
    python profiles/summarize.py gpurun_out/prof.ncu-rep "title" --traffic tilt@65536 > profiles/<name>.md

`--traffic <variant>@<envs>` also records the mean dram__bytes_read.sum / dram__bytes_write.sum per launch of the
report in profiles/ncu_traffic.json, which is where bench.py reads `roofline.traffic` from.
"""
import json
import os
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


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def record_traffic(rep, key, hdr, units, rows):
    out = {}
    for metric, name in (("dram__bytes_read.sum", "dram_read_bytes"), ("dram__bytes_write.sum", "dram_write_bytes"),
                         ("gpu__time_duration.sum", "duration_ns")):
        i = hdr.index(metric)
        scale = UNIT.get(units[i], {"us": 1e3, "ns": 1.0, "ms": 1e6}.get(units[i], 1.0))
        vals = [float(r[i].replace(",", "")) * scale for r in rows]
        out[name] = sum(vals) / len(vals)
    out["launches"] = len(rows)
    out["source"] = rep
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = out
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    if "--traffic" in sys.argv:
        record_traffic(rep, sys.argv[sys.argv.index("--traffic") + 1], hdr, units, rows[2:])
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
