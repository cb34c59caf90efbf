#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into the handful of numbers the roofline
argument uses.  Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [...] > profiles/x.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of 64/SM"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"### {path.split('/')[-1]}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print(f"**{d.get('Kernel Name', '?')}**\n")
            print("| metric | value |\n|---|---|")
            for key, label in WANT:
                if key in d and d[key] != "":
                    print(f"| {label} (`{key}`) | {d[key]} {u.get(key, '')} |")
            try:
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tb = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
                ms = float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]]
                print(f"| **traffic = read + write** | {tb / 1e9:.3f} GB -> {tb / ms / 1e6:.0f} GB/s over the ncu duration |")
            except Exception:
                pass
            print()


if __name__ == "__main__":
    main()
