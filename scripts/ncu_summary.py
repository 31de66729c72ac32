#!/usr/bin/env python3
"""Condense an ncu report (--set full) into the handful of counters DESIGN.md / profiles/ quote.

    python scripts/ncu_summary.py gpurun_out/t6/prof.ncu-rep [--json out.json]
Prints one markdown table row block per profiled launch (duration, DRAM bytes, issue / fp64 utilisation, warps, stalls).
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    ("Grid Size", "grid"), ("gpu__time_duration.sum", "duration_us"), ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_static", "smem_static"), ("launch__waves_per_multiprocessor", "waves"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__inst_executed.sum", "warp_instructions"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for name, key in WANT:
            if name in hdr:
                i = hdr.index(name)
                v = r[i]
                if key.startswith("dram_") and key != "dram_throughput_pct":
                    rec[key + "_bytes"] = to_bytes(v, units[i])
                elif key == "duration_us":
                    f = float(v.replace(",", ""))
                    rec[key] = f / 1e3 if units[i] in ("ns", "nsecond") else f
                elif key == "grid":
                    rec[key] = v
                else:
                    try:
                        rec[key] = float(v.replace(",", ""))
                    except ValueError:
                        rec[key] = v
        out.append(rec)
    for rec in out:
        print("| " + " | ".join(f"{k}={rec[k]:.4g}" if isinstance(rec[k], float) else f"{k}={rec[k]}" for k in rec) + " |")
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
