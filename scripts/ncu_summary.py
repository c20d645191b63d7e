#!/usr/bin/env python
"""Turn an `ncu --set full` report (.ncu-rep, read here with the ncu CLI) into the markdown table kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [title] > profiles/rNN_x.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(v) * f


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    launches, units = load(rep)
    print(f"# {title}\n")
    print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one B200); read with "
          f"`ncu -i ... --page raw --csv` by scripts/ncu_summary.py.  Times under ncu are serialised and cold-cache.\n")
    for i, d in enumerate(launches):
        name = d["Kernel Name"].split("(")[0]
        print(f"## launch {i}: `{name}`\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"| {label} (`{k}`) | {d[k]} {units.get(k, '')} |")
        try:
            tr = to_bytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + to_bytes(
                d["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
            u = units["gpu__time_duration.sum"]
            sec = float(d["gpu__time_duration.sum"]) * {"s": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}[u]
            print(f"| **DRAM traffic per launch** | {tr / 1e9:.4f} GB -> {tr / sec / 1e9:.0f} GB/s under ncu |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main()
