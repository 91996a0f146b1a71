#!/usr/bin/env python3
"""Condense an `ncu --page raw --csv` export into one line per launch (the numbers DESIGN.md quotes)."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "winst"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel | " + " | ".join(n for _, n in COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        name = name.replace("void ", "").split("(")[0][:52]
        vals = []
        for c, _ in COLS:
            if c in idx:
                v = r[idx[c]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.0f}" if f >= 100 else f"{f:.2f}"
                except ValueError:
                    pass
                u = units[idx[c]]
                vals.append(v + (u if u in ("Kbyte", "Gbyte") else ""))
            else:
                vals.append("-")
        print(name + " | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
