#!/usr/bin/env python3
"""Kernel shares of a bench run, two ways side by side: the ncu launch list (gpu__time_duration.sum, cold cache,
serialised) and the live CUDA-event brackets bench.py reports in `kernel_ms`.

usage: kernel_shares.py <ncu launch list csv> <bench json> [title]"""
import collections
import csv
import json
import re
import sys


def main(launches, bench, title):
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"^void ", "", r[ik]).split("(")[0].replace("bra::", "").split("<")[0]
        t[name] += float(r[iv].replace(",", "")) / 1e6
        n[name] += 1
    tot = sum(t.values())
    print(f"# {title}\n")
    print(f"ncu launch list `{launches}` (gpu__time_duration.sum, cold cache, serialised) next to the live CUDA-event")
    print(f"brackets of `{bench}` (`kernel_ms`, per step).\n")
    print("| kernel | launches (ncu run) | ncu time ms | ncu share |\n|---|---|---|---|")
    for k, v in t.most_common(26):
        print(f"| {k} | {n[k]} | {v:.2f} | {v / tot * 100:.1f}% |")
    d = json.loads(open(bench).read().strip().split("\n")[-1])
    km = d["kernel_ms"]
    tot2 = sum(km.values())
    print("\n| kernel family (live) | ms per step | share |\n|---|---|---|")
    for k, v in sorted(km.items(), key=lambda kv: -kv[1])[:26]:
        print(f"| {k} | {v:.2f} | {v / tot2 * 100:.1f}% |")
    print(f"\nlive sum {tot2:.1f} ms of a {d['ms_per_step']:.1f} ms step; dominant kernel by both measures: "
          f"{t.most_common(1)[0][0]} ({t.most_common(1)[0][1] / tot * 100:.1f}% ncu) / {max(km, key=km.get)} ({max(km.values()) / tot2 * 100:.1f}% live)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "Kernel shares")
