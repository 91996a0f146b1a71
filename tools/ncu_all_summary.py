#!/usr/bin/env python3
"""Per-kernel table from tools/ncu_all.sh's raw CSV (one row per launch and metric): launches, total and mean duration,
DRAM bytes read/written, L2 bytes, occupancy, bank conflicts, registers. With --traffic ELEMS also prints DRAM bytes per
element and launch for the kernel families bench.py reports (profiles/ncu_traffic.json)."""
import csv
import json
import sys
from collections import defaultdict

FAMILY = [("rs_onesweep_kernel<0>", "radix_scatter"), ("rs_onesweep_kernel<1>", "radix_scatter_implicit"), ("rs_onesweep_kernel<2>", "radix_scatter_u8"),
          ("bwt_ranks_kernel<2>", "bwt_ranks"), ("bwt_dense_ranks_kernel", "bwt_ranks"), ("bwt_heads_stats_kernel<1>", "bwt_heads"), ("bwt_heads_kernel<1>", "bwt_heads"), ("bwt_dbl_prepare", "bwt_prepare"),
          ("bwt_finish_kernel", "bwt_finish"), ("bwt_gather_kernel", "bwt_gather"), ("mtf_lane_kernel", "mtf_apply"), ("mtf_enc_summary_kernel", "mtf_summary"),
          ("ibwt_walk_len_kernel", "ibwt_walk_len"), ("ibwt_walk_emit_kernel", "ibwt_walk_emit"), ("crc_raw_kernel", "crc32c"), ("huf_dec_sync_kernel", "huf_dec_sync"),
          ("huf_dec_write_kernel", "huf_dec_write"), ("huf_pack_kernel", "huf_pack"), ("rle_enc_out_kernel", "rle_enc_emit"),
          ("rle_dec_expand_kernel", "rle_dec_expand"), ("rle_dec_mark_kernel", "rle_dec_mark")]


def load(path):
    launches = defaultdict(dict)
    for row in csv.DictReader(l for l in open(path) if l.startswith('"')):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
        launches[(int(row["ID"]), row["Kernel Name"])][row["Metric Name"]] = v * scale
    return launches


def short(name):
    return name.replace("void ", "").replace("bra::", "").split("(")[0]


def main():
    path = sys.argv[1]
    launches = load(path)
    agg = defaultdict(lambda: defaultdict(float))
    for (_, name), m in launches.items():
        a = agg[short(name)]
        a["n"] += 1
        a["ms"] += m.get("gpu__time_duration.sum", 0)
        a["rd"] += m.get("dram__bytes_read.sum", 0)
        a["wr"] += m.get("dram__bytes_write.sum", 0)
        a["l2"] += m.get("lts__t_bytes.sum", 0)
        a["occ"] += m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0) * m.get("gpu__time_duration.sum", 0)
        a["bank"] += m.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0)
        a["regs"] = max(a["regs"], m.get("launch__registers_per_thread", 0))
        a["winst"] += m.get("smsp__inst_executed.sum", 0)
    total = sum(a["ms"] for a in agg.values())
    print(f"{'kernel':48s} {'n':>4s} {'ms':>9s} {'share':>6s} {'ms/launch':>9s} {'DRAM rd GB':>10s} {'DRAM wr GB':>10s} {'L2 GB':>8s} {'GB/s DRAM':>9s} {'occ%':>5s} {'bank confl':>11s} {'regs':>4s} {'Gwinst':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        ms = a["ms"]
        print(f"{k[:48]:48s} {int(a['n']):4d} {ms:9.3f} {100 * ms / total:5.1f}% {ms / a['n']:9.3f} {a['rd'] / 1e9:10.3f} {a['wr'] / 1e9:10.3f} {a['l2'] / 1e9:8.2f} "
              f"{(a['rd'] + a['wr']) / 1e9 / (ms / 1e3) if ms else 0:9.0f} {a['occ'] / ms if ms else 0:5.1f} {a['bank']:11.3g} {int(a['regs']):4d} {a['winst'] / 1e9:7.3f}")
    print(f"total kernel time {total:.3f} ms over {sum(int(a['n']) for a in agg.values())} launches (cold-cache, serialised; shares are what compares with the live brackets)")
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        # DRAM bytes per element of the FULL-BATCH launches only (the end-to-end stages of the same run launch the same
        # kernels on fewer blocks): per family, the launches lasting at least 93 % of the family's longest one
        elems = float(sys.argv[3])
        fam = defaultdict(list)
        for (_, name), m in launches.items():
            s = short(name)
            for pat, f in FAMILY:
                if s.startswith(pat):
                    fam[f].append((m.get("gpu__time_duration.sum", 0), m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)))
                    break
        out = {}
        for f, rows in sorted(fam.items()):
            gmax = max(g for g, _ in rows)
            full = [b for g, b in rows if g >= 0.93 * gmax]
            out[f] = round(sum(full) / len(full) / elems, 2)
        print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
