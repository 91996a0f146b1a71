#!/usr/bin/env python3
"""Hot SASS instructions of one kernel from an `ncu --page source --csv` export: share of stall samples, executions, top stall reasons."""
import csv
import sys


def main(path, frac=0.006):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > idx['# Samples'] and r[idx['# Samples']].isdigit()]
    tot = sum(int(r[idx['# Samples']]) for r in data)
    print(rows[0][1], '| samples', tot, '| instructions', len(data))
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {k: 0 for k in reasons}
    for i, r in enumerate(data):
        s = int(r[idx['# Samples']])
        for k in reasons:
            agg[k] += int(r[idx[k]] or 0)
        if s > tot * frac:
            st = {k: int(r[idx[k]] or 0) for k in reasons}
            top = sorted(st.items(), key=lambda x: -x[1])[:2]
            print(f"{i:5d} {r[idx['Source']].strip()[:64]:64s} {100 * s / tot:5.1f}% exec {r[idx['Instructions Executed']]:>9s} {top}")
    print('stall totals:', sorted(((k, v) for k, v in agg.items() if v), key=lambda x: -x[1])[:8])


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.006)
