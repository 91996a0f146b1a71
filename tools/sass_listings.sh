#!/bin/bash
# SASS listings of the sort, CRC, Huffman pack and RLE emit kernels plus the per-kernel mnemonic summary (profiles/r2_sass_*).
cd "$(dirname "$0")/.."
LIB=br-archive_b200/libbra_b200.so
for pair in rs_onesweep:sort crc_raw:crc huf_pack:huf_pack rle_enc_out:rle_enc_out; do
    k=${pair%%:*}; o=${pair##*:}
    cuobjdump -sass $LIB 2>/dev/null | awk -v pat="$k" '/Function :/ {p = ($0 ~ pat)} p' > profiles/r2_sass_$o.txt
done
python - <<'PY'
import re, subprocess, collections
out = subprocess.run(['cuobjdump', '-sass', 'br-archive_b200/libbra_b200.so'], capture_output=True, text=True).stdout
fn = None
per = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        per[fn][m.group(1).split('.')[0]] += 1
names = subprocess.run(['c++filt'], input='\n'.join(per.keys()), capture_output=True, text=True).stdout.splitlines()
with open('profiles/r2_sass_summary.txt', 'w') as f:
    f.write('SASS mnemonic counts per kernel of br-archive_b200/libbra_b200.so (cuobjdump -sass, sm_100a); bulk-asynchronous copy = UBLKCP, mbarrier = SYNCS\n')
    f.write(f"{'kernel':70s} {'instr':>6s} {'UBLKCP':>6s} {'SYNCS':>5s} {'MATCH':>5s} {'ATOMS':>5s} {'STL':>4s} {'LDL':>4s}\n")
    for (fnm, c), nm in sorted(zip(per.items(), names), key=lambda x: -sum(x[0][1].values())):
        nm = nm.replace('bra::', '').split('(')[0][:70]
        f.write(f"{nm:70s} {sum(c.values()):6d} {c['UBLKCP']:6d} {c['SYNCS']:5d} {c['MATCH']:5d} {c['ATOMS']:5d} {c['STL']:4d} {c['LDL']:4d}\n")
PY
