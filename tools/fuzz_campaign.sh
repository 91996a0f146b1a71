#!/bin/bash
# Multi-seed fuzz campaign of the batched path against the oracle (tests/test_gpu_parity.py::test_batched_fuzz_shapes).
# Usage: tools/fuzz_campaign.sh <tag> <iters> <seed> [<seed> ...]; one log per seed under gpurun_out/fuzz_<tag>/,
# failing cases are kept as gpurun_out/fuzz_<tag>/fail_<seed>.json.
cd "$(dirname "$0")/.."
tag=$1; iters=$2; shift 2
out=gpurun_out/fuzz_$tag
mkdir -p "$out"
: > "$out/summary.txt"
for seed in "$@"; do
    rm -f gpurun_out/fuzz_fail.json
    start=$(date +%s)
    BRA_FUZZ_SEED=$seed BRA_FUZZ_ITERS=$iters timeout ${PER_SEED_TIMEOUT:-900} \
        python -m pytest "tests/test_gpu_parity.py::test_batched_fuzz_shapes" -x -q -m gpu > "$out/seed_$seed.log" 2>&1
    rc=$?
    [ -f gpurun_out/fuzz_fail.json ] && mv gpurun_out/fuzz_fail.json "$out/fail_$seed.json"
    echo "seed=$seed iters=$iters rc=$rc seconds=$(( $(date +%s) - start )) $(tail -1 "$out/seed_$seed.log")" >> "$out/summary.txt"
done
cat "$out/summary.txt"
