#!/bin/bash
# Runs every GPU parity test in its own process (a CUDA fault in one test must not poison the rest)
# and leaves one log per test plus a summary under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/tests
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
tests=$(python -m pytest tests/test_gpu_parity.py -m gpu --collect-only -q 2>/dev/null | grep "::")
: > gpurun_out/summary.txt
for t in $tests; do
    name=$(echo "$t" | sed 's/.*:://' | tr '[]' '__')
    timeout ${PER_TEST_TIMEOUT:-420} python -m pytest "$t" -x -q -m gpu > "gpurun_out/tests/$name.log" 2>&1
    rc=$?
    echo "$rc $name" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
