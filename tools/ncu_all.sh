#!/bin/bash
# One ncu pass over EVERY kernel launch of one encode+decode step (256 MiB per shape so that the replays stay short):
# duration, DRAM and L2 bytes, occupancy, shared-memory bank conflicts, registers, executed warp instructions.
# Output: gpurun_out/<tag>_ncu_all_<shape>.csv (raw, one row per launch and metric). Summarise with tools/ncu_all_summary.py.
cd "$(dirname "$0")/.."
tag=${1:-r2}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,smsp__inst_executed.sum,launch__grid_size,launch__block_size
for shape in ${SHAPES:-text random repeat251}; do
    extra=""
    [ "$shape" = repeat251 ] && extra="--block-kib 8192"
    timeout ${NCU_TIMEOUT:-900} ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_ncu_all_$shape.csv \
        python bench.py --workload $shape --size-mib 256 --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --parity-blocks 0 $extra > gpurun_out/${tag}_ncu_all_$shape.log 2>&1
    echo "ncu $shape rc=$?"
done
