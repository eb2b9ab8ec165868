#!/bin/bash
# GPU call 20: tcgen05.mma issue rate by instruction shape / operand source / CTA group; M = 64 / 128 accumulator layout
LOG=gpurun_out/r2_call20.log
. "$(dirname "$0")/r2_common.sh"
step "MMA rate micro-benchmark"
t 200 python scripts/bench_mma_rate.py >> $LOG 2>&1
step done
cat $LOG
