#!/bin/bash
# GPU call 13: timing experiments on the v4 schedule (which bytes pace it): weights once per tile, no A copies, both; 256-row tiles at N=128
LOG=gpurun_out/r2_call13.log
. "$(dirname "$0")/r2_common.sh"
step "GEMM per layer: timing experiments on v4"
t 400 python scripts/bench_gemm_layers.py --opt gemm_debug=1,gemm_debug=4,gemm_debug=5,gemm_mt=2 >> $LOG 2>&1
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 45
