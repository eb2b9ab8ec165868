#!/bin/bash
# GPU call 31: where the 0.7 ms between the no-load floor and the shipped GEMM goes (weights once per tile / no A copies / neither)
LOG=gpurun_out/r2_call31.log
. "$(dirname "$0")/r2_common.sh"
step "GEMM per layer: timing experiments on the shipped configuration"
t 300 python scripts/bench_gemm_layers.py --opt gemm_debug=1,gemm_debug=4,gemm_debug=5,gemm_debug=2 >> $LOG 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 32
