#!/bin/bash
# GPU call 21: accumulators interleaved in the MMA issue order (sparse GEMM MT = 2, dense conv MT = 2): tests, per-layer table, bench
LOG=gpurun_out/r2_call21.log
. "$(dirname "$0")/r2_common.sh"
step "spconv + conv tests"
t 300 python -m pytest tests/test_spconv_gpu.py tests/test_conv2d_own.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer"
t 400 python scripts/bench_gemm_layers.py --opt gemm_pair=0,gemm_mt=2,gemm_mt=1,gemm_debug=5 >> $LOG 2>&1
step "bench"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c21_bench_default.json 2>> $LOG
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 50
for f in gpurun_out/r2c21_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
