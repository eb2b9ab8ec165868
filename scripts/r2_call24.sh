#!/bin/bash
# GPU call 24: one elect.sync per K chunk (umma_chunk) in every tcgen05 kernel
LOG=gpurun_out/r2_call24.log
. "$(dirname "$0")/r2_common.sh"
step "MMA rate micro-benchmark (legacy vs elected issue form)"
t 120 python scripts/bench_mma_rate.py > gpurun_out/mma_rate_r02.txt 2>&1
step "spconv + conv tests"
t 300 python -m pytest tests/test_spconv_gpu.py tests/test_conv2d_own.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer"
t 300 python scripts/bench_gemm_layers.py --opt gemm_pair=1,gemm_mt=2,gemm_mt=1,gemm_debug=5 >> $LOG 2>&1
step "default GPU suite"
t 600 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench"
t 150 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_bench_default.json 2>> $LOG; echo "rc=$?" >> $LOG
ISF_OPTS=gemm_pair=1 t 150 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_bench_pair.json 2>> $LOG; echo "rc=$?" >> $LOG
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 60
for f in gpurun_out/r2c24_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
