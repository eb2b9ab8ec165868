#!/bin/bash
# GPU call 7: v4 schedule (dedicated epilogue warps, double-buffered TMEM) bring-up + A/B
LOG=gpurun_out/r2_call7.log
. "$(dirname "$0")/r2_common.sh"
step "v4 bit-identity + spconv tests"
t 300 python -m pytest tests/test_spconv_gpu.py tests/test_abi.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: v4"
t 300 python scripts/bench_gemm_layers.py --opt gemm_v4=1 >> $LOG 2>&1
step "bench default / v4 sparse / v4 all"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c7_bench_default.json 2>> $LOG
ISF_OPTS=gemm_v4=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c7_bench_v4.json 2>> $LOG
ISF_OPTS=gemm_v4=2 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c7_bench_v4all.json 2>> $LOG
step "suite with v4 everywhere (numerics of every GEMM user)"
ISF_OPTS=gemm_v4=2 t 600 python -m pytest tests/test_fusion_gpu.py tests/test_gemm_gpu.py tests/test_voxel_gpu.py -q -m gpu >> $LOG 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 70
for f in gpurun_out/r2c7_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
