#!/bin/bash
# GPU call 3 of round 2: conv cache fix, TMA-gather sparse-conv schedule (v3), split-row LayerNorm epilogue, overflow fix,
# bench-workload parity; then the whole default suite and benches.
LOG=gpurun_out/r2_call3.log
. "$(dirname "$0")/r2_common.sh"
step "conv + spconv + gemm tests"
t 300 python -m pytest tests/test_conv2d_own.py tests/test_gemm_gpu.py -x -q -m gpu >> $LOG 2>&1
t 300 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: v3"
t 300 python scripts/bench_gemm_layers.py --opt gemm_v3=1 >> $LOG 2>&1
step "fusion tests (with measured errors)"
t 600 python -m pytest tests/test_fusion_gpu.py -x -q -s -m gpu >> $LOG 2>&1
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench: default, v3, old LN epilogue"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c3_bench_default.json 2>> $LOG
ISF_OPTS=gemm_v3=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c3_bench_v3.json 2>> $LOG
ISF_OPTS=gemm_ln_fast=0 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c3_bench_oldln.json 2>> $LOG
step done
grep -v "^E   \|^$" $LOG | tail -n 170
for f in gpurun_out/r2c3_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
