#!/bin/bash
# GPU call 16: warp-per-chunk producers that own their smem stage (fix of call 14), per-layer A/B, suite, bench
LOG=gpurun_out/r2_call16.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests (bit-identity of every variant)"
t 300 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: new defaults vs each change switched off"
t 400 python scripts/bench_gemm_layers.py --opt gemm_wpc=0,gemm_wpc=2,gemm_rev=0,gemm_v4_32=0,gemm_v4_32=2,gemm_mt=2,gemm_debug=5 >> $LOG 2>&1
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_bench_default.json 2>> $LOG
ISF_OPTS=gemm_wpc=0,gemm_rev=0,gemm_v4_32=0 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_bench_old.json 2>> $LOG
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 70
for f in gpurun_out/r2c16_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
