#!/bin/bash
# GPU call 12: zero-skip of the v4 producers (bit-identity + per-layer A/B + bench)
LOG=gpurun_out/r2_call12.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests (incl. zero-skip bit-identity)"
t 300 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: zero-skip"
t 300 python scripts/bench_gemm_layers.py --opt gemm_zskip=1 >> $LOG 2>&1
step "bench default vs zero-skip"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c12_bench_default.json 2>> $LOG
ISF_OPTS=gemm_zskip=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c12_bench_zskip.json 2>> $LOG
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 45
for f in gpurun_out/r2c12_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
