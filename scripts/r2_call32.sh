#!/bin/bash
# GPU call 32: L2 prefetch of the next tile's gathered rows (gemm_pf): per-layer A/B and bench
LOG=gpurun_out/r2_call32.log
. "$(dirname "$0")/r2_common.sh"
step "GEMM per layer: gemm_pf"
t 300 python scripts/bench_gemm_layers.py --opt gemm_pf=1 >> $LOG 2>&1
step "bench: default vs gemm_pf=1"
t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c32_bench_default.json 2>> $LOG; echo "rc=$?" >> $LOG
ISF_OPTS=gemm_pf=1 t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c32_bench_pf.json 2>> $LOG; echo "rc=$?" >> $LOG
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 34
for f in gpurun_out/r2c32_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
