#!/bin/bash
# GPU call 36: dynamic tile scheduling (v4 kernels: 32- / 64-column layers): bit-identity, per-layer A/B, bench A/B
LOG=gpurun_out/r2_call36.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests"
t 200 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: dynamic (default)"
t 200 python scripts/bench_gemm_layers.py >> $LOG 2>&1
step "GEMM per layer: static"
ISF_DYNAMIC_TILES=0 t 200 python scripts/bench_gemm_layers.py >> $LOG 2>&1
step "bench A/B"
t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c36_bench_dyn.json 2>> $LOG; echo "rc=$?" >> $LOG
ISF_DYNAMIC_TILES=0 t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c36_bench_static.json 2>> $LOG; echo "rc=$?" >> $LOG
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 64
for f in gpurun_out/r2c36_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
