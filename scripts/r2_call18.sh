#!/bin/bash
# GPU call 18: CTA-pair schedule with plain remote arrives + st.async mask exchange (no MEMBAR.GPU / CCTL.IVALL per chunk)
LOG=gpurun_out/r2_call18.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests (pair kernel is the default; bit-identity of every variant)"
t 300 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: pairs vs v4"
t 400 python scripts/bench_gemm_layers.py --opt gemm_pair=0,gemm_pair=256,gemm_pair=128,gemm_rev=1,gemm_mt=1,gemm_debug=5 >> $LOG 2>&1
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c18_bench_default.json 2>> $LOG
ISF_OPTS=gemm_pair=0 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c18_bench_v4.json 2>> $LOG
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 70
for f in gpurun_out/r2c18_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
