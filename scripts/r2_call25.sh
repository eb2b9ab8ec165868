#!/bin/bash
# GPU call 25: CTA-pair kernels at ONE CTA per SM (256-row CTA tiles for 64 / 128 columns); repeated bench runs with pairs on
LOG=gpurun_out/r2_call25.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests"
ISF_OPTS=gemm_pair=1 t 200 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer (pairs on: 1 CTA/SM vs 2 CTAs/SM; pairs only for 256 / 128+ columns)"
ISF_OPTS=gemm_pair=1 t 300 python scripts/bench_gemm_layers.py --opt gemm_mt=1,gemm_pair=256,gemm_pair=128,gemm_pair=0 >> $LOG 2>&1
step "bench x6, pairs on"
for i in 1 2 3 4 5 6; do
  ISF_OPTS=gemm_pair=1 t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c25_bench_pair$i.json 2>> $LOG; echo "rc=$?" >> $LOG
done
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 50
for f in gpurun_out/r2c25_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
