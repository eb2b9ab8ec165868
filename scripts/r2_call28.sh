#!/bin/bash
# GPU call 28: two MMA-issuing warps in the 256-row CTA-pair kernels (gemm_mma2); per-layer A/B; bench x3
LOG=gpurun_out/r2_call28.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests"
t 200 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer"
t 300 python scripts/bench_gemm_layers.py --opt gemm_mma2=0,gemm_pair=1,gemm_pair=256,gemm_debug=5 >> $LOG 2>&1
step "bench x3"
for i in 1 2 3; do
  t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c28_bench_$i.json 2>> $LOG; echo "rc=$?" >> $LOG
done
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 50
for f in gpurun_out/r2c28_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
