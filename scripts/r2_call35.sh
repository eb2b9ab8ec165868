#!/bin/bash
# GPU call 35: sparse-conv GEMMs launch the fewest CTAs that keep the number of tile rounds (gemm_trim): per-layer check, bench A/B x3
LOG=gpurun_out/r2_call35.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests"
t 200 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: trim on (default) / off"
t 200 python scripts/bench_gemm_layers.py --opt gemm_trim=0 >> $LOG 2>&1
step "bench A/B"
for i in 1 2 3; do
  t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c35_bench_trim_$i.json 2>> $LOG; echo "rc=$?" >> $LOG
  ISF_OPTS=gemm_trim=0 t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c35_bench_full_$i.json 2>> $LOG; echo "rc=$?" >> $LOG
done
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 34
for f in gpurun_out/r2c35_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
