#!/bin/bash
# GPU call 8: v4 as default; 256-row tiles and timing experiments on v4; source-level ncu of the v4 stage-3 launch
LOG=gpurun_out/r2_call8.log
. "$(dirname "$0")/r2_common.sh"
step "spconv tests"
t 300 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu >> $LOG 2>&1
step "GEMM per layer: v4 default vs 256-row tiles vs no-weights / no-A timing experiments"
t 300 python scripts/bench_gemm_layers.py --opt gemm_mt=2,gemm_debug=1,gemm_debug=4 >> $LOG 2>&1
step "ncu --set full: v4 stage-3 (launch 35) and stage-2 (launch 30)"
for spec in 35:gemm4_128 30:gemm4_64; do
  idx=${spec%%:*}; name=${spec##*:}
  t 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s $idx -c 1 -o gpurun_out/${name}_r02 -f \
      python scripts/profile_spconv.py >> $LOG 2>&1
  ncu -i gpurun_out/${name}_r02.ncu-rep --page raw --csv > gpurun_out/${name}_r02_raw.csv 2>> $LOG
done
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c8_bench_default.json 2>> $LOG
step done
grep -v "^E   \|^$\|==PROF==\|==WARNING==" $LOG | tail -n 70
for f in gpurun_out/r2c8_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
