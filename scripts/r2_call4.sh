#!/bin/bash
# GPU call 4 of round 2: fixed tests, gather micro-benchmark, v3 per-layer record, source-level ncu of the stage-3 / stage-2
# implicit-GEMM launches (stall reasons per line), launch list of a frame.
LOG=gpurun_out/r2_call4.log
. "$(dirname "$0")/r2_common.sh"
step "re-run of the tests that failed in call 3"
t 300 python -m pytest tests/test_spconv_gpu.py -q -m gpu -k "bitmap or overflow" >> $LOG 2>&1
t 400 python -m pytest tests/test_fusion_gpu.py -q -s -m gpu -k "bench_workload" >> $LOG 2>&1
step "gather micro-benchmark"
t 300 python scripts/bench_gather.py >> $LOG 2>&1
step "GEMM per layer: v3 (record)"
t 300 python scripts/bench_gemm_layers.py --opt gemm_v3=1 >> $LOG 2>&1
step "ncu --set full: stage-3 SubM 128->128 (launch 35), stage-2 SubM 64->64 (launch 30), stage-4 (launch 40)"
for spec in 35:gemm128 30:gemm64 40:gemm256; do
  idx=${spec%%:*}; name=${spec##*:}
  t 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s $idx -c 1 -o gpurun_out/${name}_r02 -f \
      python scripts/profile_spconv.py >> $LOG 2>&1
  ncu -i gpurun_out/${name}_r02.ncu-rep --page raw --csv > gpurun_out/${name}_r02_raw.csv 2>> $LOG
  ncu -i gpurun_out/${name}_r02.ncu-rep --page source --csv > gpurun_out/${name}_r02_source.csv 2>> $LOG
done
step "launch list of one warm eager frame"
t 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
step done
ls -la gpurun_out | tail -20
grep -v "^E   \|^$" $LOG | tail -n 120
