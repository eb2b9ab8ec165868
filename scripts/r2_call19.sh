#!/bin/bash
# GPU call 19: ncu --set full of the stage-3 / stage-4 launches on the CTA-pair schedule (tc::gemm5_kernel) + launch list of a frame
LOG=gpurun_out/r2_call19.log
. "$(dirname "$0")/r2_common.sh"
step "ncu --set full: stage-3 SubM 128->128 and stage-4 SubM 256->256 (pair kernel)"
t 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 35 -c 1 -o gpurun_out/gemm5_128_r02 -f python scripts/profile_spconv.py >> $LOG 2>&1
t 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 40 -c 1 -o gpurun_out/gemm5_256_r02 -f python scripts/profile_spconv.py >> $LOG 2>&1
for k in gemm5_128 gemm5_256; do
  ncu -i gpurun_out/${k}_r02.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${k}_r02.csv 2>> $LOG
  python scripts/ncu_source_top.py gpurun_out/${k}_r02.ncu-rep > gpurun_out/ncu_source_${k}_r02.txt 2>> $LOG
done
step "launch list of one warm eager frame"
t 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_pair.csv python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
step done
tail -n 5 $LOG; ls -la gpurun_out | tail -8; head -c 1500 gpurun_out/ncu_source_gemm5_128_r02.txt
