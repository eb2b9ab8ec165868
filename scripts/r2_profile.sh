#!/bin/bash
# Round-2 profile capture (one gpurun, ~8 min), run AFTER scripts/r2_bringup.sh has shown which variant to keep:
#   gpurun --timeout 1200 -- 'ISF_TILE_ORDER=31 bash scripts/r2_profile.sh r02'      (env = the variant being profiled)
# Produces, under gpurun_out/ (copy the summaries into profiles/ with the round tag):
#   launches_<tag>.csv          ncu launch list of one warm frame of `bench.py --eager-only` (per-launch gpu time)
#   gemm_<tag>.ncu-rep          `--set full` captures of the two dominant launches (stage-3 SubM 128->128, stage-4 256->256)
#   gemm_<tag>_raw.txt          the roofline metrics of those captures
#   own_kernels_<tag>.csv       metric table of every library kernel of one frame
# Numbers printed by a run under ncu are never bench values (B200_PROFILING.md).
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 > gpurun_out/profile_$TAG.log 2>&1
# profile_spconv.py launches 2 DynamicVFE GEMMs, then the sparse encoder twice (21 launches each: conv_input, 4+1 per
# stage 1-3, 4 of stage 4, conv_out).  Second (warm) pass: index 2 + 21 + 12 = 35 is a stage-3 SubM 128->128 with residual,
# 2 + 21 + 17 = 40 a stage-4 SubM 256->256 with residual (check the kernel's grid / rows in the report).
ncu --set full --clock-control none --import-source on -k regex:gemm -s 35 -c 1 -o gpurun_out/gemm128_$TAG -f \
    python scripts/profile_spconv.py >> gpurun_out/profile_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 40 -c 1 -o gpurun_out/gemm256_$TAG -f \
    python scripts/profile_spconv.py >> gpurun_out/profile_$TAG.log 2>&1
for k in gemm128 gemm256; do
  ncu -i gpurun_out/${k}_$TAG.ncu-rep --page raw --csv > gpurun_out/${k}_${TAG}_raw.csv 2>> gpurun_out/profile_$TAG.log
done
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
ncu --clock-control none --metrics $METRICS -k regex:'gemm|nbr_|bitmap_|vfe_|voxel|cam2bev|window_attention|i2s_|channel_attention|msda|heatmap|take_topk|tile_order|pillar|scatter|nhwc|nchw|to_bev' \
    --launch-skip 85 --launch-count 120 --csv --log-file gpurun_out/own_kernels_$TAG.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> gpurun_out/profile_$TAG.log 2>&1
ls -la gpurun_out | tail -12
# input-contract kernels (multi-sweep merge, camera resample + normalise) at the config's sizes: time + DRAM bytes per launch
ncu --clock-control none --metrics $METRICS -k regex:'merge_|image_' --launch-count 40 --csv --log-file gpurun_out/input_kernels_$TAG.csv \
    python scripts/bench_input_pipeline.py --reps 2 >> gpurun_out/profile_$TAG.log 2>&1
