#!/bin/bash
# GPU call 29: final evidence of the round: ncu --set full of the shipped stage-3 / stage-4 launches, metric table + launch list of
# one eager frame, whole suite, smoke(), full bench line (CPU baseline) + reference arm
LOG=gpurun_out/r2_call29.log
. "$(dirname "$0")/r2_common.sh"
BUDGET=1700
step "default GPU suite"
t 600 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "smoke()"
t 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
step "ncu --set full: stage-3 SubM 128->128 (launch 35) and stage-4 SubM 256->256 (launch 40), shipped kernels"
for spec in 35:gemm5_128 40:gemm5_256; do
  idx=${spec%%:*}; name=${spec##*:}
  t 200 ncu --set full --clock-control none --import-source on -k regex:gemm -s $idx -c 1 -o gpurun_out/${name}_final_r02 -f \
      python scripts/profile_spconv.py >> $LOG 2>&1
  ncu -i gpurun_out/${name}_final_r02.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${name}_final_r02.csv 2>> $LOG
  python scripts/ncu_source_top.py gpurun_out/${name}_final_r02.ncu-rep > gpurun_out/ncu_source_${name}_final_r02.txt 2>> $LOG
done
step "metric table of every library kernel of one warm eager frame"
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
t 300 ncu --clock-control none --metrics $METRICS -k regex:'gemm|conv3x3|decoder|nbr_|bitmap_|vfe_|voxel|cam2bev|window_attention|i2s_|channel_attention|msda|heatmap|take_topk|tile_order|pillar|scatter|nhwc|nchw|to_bev' \
    --launch-skip 85 --launch-count 130 --csv --log-file gpurun_out/own_kernels_metrics_final_r02.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
step "bench: full default line, reference arm"
t 400 python bench.py > gpurun_out/r2c29_bench_full.json 2>> $LOG; echo "rc=$?" >> $LOG
t 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c29_bench_reference.json 2>> $LOG
step done
grep -v "^E   \|^$\|==PROF==\|==WARNING==\|^\[bench" $LOG | tail -n 30
python -c "import json; d=json.loads(open('gpurun_out/r2c29_bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d.get('cpu_baseline'), d.get('clocks'))"
tail -c 500 gpurun_out/r2c29_bench_reference.json
exit 0
