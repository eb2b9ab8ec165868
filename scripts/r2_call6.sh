#!/bin/bash
# GPU call 6: suite, 256-row tiles for the 32/256-column layers (A/B), HSF-only bench (configs[2]), configs[4] at N=1,
# ncu of the dense conv + decoder kernels, launch list of a whole frame.
LOG=gpurun_out/r2_call6.log
. "$(dirname "$0")/r2_common.sh"
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "GEMM per layer: auto 256-row tiles"
t 300 python scripts/bench_gemm_layers.py --opt gemm_mt=3 >> $LOG 2>&1
step "bench default vs gemm_mt=3"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_default.json 2>> $LOG
ISF_OPTS=gemm_mt=3 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_mt3.json 2>> $LOG
step "HSF-only (configs[2])"
t 300 python scripts/bench_hsf.py > gpurun_out/r2_hsf.json 2>> $LOG
t 300 python scripts/bench_hsf.py --levels 2 > gpurun_out/r2_hsf_2levels.json 2>> $LOG
step "configs[4] N=1"
t 300 python scripts/bench_config5.py > gpurun_out/r2_config5_n1.json 2>> $LOG
step "ncu: launch list of one warm eager frame + metric table of the new kernels"
t 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
t 400 ncu --clock-control none --metrics $METRICS -k regex:'gemm_kernel|conv3x3|ins_context|nbr_|bitmap_|vfe_|voxel|cam2bev|window_attention|i2s_|channel_attention|msda|heatmap|take_topk|tile_order|pillar|scatter|nhwc|nchw|to_bev' \
    --launch-skip 95 --launch-count 130 --csv --log-file gpurun_out/own_kernels_r02.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 70
for f in gpurun_out/r2c6_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
for f in gpurun_out/r2_hsf.json gpurun_out/r2_hsf_2levels.json gpurun_out/r2_config5_n1.json; do [ -s "$f" ] && { echo $f; tail -c 1800 $f; echo; }; done
