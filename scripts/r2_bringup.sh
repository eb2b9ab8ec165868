#!/bin/bash
# First GPU call of round 2 (one gpurun, ~10 min): bring up the opt-in kernels written at the end of round 1 without
# GPU time, each step under its own timeout so that a protocol error cannot hang the box, everything logged under
# gpurun_out/.   gpurun --timeout 1500 -- 'bash scripts/r2_bringup.sh'
set -u
mkdir -p gpurun_out
export ISF_EXPERIMENTAL=1
echo "== tile order bring-up"      | tee gpurun_out/r2_bringup.log
timeout 300 python -m pytest tests/test_tile_order.py -x -q -m gpu            >> gpurun_out/r2_bringup.log 2>&1
echo "== v2 GEMM bring-up"         | tee -a gpurun_out/r2_bringup.log
timeout 300 python -m pytest tests/test_gemm_v2_gpu.py -x -q -m gpu           >> gpurun_out/r2_bringup.log 2>&1
echo "== direct parity against the reference fixtures" | tee -a gpurun_out/r2_bringup.log
timeout 300 python -m pytest tests/test_reference_fixtures_gpu.py -x -q -m gpu >> gpurun_out/r2_bringup.log 2>&1
echo "== configs[4] geometry (0.05 m voxels, 256x256 BEV)" | tee -a gpurun_out/r2_bringup.log
timeout 300 python -m pytest tests/test_config5_gpu.py -x -q -m gpu            >> gpurun_out/r2_bringup.log 2>&1
echo "== own dense 3x3 convs on the implicit-GEMM kernel" | tee -a gpurun_out/r2_bringup.log
timeout 300 python -m pytest tests/test_conv2d_own.py -x -q -m gpu             >> gpurun_out/r2_bringup.log 2>&1
echo "== per-layer A/B"            | tee -a gpurun_out/r2_bringup.log
timeout 400 python scripts/bench_gemm_layers.py --v2 --tile-order 12,16,31    >> gpurun_out/r2_bringup.log 2>&1
unset ISF_EXPERIMENTAL
echo "== neighbour-table kernel A/B (launch list, second frame = warm)" | tee -a gpurun_out/r2_bringup.log
for opt in 0 1; do
  ISF_NBR_ROWS=$opt timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nbr_bitmap --csv \
    --log-file gpurun_out/r2_nbr_rows_$opt.csv python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> gpurun_out/r2_bringup.log 2>&1
done
echo "== bench: default, then tile order 31 / 16" | tee -a gpurun_out/r2_bringup.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline           > gpurun_out/r2_bench_default.json 2>> gpurun_out/r2_bringup.log
ISF_TILE_ORDER=31 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ord31.json 2>> gpurun_out/r2_bringup.log
ISF_TILE_ORDER=16 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ord16.json 2>> gpurun_out/r2_bringup.log
ISF_OWN_CONV=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ownconv.json 2>> gpurun_out/r2_bringup.log
tail -n 60 gpurun_out/r2_bringup.log
for f in gpurun_out/r2_bench_*.json; do echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])" 2>/dev/null; done
