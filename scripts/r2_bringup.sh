#!/bin/bash
# First GPU call of round 2 (one gpurun, ~15 min): the HOT PATH's opt-in variants written at the end of round 1 without GPU
# time — neighbour-pattern tile order, row-per-thread neighbour kernel, own dense convs, then the warp-specialised GEMM
# schedule (v2).  Every step runs under its own timeout and the script's overall budget (R2_BUDGET seconds, default 1500) and
# logs under gpurun_out/; the paths that only re-use validated kernels come first, v2 (new barrier protocol; its waits are
# bounded and trap) last, so that an error there cannot cost the other measurements.
#   gpurun --timeout 1800 -- 'bash scripts/r2_bringup.sh'          then: scripts/r2_bringup_rows.sh, scripts/r2_profile.sh
LOG=gpurun_out/r2_bringup.log
. "$(dirname "$0")/r2_common.sh"
export ISF_EXPERIMENTAL=1
step "tile order (index kernels + the validated GEMM through out_rows / res_rows)"
t 300 python -m pytest tests/test_tile_order.py -x -q -m gpu             >> $LOG 2>&1
step "per-layer A/B: tile orders on the validated kernel"
t 400 python scripts/bench_gemm_layers.py --tile-order 12,16,31         >> $LOG 2>&1
unset ISF_EXPERIMENTAL
step "neighbour-table kernel A/B (ncu launch list, second frame = warm)"
for opt in 0 1; do
  ISF_NBR_ROWS=$opt t 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nbr_bitmap --csv \
    --log-file gpurun_out/r2_nbr_rows_$opt.csv python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
done
step "bench: default, tile order 31 / 16 (+ row kernel)"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline                                > gpurun_out/r2_bench_default.json 2>> $LOG
ISF_TILE_ORDER=31 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline              > gpurun_out/r2_bench_ord31.json 2>> $LOG
ISF_TILE_ORDER=16 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline              > gpurun_out/r2_bench_ord16.json 2>> $LOG
ISF_TILE_ORDER=31 ISF_NBR_ROWS=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ord31_rows.json 2>> $LOG
export ISF_EXPERIMENTAL=1
step "own dense 3x3 convs on the implicit-GEMM kernel"
t 300 python -m pytest tests/test_conv2d_own.py -x -q -m gpu             >> $LOG 2>&1
unset ISF_EXPERIMENTAL
ISF_OWN_CONV=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline                 > gpurun_out/r2_bench_ownconv.json 2>> $LOG
export ISF_EXPERIMENTAL=1
step "size-independent properties at the full-frame sizes (storage-order invariance, identity, scaling, output sites)"
t 400 python -m pytest tests/test_properties.py -x -q -m gpu             >> $LOG 2>&1
step "v2 warp-specialised GEMM schedule: bring-up tests (bounded waits trap instead of hanging)"
t 300 python -m pytest tests/test_gemm_v2_gpu.py -x -q -m gpu            >> $LOG 2>&1 && {
  step "per-layer A/B: v2 alone and with tile orders"
  t 400 python scripts/bench_gemm_layers.py --v2 --tile-order 16,31     >> $LOG 2>&1
  unset ISF_EXPERIMENTAL
  step "bench: v2 on the sparse convs / on every GEMM, with tile order 31"
  ISF_GEMM_V2=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline                   > gpurun_out/r2_bench_v2.json 2>> $LOG
  ISF_GEMM_V2=2 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline                   > gpurun_out/r2_bench_v2all.json 2>> $LOG
  ISF_GEMM_V2=2 ISF_TILE_ORDER=31 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_v2all_ord31.json 2>> $LOG
}
summarise
