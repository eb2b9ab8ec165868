#!/bin/bash
# First GPU call of round 2 (one gpurun, ~15 min): bring up the opt-in paths written at the end of round 1 without GPU
# time.  Every step runs under its own timeout and logs under gpurun_out/; the steps are ordered by risk — the paths
# that only re-use validated kernels first, the new tcgen05 schedule (v2, bounded waits that trap) last, so that a
# protocol error there cannot cost the other measurements.
#   gpurun --t 1800 -- 'bash scripts/r2_bringup.sh'
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_bringup.log
: > $LOG
step() { echo "== [$SECONDS s] $1" | tee -a $LOG; }
# every command runs under its own limit AND the script's overall budget (seconds of box time this call may use)
BUDGET=${R2_BUDGET:-1500}
t() {
  local lim=$1; shift
  local left=$((BUDGET - SECONDS))
  if [ $left -lt 20 ]; then echo "-- skipped (budget spent): $*" | tee -a $LOG; return 99; fi
  [ $lim -gt $left ] && lim=$left
  timeout $lim "$@"
}
export ISF_EXPERIMENTAL=1
step "tile order (index kernels + the validated GEMM through out_rows / res_rows)"
t 300 python -m pytest tests/test_tile_order.py -x -q -m gpu             >> $LOG 2>&1
step "CUDA path directly against the reference-run fixtures (small grids, head, image neck, detector e2e)"
t 400 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu    >> $LOG 2>&1
step "head with the own split-K attention kernel"
ISF_HEAD_ATTN=own t 300 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu -k "head or detector" >> $LOG 2>&1
step "Swin with the own shifted-window attention kernel"
t 300 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu -k "swin" >> $LOG 2>&1
step "configs[4] geometry (0.05 m voxels, 256x256 BEV)"
t 300 python -m pytest tests/test_config5_gpu.py -x -q -m gpu            >> $LOG 2>&1
step "size-independent properties at the full-frame sizes (storage-order invariance, identity, scaling, output sites)"
t 400 python -m pytest tests/test_properties.py -x -q -m gpu             >> $LOG 2>&1
step "input contract: multi-sweep merge kernels"
t 200 python -m pytest tests/test_input_contract.py -x -q -m gpu        >> $LOG 2>&1
step "input-contract kernels at full size: time and GB/s"
t 200 python scripts/bench_input_pipeline.py > gpurun_out/r2_input_pipeline.json 2>> $LOG
step "backward kernels (dynamic scatter K3b, deformable attention K9b)"
t 300 python -m pytest tests/test_backward_ops.py -x -q -m gpu           >> $LOG 2>&1
step "configs[4] at full size: 1 M points, eager, per-stage times"
t 300 python scripts/bench_config5.py > gpurun_out/r2_config5.json 2>> $LOG
step "own dense 3x3 convs on the implicit-GEMM kernel"
t 300 python -m pytest tests/test_conv2d_own.py -x -q -m gpu             >> $LOG 2>&1
step "per-layer A/B: tile orders on the validated kernel"
t 400 python scripts/bench_gemm_layers.py --tile-order 12,16,31         >> $LOG 2>&1
unset ISF_EXPERIMENTAL
step "neighbour-table kernel A/B (ncu launch list, second frame = warm)"
for opt in 0 1; do
  ISF_NBR_ROWS=$opt t 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nbr_bitmap --csv \
    --log-file gpurun_out/r2_nbr_rows_$opt.csv python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
done
step "bench: default, tile order 31 / 16 (+ row kernel), own convs"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline                                > gpurun_out/r2_bench_default.json 2>> $LOG
ISF_TILE_ORDER=31 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline              > gpurun_out/r2_bench_ord31.json 2>> $LOG
ISF_TILE_ORDER=16 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline              > gpurun_out/r2_bench_ord16.json 2>> $LOG
ISF_TILE_ORDER=31 ISF_NBR_ROWS=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ord31_rows.json 2>> $LOG
ISF_OWN_CONV=1 t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline                 > gpurun_out/r2_bench_ownconv.json 2>> $LOG
step "whole-config forward (images + points -> boxes), default and with the own attention kernels"
t 300 python scripts/bench_full_forward.py                                                   > gpurun_out/r2_full_forward.json 2>> $LOG
t 300 python scripts/bench_full_forward.py --raw                                             > gpurun_out/r2_full_forward_raw.json 2>> $LOG
ISF_HEAD_ATTN=own ISF_SWIN_ATTN=own t 300 python scripts/bench_full_forward.py               > gpurun_out/r2_full_forward_own_attn.json 2>> $LOG
export ISF_EXPERIMENTAL=1
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
tail -n 80 $LOG
for f in gpurun_out/r2_bench_*.json; do echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])" 2>/dev/null; done
