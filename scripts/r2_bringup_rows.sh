#!/bin/bash
# Second GPU call of round 2 (one gpurun, ~12 min): the rows AROUND the hot path written at the end of round 1 without GPU
# time: CUDA path against the reference-run fixtures, head / Swin with their own attention kernels, the configs[4]
# geometry, the input-contract kernels, the backward kernels, and the whole-config forward.  Same step / budget rules as
# scripts/r2_bringup.sh; every test here is gated by ISF_EXPERIMENTAL=1 until it has passed once on a GPU.
#   gpurun --timeout 1500 -- 'bash scripts/r2_bringup_rows.sh'
LOG=gpurun_out/r2_bringup_rows.log
. "$(dirname "$0")/r2_common.sh"
export ISF_EXPERIMENTAL=1
step "CUDA path directly against the reference-run fixtures (small grids, head, image neck, detector e2e)"
t 400 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu    >> $LOG 2>&1
step "head with the own split-K attention kernel"
ISF_HEAD_ATTN=own t 300 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu -k "head or detector" >> $LOG 2>&1
step "Swin with the own shifted-window attention kernel"
t 300 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu -k "swin" >> $LOG 2>&1
step "configs[4] geometry (0.05 m voxels, 256x256 BEV)"
t 300 python -m pytest tests/test_config5_gpu.py -x -q -m gpu            >> $LOG 2>&1
step "input contract: multi-sweep merge and camera preprocessing kernels"
t 200 python -m pytest tests/test_input_contract.py -x -q -m gpu        >> $LOG 2>&1
step "backward kernels (dynamic scatter K3b, deformable attention K9b)"
t 300 python -m pytest tests/test_backward_ops.py -x -q -m gpu           >> $LOG 2>&1
unset ISF_EXPERIMENTAL
step "input-contract kernels at full size: time and GB/s"
t 200 python scripts/bench_input_pipeline.py > gpurun_out/r2_input_pipeline.json 2>> $LOG
step "configs[4] at full size: 1 M points, eager, per-stage times"
t 300 python scripts/bench_config5.py > gpurun_out/r2_config5.json 2>> $LOG
step "whole-config forward (images + points -> boxes): default, from raw sensor data, with the own attention kernels"
t 300 python scripts/bench_full_forward.py                                                   > gpurun_out/r2_full_forward.json 2>> $LOG
t 300 python scripts/bench_full_forward.py --raw                                             > gpurun_out/r2_full_forward_raw.json 2>> $LOG
ISF_HEAD_ATTN=own ISF_SWIN_ATTN=own t 300 python scripts/bench_full_forward.py               > gpurun_out/r2_full_forward_own_attn.json 2>> $LOG
tail -n 60 $LOG
for f in gpurun_out/r2_input_pipeline.json gpurun_out/r2_config5.json gpurun_out/r2_full_forward*.json; do [ -s "$f" ] && { echo $f; tail -c 1500 $f; echo; }; done
