"""CPU oracle for the IS-Fusion fusion-neck hot path.

TEST INFRASTRUCTURE ONLY.  This package is a line-by-line CPU restatement of the
reference algorithm (yinjunbo/IS-Fusion @ 86c882a, files cited per function as
``ref: path:lines`` relative to the reference checkout).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker / the CPU
baseline — never as the product path.  The product (``is-fusion_b200``) must
fail loudly when its CUDA library is missing; it never falls back to this code.

Pinning status:
  * dynamic / hard voxelization  : PINNED against the reference's own C++
    (``oracle/_ref/libref_voxel.so`` compiled from
    mmdet3d/ops/voxel/src/voxelization_cpu.cpp where it lies) and against the
    reference's numpy ``VoxelGenerator`` known-answer test
    (tests/test_models/test_voxel_encoder/test_voxel_generator.py:7-23).
  * dynamic scatter              : pinned by the brute-force contract of
    tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:57-72.
  * MSDA                         : pinned by ops/functions/ms_deform_attn_func.py
    (imported from the reference to generate tests/golden/msda_*.npz).
  * sparse-conv rulebook + arithmetic: PINNED against the reference's vendored
    spconv-1.x CPU generator (``oracle/_ref/libref_spconv.so`` compiled from
    mmdet3d/ops/bevfusion-ops/spconv/include/spconv/geometry.h where it lies;
    fixture tests/golden/spconv_ref_cpp.npz) and cross-checked against dense
    F.conv3d.
  * DynamicVFE, SparseEncoder, SST input layer / SSTv2, camera->BEV, conv_fusion,
    instance-guided fusion, SECONDV2 staging, SECONDFPN: PINNED by OUTPUTS OF THE
    REFERENCE'S OWN PYTHON FILES executed unmodified from /root/reference in the
    build container (tests/golden/ref_shim.py + make_golden_ref.py -> ref_*.npz,
    checked by tests/test_oracle_vs_reference.py).  The reference's tests hold
    no golden vectors for these rows; the third-party pieces it needs and that
    are absent here (mmcv builders, spconv layer, compiled MSDA / in-group ops)
    are stand-ins listed in ref_shim.py.
"""
