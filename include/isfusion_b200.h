/*
 * isfusion_b200 — C ABI of the B200-native IS-Fusion fusion-neck hot path.
 *
 * Conventions: plain pointers + sizes, no torch types; every
 * pointer is DEVICE memory unless marked "host"; `stream` is a cudaStream_t
 * passed as void*; functions enqueue on that stream and never synchronise the
 * device; no hidden allocation — scratch memory comes from the caller through
 * (`workspace`, `ws_bytes`), sized by the matching *_workspace_size call;
 * data-dependent output counts are written to a device int the caller reads
 * when (and if) it needs them.  Return value: 0 = ISF_OK, otherwise an error
 * code; isf_last_error() gives the message of the calling thread's last error.
 * Each entry point cites the reference interface it replaces (file:line under
 * the reference checkout yinjunbo/IS-Fusion @ 86c882a).
 */
#ifndef ISFUSION_B200_H_
#define ISFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISF_OK 0
#define ISF_ERR_INVALID 1
#define ISF_ERR_WORKSPACE 2
#define ISF_ERR_CUDA 3
#define ISF_ERR_UNSUPPORTED 4

#define ISF_REDUCE_SUM 0
#define ISF_REDUCE_MEAN 1
#define ISF_REDUCE_MAX 2

/* ---- library ---------------------------------------------------------------- */
int isf_version(void);
const char* isf_last_error(void);
/* number of kernels this library has launched since the last reset (bench.py gpu_launches) */
unsigned long long isf_launch_count(int reset);
/* runtime options: "gemm_cluster" = 1 | 2 : thread-block-cluster size of the tcgen05 GEMM (CTAs of a cluster
 * share every weight chunk through TMA multicast).  "gemm_mt" = 1 | 2 | 3 : 128-row sub-tiles per CTA tile of the
 * sparse-conv GEMM (2 = one weight chunk feeds 256 output rows, 3 = only for the 32- and 256-column layers).  "gemm_v4" =
 * 0 | 1 | 2 : schedule with dedicated epilogue warps and a double-buffered TMEM accumulator (csrc/tc_gemm4.cuh; 1 = sparse-
 * conv launches, 2 = dense GEMMs too).  "gemm_ln_fast" = 0 | 1 : split-row LayerNorm epilogue (default 1).  "gemm_debug" = bit mask of TIMING experiments
 * (results meaningless; csrc/tc_gemm.cuh GemmParams::debug).  "tile_order_drop_bits" = 0..26 : low occupancy-mask
 * bits left out of the sort of isf_spconv_tile_order (fewer radix passes).  "nbr_rows" = 0 | 1 : row-per-thread
 * kernel behind isf_spconv_neighbors_bitmap (one bitmap word load per (kz, ky) instead of per tap; same results;
 * default 1). */
int isf_set_option(const char* name, int value);

/* ---- a1: dynamic voxelization -------------------------------------------------
 * replaces voxel_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3)
 *   mmdet3d/ops/voxel/src/voxelization.cpp:8, voxelization.h:83-95,
 *   kernel voxelization_cuda.cu:24-61 (launch :485-528), CPU voxelization_cpu.cpp:7-43.
 * points [n,nfeat] fp32 (xyz first), coors [n,3] int32 (z,y,x).  voxel_size[3] and
 * coors_range[6] are HOST arrays.  cpu_semantics=0 reproduces the CUDA kernel's partial
 * -1 rows ((-1,0,0) / (-1,-1,0) / (-1,-1,-1)), 1 writes (-1,-1,-1) like the CPU path. */
int isf_dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                         const float* coors_range, int* coors, int cpu_semantics, void* stream);

/* the same op in two calls: `group` (sort + run-length encode + voxel coordinates + count) then `compute` (the
 * fused feature kernel on the grouping left in the same workspace, same stream, same n/dims).  The coordinates
 * - all the sparse-conv rulebooks need - are thus available before the feature kernel has run. */
int isf_dynamic_vfe_group(const int* coors, int n, const int* dims, int* voxel_coors, int* num_voxels_dev,
                          void* workspace, size_t ws_bytes, void* stream);
int isf_dynamic_vfe_compute(const float* points, int n, int nfeat, const int* dims, const float* voxel_size,
                            const float* offsets, const float* w1, const float* scale1, const float* shift1,
                            const float* w2, const float* scale2, const float* shift2, int in_ch, int feat_ch,
                            float* voxel_feats_f32, void* voxel_feats_f16, void* workspace, size_t ws_bytes,
                            void* stream);

/* tensor-core variant of `compute` (the default of the module): the two 64x64 halves of layer 2 run on the tcgen05
 * GEMM (w2a = W2[:, :64], w2b = W2[:, 64:], fp16 K-major [64][64]); fp16 output only.  Needs
 * isf_dynamic_vfe_tc_workspace_size(n) bytes for both the preceding `group` call and this call. */
int isf_dynamic_vfe_tc_workspace_size(int n, size_t* bytes);
int isf_dynamic_vfe_compute_tc(const float* points, int n, int nfeat, const int* dims, const float* voxel_size,
                               const float* offsets, const float* w1, const float* scale1, const float* shift1,
                               const void* w2a, const void* w2b, const float* scale2, const float* shift2,
                               void* voxel_feats_f16, void* workspace, size_t ws_bytes, void* stream);

/* ---- a2: dynamic scatter (per-voxel mean / max / sum) -------------------------
 * replaces voxel_layer.dynamic_point_to_voxel_forward(feats, coors, reduce_type)
 *   mmdet3d/ops/voxel/src/voxelization.h:108-120, scatter_points_cuda.cu:183-239.
 * feats [n,c] fp32, coors [n,ndim] int32 (ndim 3 or 4).  dims (HOST, may be NULL) = exclusive
 * upper bound of every coordinate column; NULL assumes < 2^21 (ndim 3) / < 2^15 (ndim 4).
 * Outputs are caller-allocated with capacity n rows: reduced [n,c], out_coors [n,ndim]
 * (sorted-unique order, as at::unique_dim), coors_map [n] (-1 for rows with a negative
 * coordinate), reduce_count [n]; *num_voxels_dev receives the number of valid rows. */
int isf_dynamic_scatter_workspace_size(int n, size_t* bytes);
int isf_dynamic_point_to_voxel_forward(const float* feats, const int* coors, int n, int c, int ndim,
                                       const int* dims, int reduce_type, float* reduced, int* out_coors,
                                       int* coors_map, int* reduce_count, int* num_voxels_dev,
                                       void* workspace, size_t ws_bytes, void* stream);

/* ---- K3b: dynamic scatter backward ----------------------------------------------
 * replaces voxel_layer.dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats,
 *   coors_map, reduce_count, reduce_type)   mmdet3d/ops/voxel/src/voxelization.h:122-140,
 *   scatter_points_cuda.cu:105-179 (kernels), :241-320 (driver).
 * grad_feats [n,c] fp32 is written completely (no pre-zeroing needed): sum -> grad_reduced[coors_map[x]];
 * mean -> that / reduce_count; max -> the gradient of (voxel, channel) goes to the LOWEST-index point whose feature
 * equals the reduced value, 0 elsewhere; rows with coors_map -1 get 0.  reduce_type as in the forward (0 sum, 1 mean,
 * 2 max).  feats / reduced are only read for max, reduce_count only for mean.  workspace: num_reduced*c int32 (max). */
int isf_dynamic_scatter_backward_workspace_size(int num_reduced, int c, size_t* bytes);
int isf_dynamic_point_to_voxel_backward(float* grad_feats, const float* grad_reduced, const float* feats,
                                        const float* reduced, const int* coors_map, const int* reduce_count, int n,
                                        int num_reduced, int c, int reduce_type, void* workspace, size_t ws_bytes,
                                        void* stream);

/* ---- input contract: multi-sweep merge --------------------------
 * replaces the per-sweep numpy work of LoadPointsFromMultiSweeps.__call__
 *   (mmdet3d/datasets/pipelines/loading.py:846-906; _remove_close :824-844).
 * raw [m, load_dim] fp32 on the DEVICE (load_dim >= 5) = key frame followed by the sweeps, copied in one piece;
 * the segment tables are HOST arrays (num_segments <= 24): seg_start [S+1] row offsets (seg_start[0] = 0,
 * seg_start[S] = m), rot [S,3,3] / trans [S,3] = sensor2lidar_rotation / _translation (float64, read only where
 * seg_flags bit 0 is set), lag [S] = value of column 4 (ts - sweep_ts, 0 for the key frame), seg_flags [S]: bit 0 =
 * rotate + translate, bit 1 = drop points with |x| < radius and |y| < radius (tested on the raw coordinates).
 * range (HOST [6] or NULL): PointsRangeFilter fused (transforms_3d.py:2002-2037, BasePoints.in_range_3d): only points
 * strictly inside (x0,y0,z0,x1,y1,z1) after the transform are kept.
 * out [m, 5] fp32 (capacity m rows, input order kept), *count_dev = rows written.  float64 arithmetic rounded to
 * float32 after the rotation and after the translation, as numpy does on the reference's float32 arrays. */
int isf_merge_sweeps_workspace_size(int m, size_t* bytes);
int isf_merge_sweeps(const float* raw, int m, int load_dim, int num_segments, const int* seg_start, const double* rot,
                     const double* trans, const double* lag, const int* seg_flags, float radius, const float* range,
                     float* out, int* count_dev, void* workspace, size_t ws_bytes, void* stream);
/* TEST HOOK: the kernels' own per-point source (shared host/device) on HOST memory; *count = rows written. */
int isf_debug_merge_sweeps_host(const float* raw, int m, int load_dim, int num_segments, const int* seg_start,
                                const double* rot, const double* trans, const double* lag, const int* seg_flags,
                                float radius, const float* range, float* out, int* count);

/* ---- input contract, camera side: ImageAug3D + ImageNormalize -----------
 * replaces the PIL / torchvision pixel work of mmdet3d/datasets/pipelines/transforms_3d.py:25-43 (ImageNormalize:
 *   ToTensor + Normalize) and :82-90 (ImageAug3D.img_transform: img.resize(resize_dims) = Pillow's 8-bit BICUBIC
 *   resampler, img.crop(crop), FLIP_LEFT_RIGHT; rotate = 0).
 * isf_image_plan: Pillow's coefficient tables (Resample.c precompute_coeffs + normalize_coeffs_8bpc) for the output
 *   columns / rows that survive the crop, written to plan_dev (isf_image_plan_size bytes); rows[2] (HOST) receives the
 *   input rows [lo, hi) the vertical pass reads.  crop = (crop_x, crop_y, crop_x + out_w, crop_y + out_h) in the
 *   resized image; parts outside it read zeros (PIL crop).
 * isf_image_aug_normalize: img [n, in_h, in_w, 3] uint8 -> out [n, 3, out_h, out_w] fp32 | fp16 =
 *   ((u8 / 255) - mean) / std of the resampled + cropped (+ flipped) frame; n frames share the plan.  mean / std HOST [3].
 *   workspace: isf_image_aug_workspace_size (the uint8 intermediate between the two passes). */
int isf_image_plan_size(int in_w, int in_h, int rs_w, int rs_h, int out_w, int out_h, size_t* plan_bytes);
int isf_image_plan(int in_w, int in_h, int rs_w, int rs_h, int crop_x, int crop_y, int out_w, int out_h, int flip,
                   void* plan_dev, size_t plan_bytes, int* rows, void* stream);
int isf_image_aug_workspace_size(int n, int out_w, int row_lo, int row_hi, size_t* bytes);
int isf_image_aug_normalize(const unsigned char* img, int n, int in_w, int in_h, int rs_w, int rs_h, int out_w, int out_h,
                            int row_lo, int row_hi, const void* plan_dev, const float* mean, const float* stdv, void* out,
                            int out_is_half, void* workspace, size_t ws_bytes, void* stream);
/* TEST HOOK: plan + both passes with the kernels' own per-pixel source on HOST memory; out fp32 [n, 3, out_h, out_w]. */
int isf_debug_image_aug_host(const unsigned char* img, int n, int in_w, int in_h, int rs_w, int rs_h, int crop_x, int crop_y,
                             int out_w, int out_h, int flip, const float* mean, const float* stdv, float* out);

/* ---- K7: rank inside a group --------------------------------------------------
 * replaces ingroup_indices.forward(group_inds, out_inds)
 *   mmdet3d/ops/TorchEx/torchex/src/ingroup_inds/ingroup_inds.cpp:28-54, kernel :17-30.
 * group_inds [n] int64 in [0, 2^key_bits); out_inds[i] = stable rank of i inside its group
 * (the reference hands ranks out in atomic arrival order; any bijection is valid). Uses the
 * workspace size of isf_dynamic_scatter_workspace_size(n). */
int isf_ingroup_indices(const long long* group_inds, int n, int key_bits, long long* out_inds,
                        void* workspace, size_t ws_bytes, void* stream);

/* ---- a5: deterministic hard voxelization (pillars) ----------------------------
 * replaces voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
 *   coors_range, max_points, max_voxels, NDim, deterministic=True) -> voxel_num
 *   mmdet3d/ops/voxel/src/voxelization.h:58-81, voxelization_cuda.cu:231-373,
 *   voxelization_cpu.cpp:45-144.
 * voxels [max_voxels,max_points,nfeat] fp32, coors [max_voxels,3], num_points_per_voxel
 * [max_voxels]: rows < voxel_num are fully written (zero padded), other rows untouched. */
int isf_hard_voxelize_workspace_size(int n, size_t* bytes);
int isf_hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                      const float* coors_range, int max_points, int max_voxels, float* voxels, int* coors,
                      int* num_points_per_voxel, int* voxel_num_dev, void* workspace, size_t ws_bytes,
                      void* stream);

/* ---- a3: fused DynamicVFE ------------------------------------------------------
 * replaces DynamicVFE.forward (mmdet3d/models/voxel_encoders/voxel_encoder.py:452-547) with
 * its three DynamicScatter calls and two map_voxel_center_to_point canvases (:410-450), eval
 * mode (BN folded into scale/shift by the caller).  points [n,nfeat>=5], coors [n,4] (b,z,y,x)
 * all in range, dims[4] HOST = (B,Z,Y,X).  voxel_size[3], offsets[3] (= v/2 + range_min) HOST.
 * w1 [64,11], w2 [64,128] fp32 row-major (nn.Linear layout).  Outputs (capacity n rows):
 * voxel_feats_f32 and/or voxel_feats_f16 [n,64] (either may be NULL), voxel_coors [n,4]. */
int isf_dynamic_vfe_workspace_size(int n, size_t* bytes);
int isf_dynamic_vfe_forward(const float* points, const int* coors, int n, int nfeat, const int* dims,
                            const float* voxel_size, const float* offsets, const float* w1,
                            const float* scale1, const float* shift1, const float* w2, const float* scale2,
                            const float* shift2, int in_ch, int feat_ch, float* voxel_feats_f32,
                            void* voxel_feats_f16, int* voxel_coors, int* num_voxels_dev, void* workspace,
                            size_t ws_bytes, void* stream);

/* ---- tensor-core GEMM with fused epilogue (tcgen05, fp16 in / fp32 accumulate) ----
 * out[m, n] = act( LN( (A[m,k] . W[n,kpad]^T) * scale[n] + shift[n] + residual[m,n] ) )
 * Replaces the cuBLAS calls behind nn.Linear / nn.MultiheadAttention projections / FFN on the path
 * (mmdet3d/models/sst/sst_basic_block_v2.py:105-126, backbones/sst_v2.py:78-79,
 *  middle_encoders/fusion_encoder.py:576,599,648-651).  A [m,k] fp16 row-major (k % 8 == 0);
 * W [n,kpad] fp16 K-major, zero padded to kpad % 64 == 0; n % 32 == 0.  scale/shift/residual/
 * ln_gamma/ln_beta may be NULL; act: 0 none, 1 ReLU, 2 GELU(erf).  LayerNorm needs n in
 * {32,64,128,256}.  out_rows (optional) scatters output row i to row out_rows[i] (-1 = drop).
 * m_dev (optional) overrides m with a device-side row count (m is then the upper bound). */
int isf_gemm_f16(const void* A, const void* W, int m, int k, int kpad, int n, const float* scale,
                 const float* shift, const void* residual, int ldr, int act, const float* ln_gamma,
                 const float* ln_beta, float ln_eps, void* out_f16, float* out_f32, int ldo,
                 const int* out_rows, const int* m_dev, void* stream);

/* ---- a7 / a10 / a13 / a14: dense 3x3 BEV convolution (tcgen05 implicit GEMM, TMA-tiled halos) ------------
 * replaces the nn.Conv2d(3x3, padding 1) + BN (+ReLU) modules of the neck, which the reference runs on cuDNN:
 *   ISFusionEncoder.conv_fusion mmdet3d/models/middle_encoders/fusion_encoder.py:862-869, conv_ins / conv_scene /
 *   conv_heatmap / heatmap_head_{1,2,3} :915-960 (call sites :1093-1146), SECONDV2 blocks + ds_layer
 *   mmdet3d/models/backbones/second.py:127-167 (forward :200-239).
 * x [batch,h,w,ldx] fp16 channels-last (cin <= ldx, cin % 64 == 0); w_packed [cout_pad, 9*cin] fp16, column
 * (ky*3 + kx)*cin + c (BN folded), cout_pad in {32,64,128,256}; shift [cout_pad] fp32 (folded BN / bias) or NULL;
 * stride 1 | 2 (stride 2: h, w even); out [batch,oh,ow,ldo] fp16, channels [0, cout_store) are written
 * (cout_store % 8 == 0; the padded channels of w_packed are zero).  No workspace. */
int isf_conv3x3_nhwc_f16(const void* x, int batch, int h, int w, int cin, int ldx, const void* w_packed, int cout_pad,
                         const float* shift, int stride, int relu, void* out, int ldo, int cout_store, void* stream);

/* ---- a4: sparse 3-D convolution ------------------------------------------------
 * Replaces the spconv / mmcv.ops sparse-conv stack the reference delegates to
 * (call sites mmdet3d/models/middle_encoders/sparse_encoder.py:122-136, mmdet3d/ops/sparse_block.py:117-134;
 *  semantic spec mmdet3d/ops/bevfusion-ops/spconv/{ops.py:45-125, conv.py:114-223,
 *  include/spconv/spconv_ops.h:27-141,260-361, include/spconv/geometry.h:24-85,144-194}).
 * Indices are [V,4] int32 (b,z,y,x); shapes / kernel geometry are HOST int[3] arrays (z,y,x order).
 *
 * isf_hash_build: coordinate -> row hash table over `indices`.  `table` = 2*capacity 32-bit words
 *   (capacity = isf_hash_capacity(v), a power of two >= 2v).
 * (every *_dev count argument is optional: when given it overrides the host count, which then is the capacity
 *  of the buffer — this is what keeps the whole encoder free of host synchronisation / graph-capturable.)
 * isf_spconv_out_indices: output sites of a strided SparseConv3d = sorted-unique linear index of every
 *   (input, kernel offset) candidate (the GPU ordering of spconv_ops.h:129-137): the valid candidates
 *   (<= prod ceil(k/s) <= 8 per input site) are compacted, radix-sorted and made unique; writes up to out_cap rows
 *   and the count (clamped to out_cap; == out_cap means "possibly overflowed") to *num_out_dev.
 * isf_spconv_neighbors: neighbour table nbr[k * nbr_stride + o] = input row reaching output row o via kernel
 *   offset k (k = (kz*KH+ky)*KW+kx, geometry.h:62-73) or -1.  subm=1: outputs = inputs, centre aligned.
 * isf_spconv_implicit_gemm: out[o, :] = act((sum_k in[nbr[k,o], :] . W[k]) * scale + shift + residual[o, :])
 *   in ONE tcgen05 kernel.  w_packed [cout, kpad] fp16, column k*cin + c = W[k][c][cout] (kpad % 64 == 0,
 *   zero padded).  cin a power of two >= 8, cout in {32,64,128,256}, taps <= 27.
 * isf_sparse_to_dense: SparseConvTensor.dense() -> [B,C,D,H,W] fp16 (structure.py:49-59).
 * isf_sparse_to_bev_nhwc: the [B, C*D, H, W] BEV view (sparse_encoder.py:133-136, channel c*D+z) written
 *   channels-last with row pitch ldo at channel offset ch_offset (caller zero-fills). */
int isf_hash_capacity(int n);
int isf_hash_build(const int* indices, int v, const int* v_dev, const int* spatial_shape, int batch, void* table,
                   int capacity, void* stream);
int isf_spconv_out_shape(const int* in_shape, const int* ksize, const int* stride, const int* padding,
                         const int* dilation, int* out_shape);
int isf_spconv_out_indices_workspace_size(int v_in, int kvol, int out_cap, size_t* bytes);
int isf_spconv_out_indices(const int* in_indices, int v_in, const int* v_in_dev, const int* in_shape, int batch, const int* ksize,
                           const int* stride, const int* padding, const int* dilation, int* out_indices,
                           int out_cap, int* num_out_dev, void* workspace, size_t ws_bytes, void* stream);
int isf_spconv_neighbors(const int* out_indices, int v_out, const int* v_out_dev, const void* in_table,
                         int in_capacity, const int* in_shape, const int* ksize, const int* stride,
                         const int* padding, const int* dilation, int subm, int* nbr, long long nbr_stride,
                         void* stream);
/* bitmap-rank rulebook for tensors whose rows are in ascending linear-key order (every tensor this library
 * produces): one bit per cell (`bits`, isf_bitmap_words(shape,batch) words) + exclusive prefix of the word popcounts
 * (`prefix`, same length) give row(coordinate) in O(1) and the sorted-unique output sites of a strided conv with no
 * sort.  isf_bitmap_build: map of an existing tensor.  isf_bitmap_out_sites: output sites of a strided conv (also
 * the map of the output tensor).  isf_spconv_neighbors_bitmap: isf_spconv_neighbors on such a map.  Workspace:
 * isf_bitmap_workspace_size(words). */
long long isf_bitmap_words(const int* spatial_shape, int batch);
int isf_bitmap_workspace_size(long long words, size_t* bytes);
int isf_bitmap_build(const int* indices, int v, const int* v_dev, const int* spatial_shape, int batch,
                     unsigned int* bits, int* prefix, void* workspace, size_t ws_bytes, void* stream);
int isf_bitmap_out_sites(const int* in_indices, int v_in, const int* v_in_dev, const int* in_shape, int batch,
                         const int* ksize, const int* stride, const int* padding, const int* dilation,
                         unsigned int* out_bits, int* out_prefix, int* out_indices, int out_cap, int* num_out_dev,
                         void* workspace, size_t ws_bytes, void* stream);
int isf_spconv_neighbors_bitmap(const int* out_indices, int v_out, const int* v_out_dev, const unsigned int* in_bits,
                                const int* in_prefix, const int* in_shape, const int* ksize, const int* stride,
                                const int* padding, const int* dilation, int subm, int* nbr, long long nbr_stride,
                                void* stream);
int isf_spconv_implicit_gemm(const void* feats_in, int cin, const void* w_packed, int kpad, int cout,
                             const int* nbr, long long nbr_stride, int taps, int v_out, const int* v_out_dev,
                             const float* scale, const float* shift, const void* residual, int ldr, int relu,
                             void* out_f16, int ldo, void* stream);
/* Tile order (opt-in, no reference counterpart: spconv's per-offset pair lists — spconv_ops.h:260-361 — never
 * multiply a missing neighbour; the output-stationary tiles here do, unless rows with the same neighbour pattern
 * share a tile).  isf_spconv_tile_order: stable sort of the rows of a neighbour table by their kernel-offset
 * occupancy mask (inside chunks of 2^chunk_log2 consecutive rows; >= 31 = one global sort) -> permuted table
 * `nbr_out` [taps, nbr_stride] + `row_perm` [nbr_stride] (tile slot -> storage row, -1 beyond the valid rows).
 * isf_spconv_implicit_gemm_ordered: isf_spconv_implicit_gemm over such a table; slot i reads residual row
 * row_perm[i] and writes output row row_perm[i], so `out` is bit-identical to the unordered call.
 * ksize = host int[3] (kD,kH,kW) with kD*kH*kW == taps <= 27.  Workspace: isf_spconv_tile_order_workspace_size(v). */
int isf_spconv_tile_order_workspace_size(int v, size_t* bytes);
int isf_spconv_tile_order_bits(const int* ksize, int taps, unsigned char* pos);   /* host-only: key bit of every tap */
int isf_spconv_tile_order(const int* nbr, long long nbr_stride, int taps, const int* ksize, int v, const int* v_dev,
                          int chunk_log2, int* nbr_out, int* row_perm, void* workspace, size_t ws_bytes, void* stream);
/* isf_spconv_neighbors_bitmap followed by isf_spconv_tile_order, fused: the occupancy masks come straight from the
 * bitmap, and the table is written in tile order (the plain table is never materialised).  Same outputs. */
int isf_spconv_neighbors_bitmap_ordered(const int* out_indices, int v_out, const int* v_out_dev, const unsigned int* in_bits,
                                        const int* in_prefix, const int* in_shape, const int* ksize, const int* stride,
                                        const int* padding, const int* dilation, int subm, int chunk_log2, int* nbr_out,
                                        long long nbr_stride, int* row_perm, void* workspace, size_t ws_bytes, void* stream);
/* host-only test hook: the row walk of the row-per-thread kernels (same source, compiled for the host) on HOST memory;
 * nbr [kvol, nbr_stride] and masks [v_out] (occupancy mask, offset k at bit isf_spconv_tile_order_bits()[k]) are host
 * arrays.  Lets the CPU suite check the device logic against the oracle rulebook. */
int isf_debug_nbr_rows_host(const int* out_indices, int v_out, const unsigned int* in_bits, const int* in_prefix,
                            const int* in_shape, const int* ksize, const int* stride, const int* padding,
                            const int* dilation, int subm, int* nbr, long long nbr_stride, unsigned int* masks);
int isf_spconv_implicit_gemm_ordered(const void* feats_in, int cin, const void* w_packed, int kpad, int cout,
                                     const int* nbr_ordered, long long nbr_stride, int taps, int v_out,
                                     const int* v_out_dev, const int* row_perm, const float* scale, const float* shift,
                                     const void* residual, int ldr, int relu, void* out_f16, int ldo, void* stream);
/* Dynamic tile scheduling (no reference counterpart): hands the NEXT isf_spconv_implicit_gemm[_ordered] call of this host thread
 * a device buffer of two ints that is zero on entry; the kernel's CTAs then draw their row tiles from it (densest tiles first)
 * instead of walking them round-robin, and leave it zeroed.  One buffer per rulebook: launches that share it must be
 * stream-ordered.  Without the call (or with NULL) the tiles are walked round-robin; results are bit-identical either way. */
int isf_spconv_tile_counter(int* counter2);
/* Training: weight gradient of a sparse convolution (replaces the filter-gradient half of spconv's indice_conv_backward,
 * bevfusion-ops/spconv/include/spconv/spconv_ops.h:364-470, one gather + GEMM per kernel offset there).
 *   dw[k][ci][co] += sum over p in [tap_offsets[k], tap_offsets[k+1]) of  x[in_rows[p]][ci] * g[out_rows[p]][co]
 * x [*, cin] and g [*, cout] fp16 row-major (16-byte aligned rows: cin, cout multiples of 8), in_rows / out_rows device int32 pair
 * lists grouped by kernel offset, tap_offsets device int32 [taps + 1], max_pairs_per_tap the host's bound on a group's size,
 * dw fp32 [taps, cin, cout], accumulated into (zero it first).  fp32 accumulation; the order of the additions is not fixed. */
int isf_spconv_wgrad(const void* x_f16, int cin, const void* g_f16, int cout, const int* in_rows, const int* out_rows,
                     const int* tap_offsets, int taps, int max_pairs_per_tap, float* dw, void* stream);
int isf_sparse_to_dense(const void* feats, const int* indices, int v, int c, int batch, int D, int H, int W,
                        void* out, void* stream);
int isf_sparse_to_bev_nhwc(const void* feats, const int* indices, int v, const int* v_dev, int c, int D, int H, int W,
                           void* out, int ldo, int ch_offset, void* stream);

/* ---- a6: camera -> BEV (Point-to-Grid) -----------------------------------------
 * replaces ISFusionEncoder.img_point_sampling + img_fv_to_bev (mmdet3d/models/middle_encoders/
 * fusion_encoder.py:965-1070): 6*B F.grid_sample launches with [1,256,M*12,1] temporaries become one kernel
 * (a warp per pillar).  pillars [M,P,nfeat] fp32 (zero-padded slots are projected too, Appendix C #1),
 * pillar_coors [M,4] int32 (b,z,y,x), img_feat_nhwc [B,ncam,fh,fw,256] fp16, proj_dev [B*ncam,20] fp32 =
 * {M(9) = lidar2img[:3,:3].inv(lidar_aug[:3,:3]), t(3), img_aug[:2,:3](6), img_aug[:2,3](2)}.
 * Writes 256 channels at ch_off of the channels-last row (pitch ldo) of every non-empty pillar cell.
 * pillar_coors is [M,4] (b,z,y,x) (coors_dim 4) or the per-sample [M,3] (z,y,x) with batch_index given;
 * num_pillars_dev (optional) = device-side pillar count (num_pillars is then the buffer capacity). */
int isf_cam2bev(const float* pillars, int num_pillars, const int* num_pillars_dev, int points_per_pillar, int nfeat,
                const int* pillar_coors, int coors_dim, int batch_index, const void* img_feat_nhwc,
                const float* proj_dev, int batch, int ncam, int fh, int fw, int channels,
                float img_w, float img_h, int bev_h, int bev_w, void* out_nhwc, int ldo, int ch_off, void* stream);

/* ---- a8/a9: SST window attention core ------------------------------------------
 * replaces get_window_coors / get_inner_win_inds / get_flat2win_inds_v2 / flat2window / nn.MultiheadAttention
 * core / window2flat (mmdet3d/ops/sst/sst_ops.py:22-134,219-287; sst/sst_basic_block_v2.py:41-76) on a dense
 * grid x grid BEV: qkv [B*grid*grid, 3C] fp16 (in_proj incl. bias), pos_q/pos_k [36,C] fp32 (Wq.pos, Wk.pos),
 * shifted = 0 (aligned 6x6 windows) | 1 (shifted by 3).  out [B*grid*grid, C] fp16 (before out_proj). */
int isf_window_attention(const void* qkv, const float* pos_q, const float* pos_k, int batch, int grid, int channels,
                         int heads, int shifted, void* out, void* stream);

/* ---- image backbone: Swin shifted-window attention ----------------------------
 * replaces the pad / roll / window_partition / WindowMSA core / window_reverse / roll / crop sequence of
 * mmdet3d/models/backbones/swin.py:85-123,177-283.  qkv [B*H*W, 3C] fp16 token-major (bias included), qkv_bias [3C] fp32
 * (value of a padding position), rel_bias [heads, 49, 49] fp32 -> out [B*H*W, C] fp16 (input of `proj`).  7x7 windows,
 * head_dim 32, shift 0 | 3.  isf_debug_swin_window_map: host-only, the kernel's (window, slot) -> token / region closed form. */
int isf_swin_window_attention(const void* qkv, const float* qkv_bias, const float* rel_bias, int batch, int H, int W, int channels,
                              int heads, int shift, void* out, void* stream);
int isf_debug_swin_window_map(int H, int W, int shift, int* tok, int* region);

/* ---- a10: heat-map NMS + top-K --------------------------------------------------
 * replaces sigmoid / max_pool2d / argsort of fusion_encoder.py:1100-1131.  logits [B,classes,H,W] fp32;
 * classes >= plain_from keep every cell (kernel-size-1 classes); top_idx [B,k] int32 = flat cell index. */
int isf_heatmap_topk_workspace_size(int batch, int classes, int H, int W, size_t* bytes);
int isf_heatmap_topk(const float* logits, int batch, int classes, int H, int W, int plain_from, int k, int* top_idx,
                     float* top_score, void* workspace, size_t ws_bytes, void* stream);
/* same, top_idx = class * H*W + cell: the proposals of TransFusionHeadV2.forward_single
 * (mmdet3d/models/dense_heads/transfusion_head_v2.py:789-826), which also needs each proposal's class. */
int isf_heatmap_topk_class(const float* logits, int batch, int classes, int H, int W, int plain_from, int k, int* top_idx,
                           float* top_score, void* workspace, size_t ws_bytes, void* stream);

/* ---- a11: multi-scale deformable attention forward ------------------------------
 * replaces mmcv._ext.ms_deform_attn_forward (call site mmdet3d/models/middle_encoders/
 * multi_scale_deformable_attn_function.py:118; kernel spec ops/src/cuda/ms_deform_im2col_cuda.cuh:34-83,238-298).
 * value [B,S,H,D] fp32 or fp16, spatial_shapes int64 [L,2], level_start_index int64 [L] (device),
 * sampling_loc [B,Q,H,L,P,2] fp32, attn_weight [B,Q,H,L,P] fp32 -> out [B,Q,H*D] fp32. */
int isf_ms_deform_attn_forward(const void* value, int value_is_half, const long long* spatial_shapes,
                               const long long* level_start_index, const float* sampling_loc,
                               const float* attn_weight, int B, int S, int H, int D, int L, int Q, int P, float* out,
                               void* stream);

/* ---- a11: the whole instance-context decoder in one launch (csrc/decoder.cu) ---------------------------
 * replaces InsContextAtt.forward (mmdet3d/models/middle_encoders/fusion_encoder.py:768-830) = query position embedding
 * (PositionEmbeddingLearned :176-190) + `layers` x DeformableTransformerDecoderLayer.forward (:600-674: self-attention,
 * MSDeformAttn :523-598 with ms_deform_attn_core ops/functions/ms_deform_attn_func.py:41-61, FFN, three LayerNorms), and
 * the key / value in-projection of Instane2SceneAtt's MultiheadAttention (:441-462) applied to its output.
 * scene_tokens [B, G*G, C] fp16 (conv_scene output, [Y,X] frame); top_idx [B, K] int32 cell n = y*G + x; value [B*G*G, ldv]
 * fp16 = value_proj(tokens + key_pos) of layer l at columns [l*C, (l+1)*C); out [B, K, C] fp32; kv_out [B, K, 2C] fp16 or
 * NULL.  C = 128, 8 heads, 16 points, K <= 224.  packed_weights (fp16) / params (fp32): layouts of
 * isfusion_b200.fusion_encoder.pack_decoder (sizes from isf_ins_context_decoder_sizes); every weight matrix [N,128] is
 * stored per 8 output rows in mma.sync B-fragment order. */
int isf_ins_context_decoder_sizes(int layers, size_t* packed_weight_bytes, size_t* param_floats);
int isf_ins_context_decoder(const void* scene_tokens, const int* top_idx, const void* value, int ldv,
                            const void* packed_weights, const float* params, int batch, int grid, int num_queries,
                            int channels, int heads, int points, int layers, float* out, void* kv_out, void* stream);

/* ---- K9b: multi-scale deformable attention backward -------------------------------
 * replaces mmcv._ext.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
 *   grad_output, grad_value, grad_sampling_loc, grad_attn_weight, im2col_step)
 *   (call site multi_scale_deformable_attn_function.py:61-84; kernel spec ops/src/cuda/ms_deform_im2col_cuda.cuh:87-160).
 * grad_output [B,Q,H*D] fp32.  grad_value [B,S,H,D] fp32 is ACCUMULATED INTO (the caller zeroes it, as the reference's
 * autograd function does); grad_sampling_loc [B,Q,H,L,P,2] and grad_attn_weight [B,Q,H,L,P] are overwritten. */
int isf_ms_deform_attn_backward(const void* value, int value_is_half, const long long* spatial_shapes,
                                const long long* level_start_index, const float* sampling_loc, const float* attn_weight,
                                const float* grad_output, int B, int S, int H, int D, int L, int Q, int P,
                                float* grad_value, float* grad_sampling_loc, float* grad_attn_weight, void* stream);
/* TEST HOOK: the kernel's own per-sample arithmetic (shared host/device source) on HOST memory, fp32 value. */
int isf_debug_msda_backward_host(const float* value, const long long* spatial_shapes, const long long* level_start_index,
                                 const float* sampling_loc, const float* attn_weight, const float* grad_output, int B,
                                 int S, int H, int D, int L, int Q, int P, float* grad_value, float* grad_sampling_loc,
                                 float* grad_attn_weight);

/* ---- a12: instance -> scene attention -------------------------------------------
 * isf_i2s_attention: softmax(q k^T / sqrt(d)) v of fusion_encoder.py:441-462 for tokens x num_keys per frame:
 *   q [B*tokens, C] fp16, kv [B, num_keys, 2C] fp16 (k | v), out [B*tokens, C] fp16.
 * isf_channel_attention: out = S + softmax(S I^T, -1) I per (frame, channel) over the grid x grid plane
 *   (fusion_encoder.py:495-500, no 1/sqrt(d)); scene / ins / out channels-last fp16 with pitches lds/ldi/ldo. */
int isf_i2s_attention(const void* q, const void* kv, int batch, int tokens, int num_keys, int channels, int heads,
                      void* out, void* stream);
/* query -> scene attention (TransFusionHeadV2 cross attention, transfusion_head_v2.py:107-110 through
 * multi_head_attention_forward :388-470): few queries x many keys, split over key blocks (flash-decoding), the tile code of
 * isf_i2s_attention with the roles swapped.  q [B, queries, C] fp16 (projected, un-scaled), kv [B, num_keys, 2C] fp16 (k | v)
 * -> out [B, queries, C] fp16.  C = 128, 8 heads. */
int isf_query_scene_attention_workspace_size(int batch, int queries, int num_keys, int channels, int heads, size_t* bytes);
int isf_query_scene_attention(const void* q, const void* kv, int batch, int queries, int num_keys, int channels, int heads,
                              void* out, void* workspace, size_t ws_bytes, void* stream);
int isf_channel_attention_workspace_size(int batch, int grid, int channels, size_t* bytes);
int isf_channel_attention(const void* scene, int lds, const void* ins, int ldi, int batch, int grid, int channels,
                          void* out, int ldo, void* workspace, size_t ws_bytes, void* stream);

/* bf16 twins of the dense tensor-core entry points of the Hierarchical-Scene-Fusion path (BASELINE configs[2] names a bf16
 * tensor-core path; reference: the same modules as their f16 siblings, run under torch autocast(bfloat16) semantics): identical
 * arguments, every 16-bit tensor (operands, residual, output) is bf16 instead of fp16, accumulation / epilogue / softmax are fp32. */
int isf_gemm_bf16(const void* A, const void* W, int m, int k, int kpad, int n, const float* scale, const float* shift,
                  const void* residual, int ldr, int act, const float* ln_gamma, const float* ln_beta, float ln_eps,
                  void* out_bf16, float* out_f32, int ldo, const int* out_rows, const int* m_dev, void* stream);
int isf_conv3x3_nhwc_bf16(const void* x, int batch, int h, int w, int cin, int ldx, const void* w_packed, int cout_pad,
                          const float* shift, int stride, int relu, void* out, int ldo, int cout_store, void* stream);
int isf_window_attention_bf16(const void* qkv, const float* pos_q, const float* pos_k, int batch, int grid, int channels,
                              int heads, int shifted, void* out, void* stream);
int isf_cam2bev_bf16(const float* pillars, int num_pillars, const int* num_pillars_dev, int points_per_pillar, int nfeat,
                     const int* pillar_coors, int coors_dim, int batch_index, const void* img_feat_nhwc, const float* proj_dev,
                     int batch, int ncam, int fh, int fw, int channels, float img_w, float img_h, int bev_h, int bev_w,
                     void* out_nhwc, int ldo, int ch_off, void* stream);

/* bring-up probe (NOT on the product path): TMA tile::gather4 of 128 rows x 64 fp16 columns into shared memory, raw
 * shared memory (32 KiB) returned in out32k, *status_dev = 1 when the transaction completed (csrc/debug_tma.cu). */
int isf_debug_tma_gather4(const void* a, int rows, int cols, const int* idx_dev, int col0, int variant, void* out32k,
                          int* status_dev, void* stream);

/* micro-benchmark (NOT on the product path): sustained rate of 128-row x 128-B gathers into shared memory per SM, by TMA
 * tile::gather4 (modes 0-2), a plain TMA tile load (3) or LDGSTS (4); out_cycles[block] = cycles per 16 KB chunk. */
int isf_debug_gather_bench(const void* a, int rows, int cols, const int* idx_dev, int n_idx, int mode, int stages, int iters,
                           int blocks, float* out_cycles, void* stream);

/* micro-benchmark (NOT on the product path): cycles per tcgen05.mma kind::f16 instruction of shape m x n x 16 (cta_group 1 | 2;
 * a_tmem = 1: A operand from tensor memory; chains = independent accumulators the instructions alternate between; issuers =
 * warps of the CTA issuing their own streams) issued back to back on constant operands,
 * out_cycles[block]; with out_probe (256 floats, one block) it runs one m x 32 instruction on A[i][0] = i + 1, B[j][0] = 1 and
 * returns TMEM columns 0 and 1 of the 128 lanes (the accumulator layout of the shape) (csrc/debug_tma.cu). */
int isf_debug_mma_rate(int m, int n, int cta_group, int a_tmem, int chains, int issuers, int blocks, int iters,
                       float* out_cycles, float* out_probe, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISFUSION_B200_H_ */
