// Voxelization, dynamic scatter, fused DynamicVFE, ingroup indices (SURVEY §8 rows a1,a2,a3,a5,K7).
//
// Replaces mmdet3d/ops/voxel/src/voxelization_cuda.cu (dynamic_voxelize :24-61,485-528;
// deterministic hard_voxelize :231-373) and scatter_points_cuda.cu (:183-239) of the reference,
// plus the DynamicVFE python graph (voxel_encoders/voxel_encoder.py:452-547) as one fused op.
//
// Design: every op linearises the voxel coordinate into a radix key, does ONE stable CUB radix
// sort of (key, point index) restricted to the bits in use, run-length-encodes the sorted keys
// and then works on contiguous segments: no atomics, no dense canvases, no host syncs, results
// deterministic and in the reference's sorted-unique (b,z,y,x) order.
#include <cub/cub.cuh>
#include "tc_gemm.cuh"

extern "C" void isf_count_launch(int n);

// ------------------------------------------------------------------------------------------------
// a1: dynamic voxelization
// ------------------------------------------------------------------------------------------------
struct VoxGeom {
  float vx, vy, vz;
  float xmin, ymin, zmin;
  int gx, gy, gz;
};

// fp32 subtract -> fp32 divide -> floor with no contraction / reciprocal, to be bit-identical to
// voxelization_cuda.cu:37,43,50 and voxelization_cpu.cpp:24.
__device__ __forceinline__ int vox_coord(float p, float lo, float v) {
  return (int)floorf(__fdiv_rn(__fsub_rn(p, lo), v));
}

template <bool kCpuSemantics>
__global__ void __launch_bounds__(256) dynamic_voxelize_kernel(const float* __restrict__ points, int n,
                                                               int nfeat, VoxGeom g,
                                                               int* __restrict__ coors) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = points + (size_t)i * nfeat;
  int cx = vox_coord(p[0], g.xmin, g.vx);
  int cy = vox_coord(p[1], g.ymin, g.vy);
  int cz = vox_coord(p[2], g.zmin, g.vz);
  bool bx = cx < 0 || cx >= g.gx, by = cy < 0 || cy >= g.gy, bz = cz < 0 || cz >= g.gz;
  int o0, o1, o2;
  if (kCpuSemantics) {
    if (bx || by || bz) { o0 = o1 = o2 = -1; } else { o0 = cz; o1 = cy; o2 = cx; }
  } else {
    // reference CUDA kernel: slots after the first failing axis keep the caller's zeros
    if (bx) { o0 = -1; o1 = 0; o2 = 0; }
    else if (by) { o0 = -1; o1 = -1; o2 = 0; }
    else if (bz) { o0 = -1; o1 = -1; o2 = -1; }
    else { o0 = cz; o1 = cy; o2 = cx; }
  }
  int* c = coors + (size_t)i * 3;
  c[0] = o0; c[1] = o1; c[2] = o2;
}

static int make_geom(const float* vs, const float* cr, VoxGeom* g) {
  g->vx = vs[0]; g->vy = vs[1]; g->vz = vs[2];
  g->xmin = cr[0]; g->ymin = cr[1]; g->zmin = cr[2];
  // grid = round((max-min)/v) in fp32 (voxelization_cuda.cu:508-510)
  g->gx = (int)roundf((cr[3] - cr[0]) / vs[0]);
  g->gy = (int)roundf((cr[4] - cr[1]) / vs[1]);
  g->gz = (int)roundf((cr[5] - cr[2]) / vs[2]);
  return 0;
}

extern "C" int isf_dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                                    const float* coors_range, int* coors, int cpu_semantics,
                                    void* stream) {
  ISF_CHECK_ARG(n >= 0 && nfeat >= 3 && voxel_size && coors_range);
  if (n == 0) return ISF_OK;
  ISF_CHECK_ARG(points && coors);
  VoxGeom g;
  make_geom(voxel_size, coors_range, &g);
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = isf_cdiv(n, 256);
  if (cpu_semantics)
    dynamic_voxelize_kernel<true><<<blocks, 256, 0, st>>>(points, n, nfeat, g, coors);
  else
    dynamic_voxelize_kernel<false><<<blocks, 256, 0, st>>>(points, n, nfeat, g, coors);
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// key packing
// ------------------------------------------------------------------------------------------------
struct KeyLayout {
  int ndim;
  int bits[4];   // bits per dimension (most significant dimension first)
  int total_bits;
};

static int ceil_log2(long long v) {
  int b = 0;
  while (b < 62 && (1ll << b) < v) ++b;
  return b < 1 ? 1 : b;
}

static KeyLayout make_layout(int ndim, const int* dims) {
  KeyLayout L;
  L.ndim = ndim;
  L.total_bits = 0;
  for (int d = 0; d < 4; ++d) L.bits[d] = 0;
  for (int d = 0; d < ndim; ++d) {
    L.bits[d] = dims ? ceil_log2(dims[d]) : (ndim == 3 ? 21 : 15);
    L.total_bits += L.bits[d];
  }
  return L;
}

__device__ __forceinline__ unsigned long long pack_key(const int* c, const KeyLayout& L) {
  unsigned long long k = 0;
  bool bad = false;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    if (d < L.ndim) {
      int v = c[d];
      bad |= (v < 0) || (v >= (1 << L.bits[d]));
      k = (k << L.bits[d]) | (unsigned long long)(unsigned)v;
    }
  }
  return bad ? ~0ull : k;
}

__device__ __forceinline__ void unpack_key(unsigned long long k, const KeyLayout& L, int* c) {
#pragma unroll
  for (int d = 3; d >= 0; --d) {
    if (d < L.ndim) {
      c[d] = (int)(k & ((1ull << L.bits[d]) - 1));
      k >>= L.bits[d];
    }
  }
}

__global__ void __launch_bounds__(256) make_keys_kernel(const int* __restrict__ coors, int n, KeyLayout L,
                                                        unsigned long long* __restrict__ keys,
                                                        int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[4] = {0, 0, 0, 0};
  for (int d = 0; d < L.ndim; ++d) c[d] = coors[(size_t)i * L.ndim + d];
  keys[i] = pack_key(c, L);
  idx[i] = i;
}

// Shared "group points by key" front end.  After it runs:
//   idx_sorted[j]   point index at sorted position j (stable: index order inside a voxel)
//   uniq[r], cnt[r], start[r]  key / size / first sorted position of run r,   *num_runs runs,
//   *num_valid = number of runs whose key is valid (the invalid run, if any, is the last one).
struct GroupBuffers {
  unsigned long long *keys, *keys_sorted, *uniq;
  int *idx, *idx_sorted, *cnt, *start, *num_runs, *num_valid;
  void* cub_tmp;
  size_t cub_bytes;
};

static size_t group_cub_bytes(int n) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, n);
  cub::DeviceRunLengthEncode::Encode(nullptr, b, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                     (int*)nullptr, (int*)nullptr, n);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (int*)nullptr, (int*)nullptr, n);
  size_t m = a > b ? a : b;
  return m > c ? m : c;
}

static void group_carve(IsfArena& A, int n, GroupBuffers* G) {
  G->keys = A.take<unsigned long long>(n);
  G->keys_sorted = A.take<unsigned long long>(n);
  G->uniq = A.take<unsigned long long>(n + 1);
  G->idx = A.take<int>(n);
  G->idx_sorted = A.take<int>(n);
  G->cnt = A.take<int>(n + 1);
  G->start = A.take<int>(n + 1);
  G->num_runs = A.take<int>(1);
  G->num_valid = A.take<int>(1);
  G->cub_bytes = group_cub_bytes(n);
  G->cub_tmp = A.take<char>(G->cub_bytes);
}

__global__ void finalize_runs_kernel(const unsigned long long* __restrict__ uniq, const int* num_runs,
                                     int* num_valid, int* out_count) {
  int r = *num_runs;
  int v = (r > 0 && uniq[r - 1] == ~0ull) ? r - 1 : r;
  *num_valid = v;
  if (out_count) *out_count = v;
}

static int group_by_key(GroupBuffers& G, int n, int key_bits, int* out_count_dev, cudaStream_t st) {
  size_t tmp = G.cub_bytes;
  // +1 bit so that the all-ones invalid key still sorts last
  int end_bit = key_bits + 1 > 64 ? 64 : key_bits + 1;
  // invalid key is ~0: restricting the sort to [0,end_bit) keeps it last because all its bits are 1.
  ISF_CUDA(cub::DeviceRadixSort::SortPairs(G.cub_tmp, tmp, G.keys, G.keys_sorted, G.idx, G.idx_sorted, n, 0,
                                           end_bit, st));
  tmp = G.cub_bytes;
  ISF_CUDA(cub::DeviceRunLengthEncode::Encode(G.cub_tmp, tmp, G.keys_sorted, G.uniq, G.cnt, G.num_runs, n, st));
  tmp = G.cub_bytes;
  // cnt beyond num_runs is garbage; scan all n entries (cheap) - only the first num_runs are used.
  ISF_CUDA(cub::DeviceScan::ExclusiveSum(G.cub_tmp, tmp, G.cnt, G.start, n, st));
  finalize_runs_kernel<<<1, 1, 0, st>>>(G.uniq, G.num_runs, G.num_valid, out_count_dev);
  isf_count_launch(8);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// a2: dynamic_point_to_voxel_forward
// ------------------------------------------------------------------------------------------------
enum { ISF_REDUCE_SUM = 0, ISF_REDUCE_MEAN = 1, ISF_REDUCE_MAX = 2 };

__global__ void __launch_bounds__(256) scatter_reduce_kernel(
    const float* __restrict__ feats, int c, int reduce_type, const int* __restrict__ idx_sorted,
    const unsigned long long* __restrict__ uniq, const int* __restrict__ cnt, const int* __restrict__ start,
    const int* __restrict__ num_valid, KeyLayout L, int n, float* __restrict__ reduced,
    int* __restrict__ out_coors, int* __restrict__ reduce_count) {
  const int nv = *num_valid;
  const long long total = (long long)nv * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int v = (int)(t / c), ch = (int)(t % c);
    int s = start[v], m = cnt[v];
    float acc = (reduce_type == ISF_REDUCE_MAX) ? -INFINITY : 0.f;
    for (int j = 0; j < m; ++j) {
      float x = feats[(size_t)idx_sorted[s + j] * c + ch];
      acc = (reduce_type == ISF_REDUCE_MAX) ? fmaxf(acc, x) : acc + x;
    }
    if (reduce_type == ISF_REDUCE_MEAN) acc = acc / (float)m;
    reduced[(size_t)v * c + ch] = acc;
    if (ch == 0) {
      int cc[4];
      unpack_key(uniq[v], L, cc);
      for (int d = 0; d < L.ndim; ++d) out_coors[(size_t)v * L.ndim + d] = cc[d];
      reduce_count[v] = m;
    }
  }
}

// coors_map[point] = voxel row (or -1).  One thread per sorted position; voxel id by binary search
// over the run starts (runs are short, V ~ N/2, so this is ~17 steps of L2-resident reads).
__global__ void __launch_bounds__(256) coors_map_kernel(const int* __restrict__ idx_sorted,
                                                        const int* __restrict__ start,
                                                        const int* __restrict__ num_runs,
                                                        const int* __restrict__ num_valid, int n,
                                                        int* __restrict__ coors_map) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int lo = 0, hi = *num_runs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (start[mid] <= j) lo = mid; else hi = mid - 1;
  }
  coors_map[idx_sorted[j]] = (lo < *num_valid) ? lo : -1;
}

extern "C" int isf_dynamic_scatter_workspace_size(int n, size_t* bytes) {
  ISF_CHECK_ARG(n >= 0 && bytes);
  IsfArena A(nullptr, 0);
  GroupBuffers G;
  group_carve(A, n > 0 ? n : 1, &G);
  *bytes = A.off + 1024;
  return ISF_OK;
}

extern "C" int isf_dynamic_point_to_voxel_forward(const float* feats, const int* coors, int n, int c, int ndim,
                                                  const int* dims, int reduce_type, float* reduced,
                                                  int* out_coors, int* coors_map, int* reduce_count,
                                                  int* num_voxels_dev, void* workspace, size_t ws_bytes,
                                                  void* stream) {
  ISF_CHECK_ARG(n >= 0 && c >= 1 && (ndim == 3 || ndim == 4) && reduce_type >= 0 && reduce_type <= 2);
  ISF_CHECK_ARG(num_voxels_dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ISF_CUDA(cudaMemsetAsync(num_voxels_dev, 0, sizeof(int), st));
    return ISF_OK;
  }
  ISF_CHECK_ARG(feats && coors && reduced && out_coors && coors_map && reduce_count && workspace);
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  group_carve(A, n, &G);
  if (!A.ok) { isf_set_error("dynamic scatter: workspace too small"); return ISF_ERR_WORKSPACE; }
  KeyLayout L = make_layout(ndim, dims);
  ISF_CHECK_ARG(L.total_bits <= 63);
  make_keys_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(coors, n, L, G.keys, G.idx);
  int rc = group_by_key(G, n, L.total_bits, num_voxels_dev, st);
  if (rc) return rc;
  long long total = (long long)n * c;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  scatter_reduce_kernel<<<blocks, 256, 0, st>>>(feats, c, reduce_type, G.idx_sorted, G.uniq, G.cnt, G.start,
                                                G.num_valid, L, n, reduced, out_coors, reduce_count);
  coors_map_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(G.idx_sorted, G.start, G.num_runs, G.num_valid, n,
                                                     coors_map);
  isf_count_launch(3);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// K3b: dynamic scatter backward  (scatter_points_cuda.cu:105-179, driver :241-320)
//   sum : grad_feats[x] = grad_reduced[map[x]]            mean: ... / reduce_count[map[x]]
//   max : for every (voxel, channel) the LOWEST-index point whose feature equals the reduced value receives the
//         gradient (the reference's atomicMin over point indices), every other point 0.
// Rows with map[x] == -1 get 0.  Unlike the reference there is no fill_(0) + scattered write: both variants are one
// coalesced point-major pass that writes every element of grad_feats exactly once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_backward_add_kernel(const float* __restrict__ grad_reduced,
                                                                   const int* __restrict__ coors_map,
                                                                   const int* __restrict__ reduce_count, long long total,
                                                                   int c, int mean, float* __restrict__ grad_feats) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t / c), i = (int)(t % c);
    const int to = coors_map[x];
    float g = 0.f;
    if (to >= 0) {
      g = grad_reduced[(long long)to * c + i];
      if (mean) g = g / (float)reduce_count[to];
    }
    grad_feats[t] = g;
  }
}

__global__ void __launch_bounds__(256) scatter_backward_argmax_kernel(const float* __restrict__ feats,
                                                                      const float* __restrict__ reduced,
                                                                      const int* __restrict__ coors_map, long long total,
                                                                      int c, int* __restrict__ reduce_from) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t / c), i = (int)(t % c);
    const int to = coors_map[x];
    if (to < 0) continue;
    const long long r = (long long)to * c + i;
    if (feats[t] == reduced[r]) atomicMin(&reduce_from[r], x);
  }
}

__global__ void __launch_bounds__(256) scatter_backward_max_kernel(const float* __restrict__ grad_reduced,
                                                                   const int* __restrict__ coors_map,
                                                                   const int* __restrict__ reduce_from, long long total,
                                                                   int c, float* __restrict__ grad_feats) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t / c), i = (int)(t % c);
    const int to = coors_map[x];
    float g = 0.f;
    if (to >= 0) {
      const long long r = (long long)to * c + i;
      if (reduce_from[r] == x) g = grad_reduced[r];
    }
    grad_feats[t] = g;
  }
}

extern "C" int isf_dynamic_scatter_backward_workspace_size(int num_reduced, int c, size_t* bytes) {
  ISF_CHECK_ARG(bytes && num_reduced >= 0 && c >= 1);
  *bytes = (size_t)num_reduced * c * sizeof(int) + 256;
  return ISF_OK;
}

extern "C" int isf_dynamic_point_to_voxel_backward(float* grad_feats, const float* grad_reduced, const float* feats,
                                                   const float* reduced, const int* coors_map, const int* reduce_count,
                                                   int n, int num_reduced, int c, int reduce_type, void* workspace,
                                                   size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(n >= 0 && num_reduced >= 0 && c >= 1 && reduce_type >= 0 && reduce_type <= 2);
  if (n == 0) return ISF_OK;
  ISF_CHECK_ARG(grad_feats && coors_map);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * c;
  if (num_reduced == 0) {                   // every point was dropped: the gradient is zero
    ISF_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * total, st));
    return ISF_OK;
  }
  ISF_CHECK_ARG(grad_reduced);
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  if (reduce_type != 2) {                   // 0 = sum, 1 = mean (same codes as the forward)
    ISF_CHECK_ARG(reduce_type == 0 || reduce_count);
    scatter_backward_add_kernel<<<blocks, 256, 0, st>>>(grad_reduced, coors_map, reduce_count, total, c, reduce_type == 1,
                                                        grad_feats);
    isf_count_launch(1);
  } else {
    ISF_CHECK_ARG(feats && reduced && workspace);
    if (ws_bytes < (size_t)num_reduced * c * sizeof(int)) { isf_set_error("dynamic scatter backward: workspace too small"); return ISF_ERR_WORKSPACE; }
    int* reduce_from = (int*)workspace;
    ISF_CUDA(cudaMemsetAsync(reduce_from, 0x7F, sizeof(int) * (size_t)num_reduced * c, st));   // 0x7F7F7F7F > any point index
    scatter_backward_argmax_kernel<<<blocks, 256, 0, st>>>(feats, reduced, coors_map, total, c, reduce_from);
    scatter_backward_max_kernel<<<blocks, 256, 0, st>>>(grad_reduced, coors_map, reduce_from, total, c, grad_feats);
    isf_count_launch(3);
  }
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// Input contract: multi-sweep merge on the device
//   replaces the per-sweep numpy work of LoadPointsFromMultiSweeps.__call__ (mmdet3d/datasets/pipelines/loading.py:
//   846-906, _remove_close :824-844): the raw key frame + sweeps arrive as ONE [M, load_dim] fp32 block (one H2D copy from
//   pinned memory); per segment (key frame / sweep) an optional "remove points with |x| < r and |y| < r", the rigid
//   transform into the key frame's LiDAR frame and the time-lag column.  Arithmetic as numpy does it: float32 points,
//   float64 calibration, x' = fl32(sum_j p_j R_ij) then fl32(x' + t_i); column 4 = fl32(lag).  Row order is preserved
//   (hard voxelization keeps the FIRST 12 points of a pillar): stable compaction = flags -> exclusive scan -> scatter.
// ------------------------------------------------------------------------------------------------
constexpr int MERGE_MAX_SEGMENTS = 24;
struct MergeTable {
  int n;
  int start[MERGE_MAX_SEGMENTS + 1];
  int flags[MERGE_MAX_SEGMENTS];          // bit 0: rotate + translate, bit 1: remove close points
  double rot[MERGE_MAX_SEGMENTS][9];      // sensor2lidar_rotation, row-major
  double trans[MERGE_MAX_SEGMENTS][3];
  double lag[MERGE_MAX_SEGMENTS];
  float radius;
  int use_range;                          // PointsRangeFilter fused: keep only points STRICTLY inside range (after the transform)
  float range[6];
};

__host__ __device__ __forceinline__ int merge_segment_of(const MergeTable& T, int row) {
  int s = 0;
  while (s + 1 < T.n && row >= T.start[s + 1]) ++s;
  return s;
}

__host__ __device__ __forceinline__ bool merge_keeps(const MergeTable& T, int s, const float* p) {
  if (!(T.flags[s] & 2)) return true;
  return !(fabsf(p[0]) < T.radius && fabsf(p[1]) < T.radius);
}

// BasePoints.in_range_3d (core/points/base_points.py:224-229) on a merged point: strictly inside on all six faces
__host__ __device__ __forceinline__ bool merge_in_range(const MergeTable& T, const float* v) {
  return v[0] > T.range[0] && v[1] > T.range[1] && v[2] > T.range[2] && v[0] < T.range[3] && v[1] < T.range[4] && v[2] < T.range[5];
}

// one point of segment s: in = raw row (>= 5 columns), out = 5 columns
__host__ __device__ __forceinline__ void merge_point(const MergeTable& T, int s, const float* in, float* out) {
  float x = in[0], y = in[1], z = in[2];
  if (T.flags[s] & 1) {
    const double* R = T.rot[s];
    const double px = (double)x, py = (double)y, pz = (double)z;
    // the row-times-R^T product in float64, accumulated in column order like the BLAS kernel numpy calls
    const float rx = (float)fma(pz, R[2], fma(py, R[1], px * R[0]));
    const float ry = (float)fma(pz, R[5], fma(py, R[4], px * R[3]));
    const float rz = (float)fma(pz, R[8], fma(py, R[7], px * R[6]));
    x = (float)((double)rx + T.trans[s][0]);
    y = (float)((double)ry + T.trans[s][1]);
    z = (float)((double)rz + T.trans[s][2]);
  }
  out[0] = x; out[1] = y; out[2] = z; out[3] = in[3]; out[4] = (float)T.lag[s];
}

__global__ void __launch_bounds__(256) merge_flags_kernel(const float* __restrict__ raw, int m, int load_dim,
                                                          const __grid_constant__ MergeTable T, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int s = merge_segment_of(T, i);
  bool keep = merge_keeps(T, s, raw + (size_t)i * load_dim);
  if (keep && T.use_range) {
    float v[5];
    merge_point(T, s, raw + (size_t)i * load_dim, v);
    keep = merge_in_range(T, v);
  }
  flags[i] = keep ? 1 : 0;
}

// pos == nullptr: no segment removes points (out row = in row)
__global__ void __launch_bounds__(256) merge_write_kernel(const float* __restrict__ raw, int m, int load_dim,
                                                          const __grid_constant__ MergeTable T, const int* __restrict__ flags,
                                                          const int* __restrict__ pos, float* __restrict__ out,
                                                          int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const bool keep = pos ? flags[i] != 0 : true;
  const int o = pos ? pos[i] : i;
  if (i == m - 1 && count) *count = o + (keep ? 1 : 0);
  if (!keep) return;
  float v[5];
  merge_point(T, merge_segment_of(T, i), raw + (size_t)i * load_dim, v);
#pragma unroll
  for (int c = 0; c < 5; ++c) out[(size_t)o * 5 + c] = v[c];
}

static int merge_fill_table(MergeTable* T, int m, int num_segments, const int* seg_start, const double* rot,
                            const double* trans, const double* lag, const int* seg_flags, float radius, const float* range, bool* any_remove) {
  ISF_CHECK_ARG(num_segments >= 1 && num_segments <= MERGE_MAX_SEGMENTS && seg_start && seg_flags && lag);
  ISF_CHECK_ARG(seg_start[0] == 0 && seg_start[num_segments] == m);
  T->n = num_segments;
  T->radius = radius;
  T->use_range = range ? 1 : 0;
  for (int e = 0; e < 6; ++e) T->range[e] = range ? range[e] : 0.f;
  *any_remove = range != nullptr;
  for (int s = 0; s < num_segments; ++s) {
    ISF_CHECK_ARG(seg_start[s + 1] >= seg_start[s]);
    T->start[s] = seg_start[s];
    T->flags[s] = seg_flags[s] & 3;
    T->lag[s] = lag[s];
    if (seg_flags[s] & 2) *any_remove = true;
    if (seg_flags[s] & 1) {
      ISF_CHECK_ARG(rot && trans);
      for (int e = 0; e < 9; ++e) T->rot[s][e] = rot[s * 9 + e];
      for (int e = 0; e < 3; ++e) T->trans[s][e] = trans[s * 3 + e];
    }
  }
  T->start[num_segments] = m;
  return ISF_OK;
}

extern "C" int isf_merge_sweeps_workspace_size(int m, size_t* bytes) {
  ISF_CHECK_ARG(bytes && m >= 0);
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, m > 0 ? m : 1);
  *bytes = 2 * ((size_t)(m > 0 ? m : 1) * sizeof(int) + 256) + cub_bytes + 256;
  return ISF_OK;
}

// raw [m, load_dim] fp32 (load_dim >= 5; DEVICE), segment tables on the HOST (<= 24 segments: key frame + sweeps),
// out [m, 5] fp32 (capacity m rows), *count_dev = rows written.  seg_flags: bit 0 = transform, bit 1 = remove close.
// range (HOST [6] or NULL) = PointsRangeFilter fused: only points strictly inside (x0,y0,z0,x1,y1,z1) after the transform.
extern "C" int isf_merge_sweeps(const float* raw, int m, int load_dim, int num_segments, const int* seg_start,
                                const double* rot, const double* trans, const double* lag, const int* seg_flags, float radius,
                                const float* range, float* out, int* count_dev, void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(m >= 0 && load_dim >= 5 && count_dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) { ISF_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int), st)); return ISF_OK; }
  ISF_CHECK_ARG(raw && out);
  MergeTable T;
  bool any_remove = false;
  int rc = merge_fill_table(&T, m, num_segments, seg_start, rot, trans, lag, seg_flags, radius, range, &any_remove);
  if (rc) return rc;
  const int blocks = isf_cdiv(m, 256);
  if (!any_remove) {
    merge_write_kernel<<<blocks, 256, 0, st>>>(raw, m, load_dim, T, nullptr, nullptr, out, count_dev);
    isf_count_launch(1);
  } else {
    IsfArena A(workspace, ws_bytes);
    int* flags = A.take<int>(m);
    int* pos = A.take<int>(m);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, m);
    char* tmp = A.take<char>(cub_bytes);
    if (!A.ok || !workspace) { isf_set_error("merge_sweeps: workspace too small"); return ISF_ERR_WORKSPACE; }
    merge_flags_kernel<<<blocks, 256, 0, st>>>(raw, m, load_dim, T, flags);
    ISF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, cub_bytes, flags, pos, m, st));
    merge_write_kernel<<<blocks, 256, 0, st>>>(raw, m, load_dim, T, flags, pos, out, count_dev);
    isf_count_launch(4);
  }
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// TEST HOOK (host memory, no GPU): merge_segment_of / merge_keeps / merge_point — the kernels' own source — on the CPU.
extern "C" int isf_debug_merge_sweeps_host(const float* raw, int m, int load_dim, int num_segments, const int* seg_start,
                                           const double* rot, const double* trans, const double* lag, const int* seg_flags,
                                           float radius, const float* range, float* out, int* count) {
  ISF_CHECK_ARG(m >= 0 && load_dim >= 5 && count && (m == 0 || (raw && out)));
  *count = 0;
  if (m == 0) return ISF_OK;
  MergeTable T;
  bool any_remove = false;
  int rc = merge_fill_table(&T, m, num_segments, seg_start, rot, trans, lag, seg_flags, radius, range, &any_remove);
  if (rc) return rc;
  int o = 0;
  for (int i = 0; i < m; ++i) {
    const int s = merge_segment_of(T, i);
    const float* p = raw + (size_t)i * load_dim;
    if (!merge_keeps(T, s, p)) continue;
    float v[5];
    merge_point(T, s, p, v);
    if (T.use_range && !merge_in_range(T, v)) continue;
    for (int c = 0; c < 5; ++c) out[(size_t)o * 5 + c] = v[c];
    ++o;
  }
  *count = o;
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// K7: ingroup indices  (TorchEx/.../ingroup_inds_kernel.cu:17-30: rank of each element inside its
// group; the reference hands ranks out in atomic arrival order, here rank = stable index order)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keys_from_i64_kernel(const long long* __restrict__ g, int n,
                                                            unsigned long long* __restrict__ keys,
                                                            int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = g[i];
  keys[i] = v < 0 ? ~0ull : (unsigned long long)v;
  idx[i] = i;
}

__global__ void __launch_bounds__(256) ingroup_rank_kernel(const int* __restrict__ idx_sorted,
                                                           const int* __restrict__ start,
                                                           const int* __restrict__ num_runs, int n,
                                                           long long* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int lo = 0, hi = *num_runs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (start[mid] <= j) lo = mid; else hi = mid - 1;
  }
  out[idx_sorted[j]] = (long long)(j - start[lo]);
}

extern "C" int isf_ingroup_indices(const long long* group_inds, int n, int key_bits, long long* out_inds,
                                   void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(n >= 0 && key_bits >= 1 && key_bits <= 62);
  if (n == 0) return ISF_OK;
  ISF_CHECK_ARG(group_inds && out_inds && workspace);
  cudaStream_t st = (cudaStream_t)stream;
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  group_carve(A, n, &G);
  if (!A.ok) { isf_set_error("ingroup_indices: workspace too small"); return ISF_ERR_WORKSPACE; }
  keys_from_i64_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(group_inds, n, G.keys, G.idx);
  int rc = group_by_key(G, n, key_bits, nullptr, st);
  if (rc) return rc;
  ingroup_rank_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(G.idx_sorted, G.start, G.num_runs, n, out_inds);
  isf_count_launch(2);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// a5: deterministic hard voxelization (pillars)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pillar_keys_kernel(const float* __restrict__ points, int n, int nfeat,
                                                          VoxGeom g, unsigned long long* __restrict__ keys,
                                                          int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = points + (size_t)i * nfeat;
  int cx = vox_coord(p[0], g.xmin, g.vx);
  int cy = vox_coord(p[1], g.ymin, g.vy);
  int cz = vox_coord(p[2], g.zmin, g.vz);
  bool bad = cx < 0 || cx >= g.gx || cy < 0 || cy >= g.gy || cz < 0 || cz >= g.gz;
  keys[i] = bad ? ~0ull : (unsigned long long)(((long long)cz * g.gy + cy) * g.gx + cx);
  idx[i] = i;
}

// first point index of every valid run -> key for the second (voxel-order) sort
__global__ void __launch_bounds__(256) first_point_kernel(const int* __restrict__ idx_sorted,
                                                          const int* __restrict__ start,
                                                          const int* __restrict__ num_valid, int n,
                                                          unsigned int* __restrict__ first_pt,
                                                          int* __restrict__ run_id) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (r < *num_valid) { first_pt[r] = (unsigned)idx_sorted[start[r]]; }
  else { first_pt[r] = 0xffffffffu; }
  run_id[r] = r;
}

__global__ void __launch_bounds__(256) write_pillars_kernel(
    const float* __restrict__ points, int nfeat, VoxGeom g, int max_points, int max_voxels,
    const int* __restrict__ idx_sorted, const unsigned long long* __restrict__ uniq,
    const int* __restrict__ cnt, const int* __restrict__ start, const int* __restrict__ num_valid,
    const int* __restrict__ run_sorted, float* __restrict__ voxels, int* __restrict__ coors,
    int* __restrict__ npv, int* __restrict__ voxel_num) {
  const int nv = *num_valid;
  const int m_out = (max_voxels >= 0 && nv > max_voxels) ? max_voxels : nv;
  if (blockIdx.x == 0 && threadIdx.x == 0) *voxel_num = m_out;
  const int per = max_points * nfeat;
  const long long total = (long long)m_out * per;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int v = (int)(t / per);      // output voxel number (order of first appearance)
    int rem = (int)(t % per);
    int slot = rem / nfeat, f = rem % nfeat;
    int r = run_sorted[v];
    int m = cnt[r];
    float val = 0.f;
    if (slot < m) val = points[(size_t)idx_sorted[start[r] + slot] * nfeat + f];
    voxels[t] = val;
    if (rem == 0) {
      unsigned long long k = uniq[r];
      int cx = (int)(k % g.gx);
      int cy = (int)((k / g.gx) % g.gy);
      int cz = (int)(k / ((unsigned long long)g.gx * g.gy));
      coors[(size_t)v * 3 + 0] = cz;
      coors[(size_t)v * 3 + 1] = cy;
      coors[(size_t)v * 3 + 2] = cx;
      npv[v] = m < max_points ? m : max_points;
    }
  }
}

static void hard_carve(IsfArena& A, int n, GroupBuffers* G, unsigned int** first_pt, unsigned int** first_sorted,
                       int** run_id, int** run_sorted, void** tmp2, size_t* tmp2_bytes) {
  group_carve(A, n, G);
  *first_pt = A.take<unsigned int>(n);
  *first_sorted = A.take<unsigned int>(n);
  *run_id = A.take<int>(n);
  *run_sorted = A.take<int>(n);
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (unsigned int*)nullptr, (unsigned int*)nullptr, (int*)nullptr,
                                  (int*)nullptr, n);
  *tmp2_bytes = b;
  *tmp2 = A.take<char>(b);
}

extern "C" int isf_hard_voxelize_workspace_size(int n, size_t* bytes) {
  ISF_CHECK_ARG(n >= 0 && bytes);
  IsfArena A(nullptr, 0);
  GroupBuffers G;
  unsigned int *a, *b; int *c, *d; void* t; size_t tb;
  hard_carve(A, n > 0 ? n : 1, &G, &a, &b, &c, &d, &t, &tb);
  *bytes = A.off + 1024;
  return ISF_OK;
}

extern "C" int isf_hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                                 const float* coors_range, int max_points, int max_voxels, float* voxels,
                                 int* coors, int* num_points_per_voxel, int* voxel_num_dev, void* workspace,
                                 size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(n >= 0 && nfeat >= 3 && max_points >= 1 && voxel_size && coors_range && voxel_num_dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ISF_CUDA(cudaMemsetAsync(voxel_num_dev, 0, sizeof(int), st));
    return ISF_OK;
  }
  ISF_CHECK_ARG(points && voxels && coors && num_points_per_voxel && workspace);
  VoxGeom g;
  make_geom(voxel_size, coors_range, &g);
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  unsigned int *first_pt, *first_sorted; int *run_id, *run_sorted; void* tmp2; size_t tmp2_bytes;
  hard_carve(A, n, &G, &first_pt, &first_sorted, &run_id, &run_sorted, &tmp2, &tmp2_bytes);
  if (!A.ok) { isf_set_error("hard_voxelize: workspace too small"); return ISF_ERR_WORKSPACE; }
  long long cells = (long long)g.gx * g.gy * g.gz;
  ISF_CHECK_ARG(cells > 0 && cells < (1ll << 62));
  pillar_keys_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(points, n, nfeat, g, G.keys, G.idx);
  int rc = group_by_key(G, n, ceil_log2(cells), nullptr, st);
  if (rc) return rc;
  first_point_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(G.idx_sorted, G.start, G.num_valid, n, first_pt, run_id);
  ISF_CUDA(cub::DeviceRadixSort::SortPairs(tmp2, tmp2_bytes, first_pt, first_sorted, run_id, run_sorted, n, 0, 32,
                                           st));
  int nv_cap = (max_voxels >= 0 && max_voxels < n) ? max_voxels : n;
  long long total = (long long)nv_cap * max_points * nfeat;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  if (blocks < 1) blocks = 1;
  write_pillars_kernel<<<blocks, 256, 0, st>>>(points, nfeat, g, max_points, max_voxels, G.idx_sorted, G.uniq,
                                               G.cnt, G.start, G.num_valid, run_sorted, voxels, coors,
                                               num_points_per_voxel, voxel_num_dev);
  isf_count_launch(5);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// a3: fused DynamicVFE  (11-ch point features -> Linear 11->64 + BN + ReLU -> voxel max -> concat ->
// Linear 128->64 + BN + ReLU -> voxel max), one warp per voxel, no intermediate [N,64] tensors.
// ------------------------------------------------------------------------------------------------
struct VfeParams {
  float vx, vy, vz;        // voxel size (fp32, as the reference multiplies coors by it)
  float xoff, yoff, zoff;  // v/2 + range_min, computed in double then rounded (voxel_encoder.py:351-353)
};

constexpr int VFE_C = 64;
constexpr int VFE_IN = 11;
constexpr int VFE_WARPS = 8;

__global__ void __launch_bounds__(VFE_WARPS * 32) vfe_fused_kernel(
    const float* __restrict__ points, int nfeat, const int* __restrict__ idx_sorted,
    const unsigned long long* __restrict__ uniq, const int* __restrict__ cnt, const int* __restrict__ start,
    const int* __restrict__ num_valid, KeyLayout L, VfeParams P,
    const float* __restrict__ w1,  // [64][11]
    const float* __restrict__ s1, const float* __restrict__ b1,
    const float* __restrict__ w2,  // [64][128]
    const float* __restrict__ s2, const float* __restrict__ b2,
    float* __restrict__ out_f32, __half* __restrict__ out_f16, int* __restrict__ out_coors) {
  __shared__ float w2t[2 * VFE_C][VFE_C];
  __shared__ __align__(16) float hbuf[VFE_WARPS][VFE_C][4];   // layer-1 activations of up to 4 points, [k][point]
  for (int t = threadIdx.x; t < 2 * VFE_C * VFE_C; t += blockDim.x)
    w2t[t % (2 * VFE_C)][t / (2 * VFE_C)] = w2[t];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // layer-1 weights of this lane's two output channels live in registers
  float w1a[VFE_IN], w1b[VFE_IN];
#pragma unroll
  for (int k = 0; k < VFE_IN; ++k) { w1a[k] = w1[lane * VFE_IN + k]; w1b[k] = w1[(lane + 32) * VFE_IN + k]; }
  __syncthreads();
  const int nv = *num_valid;
  const float s1a = s1[lane], s1b = s1[lane + 32], b1a = b1[lane], b1b = b1[lane + 32];
  const float s2a = s2[lane], s2b = s2[lane + 32], b2a = b2[lane], b2b = b2[lane + 32];
  float (*hb)[4] = hbuf[wid];
  for (int v = blockIdx.x * VFE_WARPS + wid; v < nv; v += gridDim.x * VFE_WARPS) {
    const int s = start[v], m = cnt[v];
    int cc[4];
    unpack_key(uniq[v], L, cc);  // (b, z, y, x)
    // voxel mean of xyz (cluster centre), summed in sorted (= index) order like index_add
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int j = 0; j < m; ++j) {
      const float* p = points + (size_t)idx_sorted[s + j] * nfeat;
      sx += p[0]; sy += p[1]; sz += p[2];
    }
    const float fm = (float)m;
    const float mx = sx / fm, my = sy / fm, mz = sz / fm;
    // voxel centre: coors*v + offset, fp32 mul then add (no FMA), as torch does
    const float ccx = __fadd_rn(__fmul_rn((float)cc[3], P.vx), P.xoff);
    const float ccy = __fadd_rn(__fmul_rn((float)cc[2], P.vy), P.yoff);
    const float ccz = __fadd_rn(__fmul_rn((float)cc[1], P.vz), P.zoff);
    auto layer1 = [&](int j, float& ha, float& hc) {
      const float* p = points + (size_t)idx_sorted[s + j] * nfeat;
      float f[VFE_IN];
      f[0] = p[0]; f[1] = p[1]; f[2] = p[2]; f[3] = p[3]; f[4] = p[4];
      f[5] = __fsub_rn(f[0], mx); f[6] = __fsub_rn(f[1], my); f[7] = __fsub_rn(f[2], mz);
      f[8] = __fsub_rn(f[0], ccx); f[9] = __fsub_rn(f[1], ccy); f[10] = __fsub_rn(f[2], ccz);
      ha = 0.f; hc = 0.f;
#pragma unroll
      for (int k = 0; k < VFE_IN; ++k) { ha = fmaf(w1a[k], f[k], ha); hc = fmaf(w1b[k], f[k], hc); }
      ha = fmaxf(fmaf(ha, s1a, b1a), 0.f);
      hc = fmaxf(fmaf(hc, s1b, b1b), 0.f);
    };
    // pass 1: voxel max of the layer-1 activations (post-ReLU values are >= 0 and every voxel has >= 1 point)
    float vmaxa = 0.f, vmaxb = 0.f;
    for (int j = 0; j < m; ++j) {
      float ha, hc;
      layer1(j, ha, hc);
      vmaxa = fmaxf(vmaxa, ha); vmaxb = fmaxf(vmaxb, hc);
    }
    // per-voxel half of layer 2: W2[:, 64:128] . vmax1
    __syncwarp();
    hb[lane][0] = vmaxa; hb[lane + 32][0] = vmaxb;
    __syncwarp();
    float ga = 0.f, gb = 0.f;
#pragma unroll 8
    for (int k = 0; k < VFE_C; ++k) {
      const float h = hb[k][0];
      ga = fmaf(w2t[VFE_C + k][lane], h, ga);
      gb = fmaf(w2t[VFE_C + k][lane + 32], h, gb);
    }
    // pass 2: per-point half of layer 2, four points per sweep over the weights
    vmaxa = 0.f; vmaxb = 0.f;
    for (int j0 = 0; j0 < m; j0 += 4) {
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float ha = 0.f, hc = 0.f;
        if (j0 + q < m) layer1(j0 + q, ha, hc);
        hb[lane][q] = ha; hb[lane + 32][q] = hc;
      }
      __syncwarp();
      float oa[4] = {ga, ga, ga, ga}, ob[4] = {gb, gb, gb, gb};
#pragma unroll 8
      for (int k = 0; k < VFE_C; ++k) {
        const float4 h = *(const float4*)hb[k];
        const float wa = w2t[k][lane], wb = w2t[k][lane + 32];
        oa[0] = fmaf(wa, h.x, oa[0]); oa[1] = fmaf(wa, h.y, oa[1]); oa[2] = fmaf(wa, h.z, oa[2]); oa[3] = fmaf(wa, h.w, oa[3]);
        ob[0] = fmaf(wb, h.x, ob[0]); ob[1] = fmaf(wb, h.y, ob[1]); ob[2] = fmaf(wb, h.z, ob[2]); ob[3] = fmaf(wb, h.w, ob[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (j0 + q < m) {
          vmaxa = fmaxf(vmaxa, fmaxf(fmaf(oa[q], s2a, b2a), 0.f));
          vmaxb = fmaxf(vmaxb, fmaxf(fmaf(ob[q], s2b, b2b), 0.f));
        }
      }
    }
    if (out_f32) { out_f32[(size_t)v * VFE_C + lane] = vmaxa; out_f32[(size_t)v * VFE_C + lane + 32] = vmaxb; }
    if (out_f16) {
      out_f16[(size_t)v * VFE_C + lane] = __float2half_rn(vmaxa);
      out_f16[(size_t)v * VFE_C + lane + 32] = __float2half_rn(vmaxb);
    }
    if (out_coors && lane < 4) out_coors[(size_t)v * 4 + lane] = cc[lane];
  }
}

extern "C" int isf_dynamic_vfe_workspace_size(int n, size_t* bytes) {
  return isf_dynamic_scatter_workspace_size(n, bytes);
}

// points [n,nfeat>=5] fp32, coors [n,4] int32 (b,z,y,x) all in range; dims[4] = (B,Z,Y,X) extents.
extern "C" int isf_dynamic_vfe_forward(const float* points, const int* coors, int n, int nfeat, const int* dims,
                                       const float* voxel_size, const float* offsets, const float* w1,
                                       const float* scale1, const float* shift1, const float* w2,
                                       const float* scale2, const float* shift2, int in_ch, int feat_ch,
                                       float* voxel_feats_f32, void* voxel_feats_f16, int* voxel_coors,
                                       int* num_voxels_dev, void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(n >= 0 && nfeat >= 5 && dims && voxel_size && offsets && num_voxels_dev);
  if (in_ch != VFE_IN || feat_ch != VFE_C) {
    isf_set_error("dynamic_vfe: only in=11 (5+3+3), feat=(64,64) is built (isfusion_0075voxel.py:64-71)");
    return ISF_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ISF_CUDA(cudaMemsetAsync(num_voxels_dev, 0, sizeof(int), st));
    return ISF_OK;
  }
  ISF_CHECK_ARG(points && coors && w1 && scale1 && shift1 && w2 && scale2 && shift2 && voxel_coors && workspace);
  ISF_CHECK_ARG(voxel_feats_f32 || voxel_feats_f16);
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  group_carve(A, n, &G);
  if (!A.ok) { isf_set_error("dynamic_vfe: workspace too small"); return ISF_ERR_WORKSPACE; }
  KeyLayout L = make_layout(4, dims);
  make_keys_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(coors, n, L, G.keys, G.idx);
  int rc = group_by_key(G, n, L.total_bits, num_voxels_dev, st);
  if (rc) return rc;
  VfeParams P{voxel_size[0], voxel_size[1], voxel_size[2], offsets[0], offsets[1], offsets[2]};
  // persistent warps: exactly one resident wave (a partial second wave would double the kernel time)
  int occ = 0;
  {   // occupancy is per-device state
    static int occ_dev[64] = {};
    int dev = 0;
    ISF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (!occ_dev[dev]) {
      ISF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dev[dev], vfe_fused_kernel, VFE_WARPS * 32, 0));
      if (occ_dev[dev] < 1) occ_dev[dev] = 1;
    }
    occ = occ_dev[dev];
  }
  int blocks = isf_cdiv(n, VFE_WARPS);
  if (blocks > ISF_NUM_SMS * occ) blocks = ISF_NUM_SMS * occ;
  vfe_fused_kernel<<<blocks, VFE_WARPS * 32, 0, st>>>(points, nfeat, G.idx_sorted, G.uniq, G.cnt, G.start,
                                                      G.num_valid, L, P, w1, scale1, shift1, w2, scale2, shift2,
                                                      voxel_feats_f32, (__half*)voxel_feats_f16, voxel_coors);
  isf_count_launch(2);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// a3 split in two calls so that the voxel COORDINATES (all the sparse-conv rulebooks need) are available before
// the feature kernel has run: group (sort + run-length encode + coordinates + count), then compute (fused VFE on
// the grouping left in the same workspace).  isf_dynamic_vfe_forward == group + compute.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_coors_kernel(const unsigned long long* __restrict__ uniq,
                                                          const int* __restrict__ num_valid, KeyLayout L, int n,
                                                          int* __restrict__ out_coors) {
  const int nv = *num_valid;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    int cc[4];
    unpack_key(uniq[v], L, cc);
    *(int4*)(out_coors + (size_t)v * 4) = make_int4(cc[0], cc[1], cc[2], cc[3]);
  }
}

extern "C" int isf_dynamic_vfe_group(const int* coors, int n, const int* dims, int* voxel_coors, int* num_voxels_dev,
                                     void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(n >= 0 && dims && num_voxels_dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { ISF_CUDA(cudaMemsetAsync(num_voxels_dev, 0, sizeof(int), st)); return ISF_OK; }
  ISF_CHECK_ARG(coors && voxel_coors && workspace);
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  group_carve(A, n, &G);
  if (!A.ok) { isf_set_error("dynamic_vfe_group: workspace too small"); return ISF_ERR_WORKSPACE; }
  KeyLayout L = make_layout(4, dims);
  make_keys_kernel<<<isf_cdiv(n, 256), 256, 0, st>>>(coors, n, L, G.keys, G.idx);
  int rc = group_by_key(G, n, L.total_bits, num_voxels_dev, st);
  if (rc) return rc;
  int blocks = isf_cdiv(n, 256);
  if (blocks > ISF_NUM_SMS * 8) blocks = ISF_NUM_SMS * 8;
  voxel_coors_kernel<<<blocks, 256, 0, st>>>(G.uniq, G.num_valid, L, n, voxel_coors);
  isf_count_launch(2);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// must follow isf_dynamic_vfe_group on the same stream with the same (n, dims, workspace)
extern "C" int isf_dynamic_vfe_compute(const float* points, int n, int nfeat, const int* dims, const float* voxel_size,
                                       const float* offsets, const float* w1, const float* scale1, const float* shift1,
                                       const float* w2, const float* scale2, const float* shift2, int in_ch, int feat_ch,
                                       float* voxel_feats_f32, void* voxel_feats_f16, void* workspace, size_t ws_bytes,
                                       void* stream) {
  ISF_CHECK_ARG(n >= 0 && nfeat >= 5 && dims && voxel_size && offsets);
  if (in_ch != VFE_IN || feat_ch != VFE_C) { isf_set_error("dynamic_vfe: only in=11, feat=(64,64) is built"); return ISF_ERR_UNSUPPORTED; }
  if (n == 0) return ISF_OK;
  ISF_CHECK_ARG(points && w1 && scale1 && shift1 && w2 && scale2 && shift2 && workspace && (voxel_feats_f32 || voxel_feats_f16));
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  group_carve(A, n, &G);
  if (!A.ok) { isf_set_error("dynamic_vfe_compute: workspace too small"); return ISF_ERR_WORKSPACE; }
  KeyLayout L = make_layout(4, dims);
  VfeParams P{voxel_size[0], voxel_size[1], voxel_size[2], offsets[0], offsets[1], offsets[2]};
  int occ = 0;
  {   // occupancy is per-device state
    static int occ_dev[64] = {};
    int dev = 0;
    ISF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (!occ_dev[dev]) {
      ISF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dev[dev], vfe_fused_kernel, VFE_WARPS * 32, 0));
      if (occ_dev[dev] < 1) occ_dev[dev] = 1;
    }
    occ = occ_dev[dev];
  }
  int blocks = isf_cdiv(n, VFE_WARPS);
  if (blocks > ISF_NUM_SMS * occ) blocks = ISF_NUM_SMS * occ;
  vfe_fused_kernel<<<blocks, VFE_WARPS * 32, 0, (cudaStream_t)stream>>>(points, nfeat, G.idx_sorted, G.uniq, G.cnt, G.start,
                                                                      G.num_valid, L, P, w1, scale1, shift1, w2, scale2, shift2,
                                                                      voxel_feats_f32, (__half*)voxel_feats_f16, nullptr);
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// a3, tensor-core formulation (default): the 64x64 halves of layer 2 are GEMMs on tcgen05, everything else is
// a light streaming kernel.  With h1 = relu(bn1(W1 f)) per point and m1 = per-voxel max of h1:
//   layer 2 = relu(bn2(W2[:, :64] h1 + W2[:, 64:] m1))  ->  g' = (W2[:, 64:] m1) * scale2           (GEMM over voxels)
//                                                           y  = relu((W2[:, :64] h1) * scale2 + shift2 + g'[voxel(p)])
//                                                                                                   (GEMM over points)
// and the output is the per-voxel max of y.  Layer 1 (11 inputs, raw coordinates) stays fp32; h1 / g' / y are stored
// fp16 (fp32 accumulate).  The ncu capture of the SIMT kernel showed it shared-memory bound (L1TEX 80 %, 265 M warp
// instructions, 0.41 ms); this path moves ~190 MB through HBM instead.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vfe_voxel_prep_kernel(const float* __restrict__ points, int nfeat,
                                                             const int* __restrict__ idx_sorted,
                                                             const int* __restrict__ cnt, const int* __restrict__ start,
                                                             const int* __restrict__ num_valid, int n,
                                                             float4* __restrict__ mean, int* __restrict__ seg_of) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nv = *num_valid;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < nv; v += warps) {
    const int s = start[v], m = cnt[v];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int j = lane; j < m; j += 32) {
      const float* p = points + (size_t)idx_sorted[s + j] * nfeat;
      sx += p[0]; sy += p[1]; sz += p[2];
      seg_of[s + j] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) mean[v] = make_float4(sx / (float)m, sy / (float)m, sz / (float)m, (float)m);
  }
  // sorted positions beyond the valid points (out-of-range padding) belong to no voxel
  if (blockIdx.x == 0 && threadIdx.x == 0 && nv > 0) seg_of[n] = start[nv - 1] + cnt[nv - 1];   // = number of valid points
}

// layer 1 per point (8 lanes per point, 8 channels per lane), fp32 math, fp16 row of h1 in sorted order
__global__ void __launch_bounds__(256, 2) vfe_point_layer1_kernel(const float* __restrict__ points, int nfeat,
                                                               const int* __restrict__ idx_sorted,
                                                               const unsigned long long* __restrict__ uniq,
                                                               const float4* __restrict__ mean,
                                                               const int* __restrict__ seg_of, int n, KeyLayout L, VfeParams P,
                                                               const float* __restrict__ w1, const float* __restrict__ s1,
                                                               const float* __restrict__ b1, __half* __restrict__ h1,
                                                               int* __restrict__ res_rows) {
  __shared__ __align__(16) float w1t[VFE_IN][VFE_C];
  for (int t = threadIdx.x; t < VFE_IN * VFE_C; t += blockDim.x) w1t[t % VFE_IN][t / VFE_IN] = w1[t];
  __syncthreads();
  const int nvalid_pts = seg_of[n];
  const int sub = threadIdx.x & 7;                    // channel group: channels sub*8 .. sub*8+7
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; j < n; j += (gridDim.x * blockDim.x) >> 3) {
    if (j >= nvalid_pts) break;                         // out-of-range padding: rows never read (GEMM stops at nvalid)
    uint4 o;
    int v;
    {
      v = seg_of[j];
      const float* p = points + (size_t)idx_sorted[j] * nfeat;
      const float4 mu = mean[v];
      int cc[4];
      unpack_key(uniq[v], L, cc);
      float f[VFE_IN];
      f[0] = p[0]; f[1] = p[1]; f[2] = p[2]; f[3] = p[3]; f[4] = p[4];
      f[5] = __fsub_rn(f[0], mu.x); f[6] = __fsub_rn(f[1], mu.y); f[7] = __fsub_rn(f[2], mu.z);
      f[8] = __fsub_rn(f[0], __fadd_rn(__fmul_rn((float)cc[3], P.vx), P.xoff));
      f[9] = __fsub_rn(f[1], __fadd_rn(__fmul_rn((float)cc[2], P.vy), P.yoff));
      f[10] = __fsub_rn(f[2], __fadd_rn(__fmul_rn((float)cc[1], P.vz), P.zoff));
      float h[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) h[c] = 0.f;
#pragma unroll
      for (int k = 0; k < VFE_IN; ++k) {
        const float4 wa = *(const float4*)&w1t[k][sub * 8], wb = *(const float4*)&w1t[k][sub * 8 + 4];
        h[0] = fmaf(wa.x, f[k], h[0]); h[1] = fmaf(wa.y, f[k], h[1]); h[2] = fmaf(wa.z, f[k], h[2]); h[3] = fmaf(wa.w, f[k], h[3]);
        h[4] = fmaf(wb.x, f[k], h[4]); h[5] = fmaf(wb.y, f[k], h[5]); h[6] = fmaf(wb.z, f[k], h[6]); h[7] = fmaf(wb.w, f[k], h[7]);
      }
      __half2* oh = (__half2*)&o;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ch = sub * 8 + 2 * c;
        const float a = fmaxf(fmaf(h[2 * c], s1[ch], b1[ch]), 0.f), b = fmaxf(fmaf(h[2 * c + 1], s1[ch + 1], b1[ch + 1]), 0.f);
        oh[c] = __floats2half2_rn(a, b);
      }
    }
    *(uint4*)(h1 + (size_t)j * VFE_C + sub * 8) = o;
    if (sub == 0) res_rows[j] = v;
  }
}

// per-voxel max over its (contiguous) rows: thread = (voxel, 8-channel group); values are >= 0 (post-ReLU)
__global__ void __launch_bounds__(256) vfe_voxel_max_kernel(const __half* __restrict__ rows, const int* __restrict__ cnt,
                                                            const int* __restrict__ start, const int* __restrict__ num_valid,
                                                            __half* __restrict__ out) {
  const int nv = *num_valid;
  const int sub = threadIdx.x & 7;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; v < nv; v += (gridDim.x * blockDim.x) >> 3) {
    const int s = start[v], m = cnt[v];
    __half2 mx[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) mx[c] = __floats2half2_rn(0.f, 0.f);
    for (int j = 0; j < m; ++j) {
      const uint4 u = *(const uint4*)(rows + (size_t)(s + j) * VFE_C + sub * 8);
      const __half2* h = (const __half2*)&u;
#pragma unroll
      for (int c = 0; c < 4; ++c) mx[c] = __hmax2(mx[c], h[c]);
    }
    uint4 o;
    o.x = *(const uint32_t*)&mx[0]; o.y = *(const uint32_t*)&mx[1]; o.z = *(const uint32_t*)&mx[2]; o.w = *(const uint32_t*)&mx[3];
    *(uint4*)(out + (size_t)v * VFE_C + sub * 8) = o;
  }
}

extern "C" int isf_dynamic_vfe_tc_workspace_size(int n, size_t* bytes) {
  ISF_CHECK_ARG(n >= 0 && bytes);
  size_t base = 0;
  int rc = isf_dynamic_scatter_workspace_size(n, &base);
  if (rc) return rc;
  const size_t nn = n > 0 ? n : 1;
  *bytes = base + isf_align_up((nn + 1) * sizeof(int)) * 2 + isf_align_up(nn * sizeof(float4)) +
           4 * isf_align_up(nn * VFE_C * sizeof(__half)) + 4096;
  return ISF_OK;
}

// tensor-core second half of the fused op; must follow isf_dynamic_vfe_group on the same stream with the same
// (n, dims) and a workspace of isf_dynamic_vfe_tc_workspace_size(n) bytes.  w2a / w2b: fp16 K-major [64][64]
// (= W2[:, :64] and W2[:, 64:]).  Output: voxel_feats_f16 [n,64] (rows < num_voxels valid).
extern "C" int isf_dynamic_vfe_compute_tc(const float* points, int n, int nfeat, const int* dims, const float* voxel_size,
                                          const float* offsets, const float* w1, const float* scale1, const float* shift1,
                                          const void* w2a, const void* w2b, const float* scale2, const float* shift2,
                                          void* voxel_feats_f16, void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(n >= 0 && nfeat >= 5 && dims && voxel_size && offsets);
  if (n == 0) return ISF_OK;
  ISF_CHECK_ARG(points && w1 && scale1 && shift1 && w2a && w2b && scale2 && shift2 && voxel_feats_f16 && workspace);
  cudaStream_t st = (cudaStream_t)stream;
  IsfArena A(workspace, ws_bytes);
  GroupBuffers G;
  group_carve(A, n, &G);                       // same carve as isf_dynamic_vfe_group: finds its buffers
  int* seg_of = A.take<int>(n + 1);
  int* res_rows = A.take<int>(n + 1);
  float4* mean = A.take<float4>(n);
  __half* h1 = A.take<__half>((size_t)n * VFE_C);
  __half* m1 = A.take<__half>((size_t)n * VFE_C);
  __half* g = A.take<__half>((size_t)n * VFE_C);
  __half* y = A.take<__half>((size_t)n * VFE_C);
  if (!A.ok) { isf_set_error("dynamic_vfe_compute_tc: workspace too small"); return ISF_ERR_WORKSPACE; }
  KeyLayout L = make_layout(4, dims);
  VfeParams P{voxel_size[0], voxel_size[1], voxel_size[2], offsets[0], offsets[1], offsets[2]};
  const int blocks = ISF_NUM_SMS * 8;
  ISF_CUDA(cudaMemsetAsync(seg_of + n, 0, sizeof(int), st));
  vfe_voxel_prep_kernel<<<blocks, 256, 0, st>>>(points, nfeat, G.idx_sorted, G.cnt, G.start, G.num_valid, n, mean, seg_of);
  vfe_point_layer1_kernel<<<blocks, 256, 0, st>>>(points, nfeat, G.idx_sorted, G.uniq, mean, seg_of, n, L, P, w1, scale1,
                                                  shift1, h1, res_rows);
  vfe_voxel_max_kernel<<<blocks, 256, 0, st>>>(h1, G.cnt, G.start, G.num_valid, m1);
  // g' = (m1 . W2b^T) * scale2          [voxels, 64]
  tc::GemmParams q{};
  q.A = m1; q.W = (const __half*)w2b; q.m_dev = G.num_valid; q.m = n; q.cin = VFE_C; q.taps = 1; q.kpad = VFE_C;
  q.n_total = VFE_C; q.scale = scale2; q.act = tc::ACT_NONE; q.out = g; q.ldo = VFE_C;
  int rc = tc::dispatch_gemm(q, VFE_C, n, st);
  if (rc) return rc;
  // y = relu((h1 . W2a^T) * scale2 + shift2 + g'[voxel(p)])     [points, 64]
  tc::GemmParams r{};
  r.A = h1; r.W = (const __half*)w2a; r.m = n; r.m_dev = seg_of + n;   // only the in-range points (device-side count)
  r.cin = VFE_C; r.taps = 1; r.kpad = VFE_C; r.n_total = VFE_C;
  r.scale = scale2; r.shift = shift2; r.residual = g; r.ldr = VFE_C; r.res_rows = res_rows; r.act = tc::ACT_RELU;
  r.out = y; r.ldo = VFE_C;
  rc = tc::dispatch_gemm(r, VFE_C, n, st);
  if (rc) return rc;
  vfe_voxel_max_kernel<<<blocks, 256, 0, st>>>(y, G.cnt, G.start, G.num_valid, (__half*)voxel_feats_f16);
  isf_count_launch(7);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}
