// Sparse 3-D convolution: rulebook construction + tcgen05 implicit GEMM.
//
// Replaces spconv's getIndicePair (mmdet3d/ops/bevfusion-ops/spconv/include/spconv/spconv_ops.h:27-141,
// indice.cu.h:22-203: dense B*D*H*W int32 grids, per-offset pair lists) and indiceConv (spconv_ops.h:260-361:
// 27 x {gather kernel, cuBLAS mm, scatter-add kernel} per layer).  Here:
//   * coordinates live in an open-addressing hash table (no dense grid);
//   * the rulebook is an output-stationary neighbour table nbr[tap][out_row] -> in_row | -1;
//   * strided-conv output sites = sorted unique of all candidate keys (one radix sort) = the reference's
//     GPU ordering (unique of the linear index, spconv_ops.h:129-137);
//   * the convolution is ONE kernel per layer (tc_gemm.cuh): gathered rows -> swizzled smem -> tcgen05.mma
//     accumulating all taps in TMEM -> fused BN scale/shift (+residual) (+ReLU) epilogue -> one fp16 store.
#include <cub/cub.cuh>
#include "tc_gemm.cuh"

extern "C" void isf_count_launch(int n);

struct ConvGeom {
  int in_shape[3];
  int out_shape[3];
  int ksize[3], stride[3], pad[3], dil[3];
  int kvol;
  int subm;
};

static const unsigned int HASH_EMPTY = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned int hash_u32(unsigned int k) {
  k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
  return k;
}

__host__ __device__ __forceinline__ unsigned int lin_key(int b, int z, int y, int x, const int* shape) {
  return (unsigned int)(((b * shape[0] + z) * shape[1] + y) * shape[2] + x);
}

__global__ void __launch_bounds__(256) hash_insert_kernel(const int* __restrict__ indices, int v, const int* v_dev,
                                                          ConvGeom g, unsigned int* __restrict__ keys,
                                                          int* __restrict__ vals, unsigned int mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = v_dev ? *v_dev : v;
  if (i >= n) return;
  const int* c = indices + (size_t)i * 4;
  unsigned int key = lin_key(c[0], c[1], c[2], c[3], g.in_shape);
  unsigned int slot = hash_u32(key) & mask;
  while (true) {
    unsigned int prev = atomicCAS(&keys[slot], HASH_EMPTY, key);
    if (prev == HASH_EMPTY || prev == key) { vals[slot] = i; return; }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ int hash_lookup(const unsigned int* __restrict__ keys, const int* __restrict__ vals,
                                           unsigned int mask, unsigned int key) {
  unsigned int slot = hash_u32(key) & mask;
  while (true) {
    unsigned int k = keys[slot];
    if (k == key) return vals[slot];
    if (k == HASH_EMPTY) return -1;
    slot = (slot + 1) & mask;
  }
}

extern "C" int isf_hash_capacity(int n) {
  long long c = 1024;
  while (c < 2ll * n) c <<= 1;
  return (int)c;
}

static int fill_geom(ConvGeom* g, const int* in_shape, const int* ksize, const int* stride, const int* pad,
                     const int* dil, int subm) {
  g->kvol = 1;
  for (int d = 0; d < 3; ++d) {
    g->in_shape[d] = in_shape[d];
    g->ksize[d] = ksize[d];
    g->dil[d] = dil ? dil[d] : 1;
    g->stride[d] = subm ? 1 : stride[d];
    g->pad[d] = subm ? (ksize[d] / 2) * g->dil[d] : pad[d];
    g->out_shape[d] = subm ? in_shape[d]
                           : (in_shape[d] + 2 * g->pad[d] - g->dil[d] * (ksize[d] - 1) - 1) / g->stride[d] + 1;
    g->kvol *= ksize[d];
  }
  g->subm = subm;
  return 0;
}

// table: [capacity] uint32 keys followed by [capacity] int32 values
extern "C" int isf_hash_build(const int* indices, int v, const int* v_dev, const int* spatial_shape, int batch,
                              void* table, int capacity, void* stream) {
  ISF_CHECK_ARG(v >= 0 && spatial_shape && table && capacity >= 1024 && (capacity & (capacity - 1)) == 0);
  ISF_CHECK_ARG(capacity >= 2 * v || v_dev);
  long long cells = (long long)batch * spatial_shape[0] * spatial_shape[1] * spatial_shape[2];
  if (cells >= 0xFFFFFFFFll) { isf_set_error("spconv: B*D*H*W = %lld does not fit a 32-bit key", cells); return ISF_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* keys = (unsigned int*)table;
  int* vals = (int*)(keys + capacity);
  ISF_CUDA(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned int) * capacity, st));
  if (v == 0) return ISF_OK;
  ISF_CHECK_ARG(indices);
  ConvGeom g{};
  for (int d = 0; d < 3; ++d) g.in_shape[d] = spatial_shape[d];
  hash_insert_kernel<<<isf_cdiv(v, 256), 256, 0, st>>>(indices, v, v_dev, g, keys, vals, (unsigned)capacity - 1);
  isf_count_launch(2);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// strided conv: output sites
// ------------------------------------------------------------------------------------------------
// every (input site, kernel offset) proposes an output site; the VALID proposals (<= ceil(k/s)^3 per input) are
// appended to a compact list (warp-aggregated counter), radix-sorted and made unique
__global__ void __launch_bounds__(256) candidate_append_kernel(const int* __restrict__ indices, int v, const int* v_dev,
                                                               ConvGeom g, unsigned int* __restrict__ list, int list_cap,
                                                               int* __restrict__ list_count) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)v * g.kvol;
  bool won = false;
  unsigned int key = HASH_EMPTY;
  if (t < total) {
    int i = (int)(t / g.kvol), k = (int)(t % g.kvol);
    if (!(v_dev && i >= *v_dev)) {
      int kk[3];
      kk[2] = k % g.ksize[2]; kk[1] = (k / g.ksize[2]) % g.ksize[1]; kk[0] = k / (g.ksize[2] * g.ksize[1]);
      const int* c = indices + (size_t)i * 4;
      int o[3];
      bool ok = true;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        int num = c[1 + d] + g.pad[d] - kk[d] * g.dil[d];
        int q = num / g.stride[d];
        ok &= (num >= 0) && (q * g.stride[d] == num) && (q < g.out_shape[d]);
        o[d] = q;
      }
      if (ok) { key = lin_key(c[0], o[0], o[1], o[2], g.out_shape); won = true; }
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, won);
  if (ballot) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == (__ffs(ballot) - 1)) base = atomicAdd(list_count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
    if (won) {
      const int pos = base + __popc(ballot & ((1u << lane) - 1));
      if (pos < list_cap) list[pos] = key;
    }
  }
}

__global__ void __launch_bounds__(256) decode_out_kernel(const unsigned int* __restrict__ sorted,
                                                         const int* __restrict__ list_count, ConvGeom g, int cap,
                                                         int* __restrict__ out_indices, int* __restrict__ num_out) {
  int n = *list_count;                                         // unique keys incl. the all-ones filler (sorts last)
  if (n > 0 && sorted[n - 1] == HASH_EMPTY) n -= 1;
  if (n > cap) n = cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) *num_out = n;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    unsigned int k = sorted[j];
    int x = k % g.out_shape[2]; k /= g.out_shape[2];
    int y = k % g.out_shape[1]; k /= g.out_shape[1];
    int z = k % g.out_shape[0]; k /= g.out_shape[0];
    int* o = out_indices + (size_t)j * 4;
    o[0] = (int)k; o[1] = z; o[2] = y; o[3] = x;
  }
}

static size_t outidx_cub_bytes(long long n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (unsigned int*)nullptr, (unsigned int*)nullptr, n);
  cub::DeviceSelect::Unique(nullptr, b, (unsigned int*)nullptr, (unsigned int*)nullptr, (int*)nullptr, n);
  return a > b ? a : b;
}

static int conv_reach(const int* ksize, const int* stride) {
  int r = 1;
  for (int d = 0; d < 3; ++d) r *= (ksize[d] + stride[d] - 1) / stride[d];
  return r;
}

static void outidx_carve(IsfArena& A, long long list_cap, unsigned int** list, unsigned int** sorted, unsigned int** uniq,
                         int** count, int** num_sel, void** cub_tmp, size_t* cub_bytes) {
  *list = A.take<unsigned int>(list_cap);
  *sorted = A.take<unsigned int>(list_cap);
  *uniq = A.take<unsigned int>(list_cap);
  *count = A.take<int>(1);
  *num_sel = A.take<int>(1);
  *cub_bytes = outidx_cub_bytes(list_cap);
  *cub_tmp = A.take<char>(*cub_bytes);
}

// workspace for isf_spconv_out_indices: the candidate list holds up to reach = prod ceil(k/s) proposals per input
// site; `reach` is passed by the caller as `kvol`'s companion through out_cap >= 0 (kept for API stability)
extern "C" int isf_spconv_out_indices_workspace_size(int v_in, int kvol, int out_cap, size_t* bytes) {
  ISF_CHECK_ARG(v_in >= 0 && kvol >= 1 && out_cap >= 0 && bytes);
  IsfArena A(nullptr, 0);
  unsigned int *a, *b, *c; int *d, *e; void* f; size_t g;
  long long cap = (long long)(v_in > 0 ? v_in : 1) * (kvol < 8 ? kvol : 8);
  outidx_carve(A, cap, &a, &b, &c, &d, &e, &f, &g);
  *bytes = A.off + 1024;
  return ISF_OK;
}

extern "C" int isf_spconv_out_shape(const int* in_shape, const int* ksize, const int* stride, const int* padding,
                                    const int* dilation, int* out_shape) {
  ISF_CHECK_ARG(in_shape && ksize && stride && padding && out_shape);
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, 0);
  for (int d = 0; d < 3; ++d) out_shape[d] = g.out_shape[d];
  return ISF_OK;
}

extern "C" int isf_spconv_out_indices(const int* in_indices, int v_in, const int* v_in_dev, const int* in_shape, int batch, const int* ksize,
                                      const int* stride, const int* padding, const int* dilation, int* out_indices,
                                      int out_cap, int* num_out_dev, void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(v_in >= 0 && in_shape && ksize && stride && padding && num_out_dev && out_cap >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (v_in == 0 || out_cap == 0) { ISF_CUDA(cudaMemsetAsync(num_out_dev, 0, sizeof(int), st)); return ISF_OK; }
  ISF_CHECK_ARG(in_indices && out_indices && workspace);
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, 0);
  long long cells = (long long)batch * g.out_shape[0] * g.out_shape[1] * g.out_shape[2];
  ISF_CHECK_ARG(cells > 0 && cells < 0xFFFFFFFFll);
  long long n = (long long)v_in * g.kvol;
  ISF_CHECK_ARG(n < 0x7fffffffll);
  int reach = conv_reach(ksize, stride);
  if (reach > g.kvol) reach = g.kvol;
  ISF_CHECK_ARG(reach <= 8);
  const long long list_cap = (long long)v_in * (g.kvol < 8 ? g.kvol : 8);
  IsfArena A(workspace, ws_bytes);
  unsigned int *list, *sorted, *uniq; int *count, *num_sel; void* cub_tmp; size_t cub_bytes;
  outidx_carve(A, list_cap, &list, &sorted, &uniq, &count, &num_sel, &cub_tmp, &cub_bytes);
  if (!A.ok) { isf_set_error("spconv_out_indices: workspace too small"); return ISF_ERR_WORKSPACE; }
  // only the first v_in*reach slots can be written; the filler (all ones) sorts last
  const long long n_sort = (long long)v_in * reach;
  ISF_CUDA(cudaMemsetAsync(list, 0xFF, sizeof(unsigned int) * n_sort, st));
  ISF_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
  candidate_append_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in_indices, v_in, v_in_dev, g, list, (int)n_sort, count);
  int bits = 1;
  while ((1ll << bits) < cells) ++bits;
  int end_bit = bits + 1 > 32 ? 32 : bits + 1;   // keeps the all-ones filler last
  size_t tmp = cub_bytes;
  ISF_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, tmp, list, sorted, (int)n_sort, 0, end_bit, st));
  tmp = cub_bytes;
  ISF_CUDA(cub::DeviceSelect::Unique(cub_tmp, tmp, sorted, uniq, num_sel, (int)n_sort, st));
  decode_out_kernel<<<ISF_NUM_SMS * 2, 256, 0, st>>>(uniq, num_sel, g, out_cap, out_indices, num_out_dev);
  isf_count_launch(8);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// neighbour table
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nbr_kernel(const int* __restrict__ out_indices, int v_out, const int* v_out_dev,
                                                  ConvGeom g, const unsigned int* __restrict__ keys,
                                                  const int* __restrict__ vals, unsigned int mask,
                                                  int* __restrict__ nbr, long long nbr_stride) {
  const int n = v_out_dev ? *v_out_dev : v_out;
  const long long total = (long long)n * g.kvol;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    // consecutive threads -> consecutive rows of one tap: coalesced nbr writes, coalesced index reads
    int k = (int)(t / n), j = (int)(t % n);
    int kk[3];
    kk[2] = k % g.ksize[2]; kk[1] = (k / g.ksize[2]) % g.ksize[1]; kk[0] = k / (g.ksize[2] * g.ksize[1]);
    const int* o = out_indices + (size_t)j * 4;
    int ip[3];
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ip[d] = o[1 + d] * g.stride[d] - g.pad[d] + kk[d] * g.dil[d];
      ok &= ip[d] >= 0 && ip[d] < g.in_shape[d];
    }
    int r = -1;
    if (ok) r = hash_lookup(keys, vals, mask, lin_key(o[0], ip[0], ip[1], ip[2], g.in_shape));
    nbr[(long long)k * nbr_stride + j] = r;
  }
}

extern "C" int isf_spconv_neighbors(const int* out_indices, int v_out, const int* v_out_dev, const void* in_table,
                                    int in_capacity, const int* in_shape, const int* ksize, const int* stride,
                                    const int* padding, const int* dilation, int subm, int* nbr, long long nbr_stride,
                                    void* stream) {
  ISF_CHECK_ARG(v_out >= 0 && in_shape && ksize && in_table && in_capacity >= 1024 && nbr_stride >= v_out);
  ISF_CHECK_ARG(subm || (stride && padding));
  if (v_out == 0) return ISF_OK;
  ISF_CHECK_ARG(out_indices && nbr);
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, subm);
  const unsigned int* keys = (const unsigned int*)in_table;
  const int* vals = (const int*)(keys + in_capacity);
  long long total = (long long)v_out * g.kvol;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  nbr_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_indices, v_out, v_out_dev, g, keys, vals,
                                                       (unsigned)in_capacity - 1, nbr, nbr_stride);
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// bitmap-rank rulebook (rows in ascending linear-key order, which is how this library produces every tensor):
// a coordinate set is one bit per cell + an exclusive prefix of the word popcounts, so
//   row(coordinate) = prefix[key >> 5] + popc(bits[key >> 5] & ((1 << (key & 31)) - 1))          (O(1), no probing)
// and the sorted-unique output sites of a strided conv are simply the set bits in order (no sort at all).
// ------------------------------------------------------------------------------------------------
struct PopcOp {
  __host__ __device__ int operator()(const unsigned int& w) const {
#ifdef __CUDA_ARCH__
    return __popc(w);
#else
    return __builtin_popcount(w);
#endif
  }
};

__global__ void __launch_bounds__(256) bitmap_set_kernel(const int* __restrict__ indices, int v, const int* v_dev, ConvGeom g,
                                                         unsigned int* __restrict__ bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = v_dev ? *v_dev : v;
  if (i >= n) return;
  const int* c = indices + (size_t)i * 4;
  unsigned int key = lin_key(c[0], c[1], c[2], c[3], g.in_shape);
  atomicOr(&bits[key >> 5], 1u << (key & 31));
}

// one thread per input site: per axis the (at most ceil(k/s)) output coordinates it reaches, then their product
__global__ void __launch_bounds__(256) bitmap_candidates_kernel(const int* __restrict__ indices, int v, const int* v_dev,
                                                                ConvGeom g, unsigned int* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = v_dev ? *v_dev : v;
  if (i >= n) return;
  const int* c = indices + (size_t)i * 4;
  int cand[3][4], cnt[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    cnt[d] = 0;
    for (int k = 0; k < g.ksize[d]; ++k) {
      const int num = c[1 + d] + g.pad[d] - k * g.dil[d];
      const int q = num / g.stride[d];
      if (num >= 0 && q * g.stride[d] == num && q < g.out_shape[d] && cnt[d] < 4) cand[d][cnt[d]++] = q;
    }
  }
  for (int a = 0; a < cnt[0]; ++a)
    for (int b = 0; b < cnt[1]; ++b)
      for (int e = 0; e < cnt[2]; ++e) {
        const unsigned int key = lin_key(c[0], cand[0][a], cand[1][b], cand[2][e], g.out_shape);
        atomicOr(&bits[key >> 5], 1u << (key & 31));
      }
}

// Sites beyond the output capacity are dropped from the BITMAP too (their bits are cleared): the bitmap is the next layer's
// coordinate -> row map, and a set bit whose rank is >= cap would send that layer's gather past the [cap, C] feature buffer.
__global__ void __launch_bounds__(256) bitmap_enumerate_kernel(unsigned int* __restrict__ bits,
                                                               const int* __restrict__ prefix, long long words, ConvGeom g,
                                                               int cap, int* __restrict__ out_indices, int* __restrict__ num_out) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
    const unsigned int orig = bits[w];
    unsigned int word = orig, keep = orig;
    int row = prefix[w];
    if (w == words - 1) { int tot = row + __popc(word); *num_out = tot < cap ? tot : cap; }
    while (word) {
      const int b = __ffs(word) - 1;
      word &= word - 1;
      if (row < cap) {
        unsigned int k = (unsigned int)(w * 32 + b);
        int x = k % g.out_shape[2]; k /= g.out_shape[2];
        int y = k % g.out_shape[1]; k /= g.out_shape[1];
        int z = k % g.out_shape[0]; k /= g.out_shape[0];
        int* o = out_indices + (size_t)row * 4;
        o[0] = (int)k; o[1] = z; o[2] = y; o[3] = x;
      } else {
        keep &= ~(1u << b);
      }
      ++row;
    }
    if (keep != orig) bits[w] = keep;
  }
}

__global__ void __launch_bounds__(256) nbr_bitmap_kernel(const int* __restrict__ out_indices, int v_out, const int* v_out_dev,
                                                         ConvGeom g, const unsigned int* __restrict__ bits,
                                                         const int* __restrict__ prefix, int* __restrict__ nbr,
                                                         long long nbr_stride) {
  const int n = v_out_dev ? *v_out_dev : v_out;
  const long long total = (long long)n * g.kvol;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int k = (int)(t / n), j = (int)(t % n);
    int kk[3];
    kk[2] = k % g.ksize[2]; kk[1] = (k / g.ksize[2]) % g.ksize[1]; kk[0] = k / (g.ksize[2] * g.ksize[1]);
    const int* o = out_indices + (size_t)j * 4;
    int ip[3];
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ip[d] = o[1 + d] * g.stride[d] - g.pad[d] + kk[d] * g.dil[d];
      ok &= ip[d] >= 0 && ip[d] < g.in_shape[d];
    }
    int r = -1;
    if (ok) {
      const unsigned int key = lin_key(o[0], ip[0], ip[1], ip[2], g.in_shape);
      const unsigned int word = bits[key >> 5], b = key & 31;
      if ((word >> b) & 1u) r = prefix[key >> 5] + __popc(word & ((1u << b) - 1u));
    }
    nbr[(long long)k * nbr_stride + j] = r;
  }
}

// Row-per-thread variant (opt-in, isf_set_option("nbr_rows", 1)): one thread owns an output row and walks the
// kernel offsets itself.  The kW taps of one (kz, ky) are adjacent cells of the input map, i.e. adjacent bits of (almost
// always) one bitmap word, so the thread loads a word + its prefix once per (kz, ky) instead of once per tap: 9-18
// dependent loads per row instead of 54; the 27 stores of a warp stay coalesced (consecutive rows).  Same results.
struct TapBits { unsigned char pos[32]; };

__host__ __device__ __forceinline__ int isf_popc(unsigned int x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

static void tile_order_bits_fwd(const int* ksize, int taps, unsigned char* pos);   // = tile_order_bits (defined below)

// sort key of an occupancy mask: the mask read as a binary-reflected Gray code, i.e. its rank along the Gray sequence.
// Neighbouring keys then differ in ONE kernel offset, so the small mask groups that end up sharing a 128-row tile have
// a smaller union than under the plain numeric order (-1.5 ... -4 % issued slots on the bench frame, tile_occupancy.py).
__host__ __device__ __forceinline__ unsigned int tile_order_rank(unsigned int mask) {
  unsigned int b = mask;
  b ^= b >> 1; b ^= b >> 2; b ^= b >> 4; b ^= b >> 8; b ^= b >> 16;
  return b;
}

// every kernel offset of the output site (b, z, y, x): MASK_ONLY = false writes the input row (or -1) of offset k to
// dst[k * dst_stride]; MASK_ONLY = true only tests the bitmap (no prefix loads, no stores) and returns the occupancy mask
// with offset k at bit pos[k].  The kW offsets of one (kz, ky) are adjacent cells = adjacent bits: one word (+ prefix)
// load serves them unless they straddle a 32-cell boundary.
template <bool MASK_ONLY>
__host__ __device__ __forceinline__ unsigned int nbr_row_walk(int ob, int oz, int oy, int ox, const ConvGeom& g,
                                                              const unsigned int* __restrict__ bits,
                                                              const int* __restrict__ prefix, const unsigned char* pos,
                                                              int* __restrict__ dst, long long dst_stride) {
  const int x0 = ox * g.stride[2] - g.pad[2];
  unsigned int mask = 0;
  int k = 0;
  for (int kz = 0; kz < g.ksize[0]; ++kz) {
    const int iz = oz * g.stride[0] - g.pad[0] + kz * g.dil[0];
    for (int ky = 0; ky < g.ksize[1]; ++ky) {
      const int iy = oy * g.stride[1] - g.pad[1] + ky * g.dil[1];
      const bool row_ok = iz >= 0 && iz < g.in_shape[0] && iy >= 0 && iy < g.in_shape[1];
      const unsigned int key_row = row_ok ? lin_key(ob, iz, iy, 0, g.in_shape) : 0u;
      unsigned int cached_w = 0xFFFFFFFFu, word = 0;
      int pre = 0;
      for (int kx = 0; kx < g.ksize[2]; ++kx, ++k) {
        const int ix = x0 + kx * g.dil[2];
        int r = -1;
        if (row_ok && ix >= 0 && ix < g.in_shape[2]) {
          const unsigned int key = key_row + (unsigned int)ix;
          const unsigned int w = key >> 5, b = key & 31;
          if (w != cached_w) {
            cached_w = w;
            word = bits[w];
            if (!MASK_ONLY) pre = (word != 0u) ? prefix[w] : 0;
          }
          if ((word >> b) & 1u) {
            if (MASK_ONLY) mask |= 1u << pos[k];
            else r = pre + isf_popc(word & ((1u << b) - 1u));
          }
        }
        if (!MASK_ONLY) dst[(long long)k * dst_stride] = r;
      }
    }
  }
  return mask;
}

// MODE 0: nbr[k][j] for row j.  MODE 1: only the tile-order sort key of row j (occupancy mask of its kernel offsets, no
// prefix loads, no table writes) -> keys[j], rows[j] = j.  MODE 2: slot i computes storage row perm[i] and writes
// nbr[k][i] + row_perm[i] (-1 beyond the valid rows): the tile-ordered table without a detour through the plain one.
template <int MODE>
__global__ void __launch_bounds__(256) nbr_bitmap_rows_kernel(const int* __restrict__ out_indices, int v_out,
                                                              const int* v_out_dev, ConvGeom g,
                                                              const unsigned int* __restrict__ bits,
                                                              const int* __restrict__ prefix, int* __restrict__ nbr,
                                                              long long nbr_stride, const int* __restrict__ perm,
                                                              int* __restrict__ row_perm, const __grid_constant__ TapBits tb, int chunk_shift,
                                                              int invalid_bit, unsigned long long* __restrict__ keys,
                                                              int* __restrict__ rows) {
  const int n = v_out_dev ? min(*v_out_dev, v_out) : v_out;
  const int limit = MODE == 0 ? n : (MODE == 1 ? v_out : (int)nbr_stride);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < limit; i += gridDim.x * blockDim.x) {
    int j = i;
    if (MODE == 1 && i >= n) {                  // capacity rows beyond the device-side count sort last
      keys[i] = 1ull << invalid_bit;
      rows[i] = i;
      continue;
    }
    if (MODE == 2) {
      j = i < n ? perm[i] : -1;
      row_perm[i] = j;
      if (j < 0) {
        for (int k = 0; k < g.kvol; ++k) nbr[(long long)k * nbr_stride + i] = -1;
        continue;
      }
    }
    const int4 o = *(const int4*)(out_indices + (size_t)j * 4);      // (b, z, y, x)
    const unsigned int mask = nbr_row_walk<MODE == 1>(o.x, o.y, o.z, o.w, g, bits, prefix, tb.pos, MODE == 1 ? nullptr : nbr + i,
                                                      nbr_stride);
    if (MODE == 1) {
      keys[i] = ((unsigned long long)(chunk_shift < 31 ? (i >> chunk_shift) : 0) << g.kvol) | tile_order_rank(mask);
      rows[i] = i;
    }
  }
}

// host-only: nbr_row_walk on HOST memory (the same source the kernels compile) — lets the CPU test suite check the
// device logic of the row-per-thread kernels against the oracle rulebook without a GPU.  nbr [kvol, nbr_stride] and
// masks [v_out] are host arrays; bits / prefix = the coordinate bitmap of the INPUT tensor in host memory.
extern "C" int isf_debug_nbr_rows_host(const int* out_indices, int v_out, const unsigned int* in_bits, const int* in_prefix,
                                       const int* in_shape, const int* ksize, const int* stride, const int* padding,
                                       const int* dilation, int subm, int* nbr, long long nbr_stride, unsigned int* masks) {
  ISF_CHECK_ARG(out_indices && in_bits && in_prefix && in_shape && ksize && nbr && masks && nbr_stride >= v_out && v_out >= 0);
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, subm);
  ISF_CHECK_ARG(g.kvol >= 1 && g.kvol <= 27);
  TapBits tb{};
  tile_order_bits_fwd(g.ksize, g.kvol, tb.pos);
  for (int j = 0; j < v_out; ++j) {
    const int* o = out_indices + (size_t)j * 4;
    nbr_row_walk<false>(o[0], o[1], o[2], o[3], g, in_bits, in_prefix, tb.pos, nbr + j, nbr_stride);
    masks[j] = nbr_row_walk<true>(o[0], o[1], o[2], o[3], g, in_bits, in_prefix, tb.pos, nullptr, nbr_stride);
  }
  return ISF_OK;
}

static long long bitmap_words(const int* shape, int batch) {
  long long cells = (long long)batch * shape[0] * shape[1] * shape[2];
  return (cells + 31) / 32;
}

static size_t bitmap_scan_bytes(long long words) {
  size_t a = 0;
  cub::TransformInputIterator<int, PopcOp, const unsigned int*> it((const unsigned int*)nullptr, PopcOp());
  cub::DeviceScan::ExclusiveSum(nullptr, a, it, (int*)nullptr, (int)words);
  return a;
}

static int bitmap_scan(const unsigned int* bits, int* prefix, long long words, void* ws, size_t ws_bytes, cudaStream_t st) {
  cub::TransformInputIterator<int, PopcOp, const unsigned int*> it(bits, PopcOp());
  size_t need = bitmap_scan_bytes(words);
  if (need > ws_bytes) { isf_set_error("bitmap scan: workspace too small"); return ISF_ERR_WORKSPACE; }
  ISF_CUDA(cub::DeviceScan::ExclusiveSum(ws, need, it, prefix, (int)words, st));
  return ISF_OK;
}

// number of 32-bit words of the bitmap (and of the prefix array) of a coordinate space
extern "C" long long isf_bitmap_words(const int* spatial_shape, int batch) {
  if (!spatial_shape) return 0;
  return bitmap_words(spatial_shape, batch);
}

extern "C" int isf_bitmap_workspace_size(long long words, size_t* bytes) {
  ISF_CHECK_ARG(words >= 1 && words < 0x7fffffffll && bytes);
  *bytes = bitmap_scan_bytes(words) + 1024;
  return ISF_OK;
}

// bitmap + prefix of a tensor whose rows are in ascending linear-key order (row == rank)
extern "C" int isf_bitmap_build(const int* indices, int v, const int* v_dev, const int* spatial_shape, int batch,
                                unsigned int* bits, int* prefix, void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(v >= 0 && spatial_shape && bits && prefix && workspace);
  long long words = bitmap_words(spatial_shape, batch);
  ISF_CHECK_ARG(words >= 1 && words < 0x7fffffffll / 32);
  cudaStream_t st = (cudaStream_t)stream;
  ISF_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned int) * words, st));
  if (v > 0) {
    ISF_CHECK_ARG(indices);
    ConvGeom g{};
    for (int d = 0; d < 3; ++d) g.in_shape[d] = spatial_shape[d];
    bitmap_set_kernel<<<isf_cdiv(v, 256), 256, 0, st>>>(indices, v, v_dev, g, bits);
  }
  int rc = bitmap_scan(bits, prefix, words, workspace, ws_bytes, st);
  if (rc) return rc;
  isf_count_launch(4);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// output sites of a strided conv as a bitmap: sets out_bits / out_prefix (the coordinate->row map of the OUTPUT
// tensor, reusable as the input map of the next layer) and writes the sorted-unique out_indices + count
extern "C" int isf_bitmap_out_sites(const int* in_indices, int v_in, const int* v_in_dev, const int* in_shape, int batch,
                                    const int* ksize, const int* stride, const int* padding, const int* dilation,
                                    unsigned int* out_bits, int* out_prefix, int* out_indices, int out_cap,
                                    int* num_out_dev, void* workspace, size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(v_in >= 0 && in_shape && ksize && stride && padding && out_bits && out_prefix && num_out_dev && workspace);
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, 0);
  long long words = bitmap_words(g.out_shape, batch);
  ISF_CHECK_ARG(words >= 1 && words < 0x7fffffffll / 32);
  cudaStream_t st = (cudaStream_t)stream;
  ISF_CUDA(cudaMemsetAsync(out_bits, 0, sizeof(unsigned int) * words, st));
  if (v_in > 0) {
    ISF_CHECK_ARG(in_indices && out_indices);
    for (int d = 0; d < 3; ++d) ISF_CHECK_ARG((ksize[d] + g.stride[d] - 1) / g.stride[d] <= 4);
    bitmap_candidates_kernel<<<isf_cdiv(v_in, 256), 256, 0, st>>>(in_indices, v_in, v_in_dev, g, out_bits);
  }
  int rc = bitmap_scan(out_bits, out_prefix, words, workspace, ws_bytes, st);
  if (rc) return rc;
  int blocks = (int)((words + 255) / 256);
  if (blocks > ISF_NUM_SMS * 8) blocks = ISF_NUM_SMS * 8;
  bitmap_enumerate_kernel<<<blocks, 256, 0, st>>>(out_bits, out_prefix, words, g, out_cap, out_indices, num_out_dev);
  isf_count_launch(5);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

extern "C" int isf_spconv_neighbors_bitmap(const int* out_indices, int v_out, const int* v_out_dev,
                                           const unsigned int* in_bits, const int* in_prefix, const int* in_shape,
                                           const int* ksize, const int* stride, const int* padding, const int* dilation,
                                           int subm, int* nbr, long long nbr_stride, void* stream) {
  ISF_CHECK_ARG(v_out >= 0 && in_shape && ksize && in_bits && in_prefix && nbr_stride >= v_out);
  ISF_CHECK_ARG(subm || (stride && padding));
  if (v_out == 0) return ISF_OK;
  ISF_CHECK_ARG(out_indices && nbr);
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, subm);
  long long total = (long long)v_out * g.kvol;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  if (::g_isf_nbr_rows && ((uintptr_t)out_indices & 15) == 0) {      // the row kernel reads a coordinate as one int4
    int rb = isf_cdiv(v_out, 256);
    if (rb > ISF_NUM_SMS * 16) rb = ISF_NUM_SMS * 16;
    nbr_bitmap_rows_kernel<0><<<rb, 256, 0, (cudaStream_t)stream>>>(out_indices, v_out, v_out_dev, g, in_bits, in_prefix, nbr,
                                                                    nbr_stride, nullptr, nullptr, TapBits{}, 0, 0, nullptr, nullptr);
  } else {
    nbr_bitmap_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_indices, v_out, v_out_dev, g, in_bits, in_prefix, nbr,
                                                                nbr_stride);
  }
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// the convolution itself: implicit GEMM over the neighbour table
// ------------------------------------------------------------------------------------------------
extern "C" int isf_spconv_implicit_gemm(const void* feats_in, int cin, const void* w_packed, int kpad, int cout,
                                        const int* nbr, long long nbr_stride, int taps, int v_out,
                                        const int* v_out_dev, const float* scale, const float* shift,
                                        const void* residual, int ldr, int relu, void* out_f16, int ldo,
                                        void* stream) {
  ISF_CHECK_ARG(v_out >= 0 && cin % 8 == 0 && cin >= 8 && taps >= 1 && kpad % 64 == 0 && kpad >= taps * cin);
  ISF_CHECK_ARG(cout == 32 || cout == 64 || cout == 128 || cout == 256);
  ISF_CHECK_ARG((cin & (cin - 1)) == 0);            // power of two: tap = k >> log2(cin), a piece never straddles taps
  if (v_out == 0) return ISF_OK;
  ISF_CHECK_ARG(feats_in && w_packed && nbr && out_f16 && ldo % 8 == 0);
  tc::GemmParams p{};
  p.A = (const __half*)feats_in; p.W = (const __half*)w_packed; p.nbr = nbr; p.nbr_stride = nbr_stride;
  p.m_dev = v_out_dev; p.m = v_out; p.cin = cin; p.taps = taps;
  p.cin_shift = 0;
  while ((1 << p.cin_shift) < cin) ++p.cin_shift; p.kpad = kpad; p.n_total = cout;
  p.scale = scale; p.shift = shift; p.residual = (const __half*)residual; p.ldr = ldr;
  p.act = relu ? tc::ACT_RELU : tc::ACT_NONE;
  p.out = (__half*)out_f16; p.ldo = ldo;
  p.tile_ctr = isf_take_tile_counter();
  int rc = tc::dispatch_gemm(p, cout, v_out, (cudaStream_t)stream);
  isf_count_launch(1);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// Tile order (opt-in): which rows share a 128-row tile of the implicit GEMM.
//
// The kernel skips a kernel offset for a whole tile when none of the tile's rows has a neighbour there, and
// zero-fills the slot of every row that lacks one otherwise.  With the rows taken in storage order (ascending
// (b,z,y,x)) a tile of a LiDAR frame touches ~25 of the 27 offsets while a row has 4-20 neighbours: 40-75 % of the
// gathered slots and of the issued MMAs are zeros (tests/analysis/tile_occupancy.py).  Spatial re-blocking does not
// change that; grouping rows by their NEIGHBOUR PATTERN does: rows are stably sorted by their 27-bit occupancy
// mask (rarest kernel offsets = most significant bits), optionally inside chunks of 2^chunk_log2 consecutive rows
// (keeps the gathers of a tile inside a bounded window of the feature matrix).  Storage order is untouched: the
// kernel reads the permuted neighbour table and scatters / gathers its epilogue rows through row_perm
// (GemmParams::out_rows / res_rows), so results are bit-identical to the unordered launch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_order_keys_kernel(const int* __restrict__ nbr, long long nbr_stride, int taps,
                                                              int v, const int* __restrict__ v_dev, int chunk_shift,
                                                              int invalid_bit, const __grid_constant__ TapBits tb,
                                                              unsigned long long* __restrict__ keys, int* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v) return;
  const int n = v_dev ? min(*v_dev, v) : v;
  unsigned long long key;
  if (i < n) {
    unsigned int mask = 0;
    for (int k = 0; k < taps; ++k)
      if (nbr[(long long)k * nbr_stride + i] >= 0) mask |= 1u << tb.pos[k];
    key = ((unsigned long long)(chunk_shift < 31 ? (i >> chunk_shift) : 0) << taps) | tile_order_rank(mask);
  } else {
    key = 1ull << invalid_bit;          // capacity rows beyond the device-side count sort last
  }
  keys[i] = key;
  rows[i] = i;
}

__global__ void __launch_bounds__(256) tile_order_apply_kernel(const int* __restrict__ nbr, long long nbr_stride, int taps,
                                                               int v, const int* __restrict__ v_dev,
                                                               const int* __restrict__ rows_sorted, int* __restrict__ nbr_out,
                                                               int* __restrict__ row_perm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbr_stride) return;
  const int n = v_dev ? min(*v_dev, v) : v;
  const int r = i < n ? rows_sorted[i] : -1;
  row_perm[i] = r;
  for (int k = 0; k < taps; ++k)
    nbr_out[(long long)k * nbr_stride + i] = r >= 0 ? nbr[(long long)k * nbr_stride + r] : -1;
}

// significance of the taps in the sort key: rank by (moves in z, in-plane L1 distance, moves in y), ties by tap index;
// pos[k] = bit of tap k (taps-1 = most significant = sorts first)
static void tile_order_bits(const int* ksize, int taps, unsigned char* pos) {
  int order[27], score[27];
  for (int k = 0; k < taps; ++k) {
    const int kx = k % ksize[2], ky = (k / ksize[2]) % ksize[1], kz = k / (ksize[2] * ksize[1]);
    const int dz = abs(kz - ksize[0] / 2), dy = abs(ky - ksize[1] / 2), dx = abs(kx - ksize[2] / 2);
    score[k] = (dz * 16 + (dy + dx)) * 4 + dy;
    order[k] = k;
  }
  for (int a = 1; a < taps; ++a)                       // insertion sort, descending score, stable
    for (int b = a; b > 0 && score[order[b]] > score[order[b - 1]]; --b) { int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t; }
  for (int rank = 0; rank < taps; ++rank) pos[order[rank]] = (unsigned char)(taps - 1 - rank);
}

static void tile_order_bits_fwd(const int* ksize, int taps, unsigned char* pos) { tile_order_bits(ksize, taps, pos); }

// host-only: the bit position of every tap in the sort key (for tests and for CPU models of the order)
extern "C" int isf_spconv_tile_order_bits(const int* ksize, int taps, unsigned char* pos) {
  ISF_CHECK_ARG(ksize && pos && taps >= 1 && taps <= 27 && ksize[0] * ksize[1] * ksize[2] == taps);
  tile_order_bits(ksize, taps, pos);
  return ISF_OK;
}

static void tile_order_carve(IsfArena& A, int v, unsigned long long** k0, unsigned long long** k1, int** r0, int** r1,
                             char** cub_tmp, size_t* cub_bytes) {
  *k0 = A.take<unsigned long long>(v);
  *k1 = A.take<unsigned long long>(v);
  *r0 = A.take<int>(v);
  *r1 = A.take<int>(v);
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                  (int*)nullptr, v);
  *cub_bytes = b;
  *cub_tmp = A.take<char>(b);
}

extern "C" int isf_spconv_tile_order_workspace_size(int v, size_t* bytes) {
  ISF_CHECK_ARG(v >= 0 && bytes);
  IsfArena A(nullptr, 0);
  unsigned long long *k0, *k1; int *r0, *r1; char* tmp; size_t cb;
  tile_order_carve(A, v > 0 ? v : 1, &k0, &k1, &r0, &r1, &tmp, &cb);
  *bytes = A.off;
  return ISF_OK;
}

// nbr [taps, nbr_stride] -> nbr_out [taps, nbr_stride] (rows regrouped) + row_perm [nbr_stride] (tile slot -> storage
// row, -1 beyond the valid rows).  ksize = the kernel's (kD,kH,kW) (taps = kD*kH*kW): it fixes the significance of
// the mask bits (offsets that move in z, then the far in-plane ones, sort first).  chunk_log2 >= 31 = one global sort.
extern "C" int isf_spconv_tile_order(const int* nbr, long long nbr_stride, int taps, const int* ksize, int v,
                                     const int* v_dev, int chunk_log2, int* nbr_out, int* row_perm, void* workspace,
                                     size_t ws_bytes, void* stream) {
  ISF_CHECK_ARG(nbr && nbr_out && row_perm && ksize && nbr != nbr_out);
  ISF_CHECK_ARG(taps >= 1 && taps <= 27 && ksize[0] * ksize[1] * ksize[2] == taps && v >= 0 && nbr_stride >= v);
  ISF_CHECK_ARG(chunk_log2 >= 7);                      // a chunk holds at least one 128-row tile
  cudaStream_t st = (cudaStream_t)stream;
  const int vv = v > 0 ? v : 1;
  IsfArena A(workspace, ws_bytes);
  unsigned long long *k0, *k1; int *r0, *r1; char* tmp; size_t cb;
  tile_order_carve(A, vv, &k0, &k1, &r0, &r1, &tmp, &cb);
  if (!A.ok || !workspace) { isf_set_error("isf_spconv_tile_order: workspace too small"); return ISF_ERR_WORKSPACE; }
  TapBits tb{};
  tile_order_bits(ksize, taps, tb.pos);
  int chunk_bits = 0;
  if (chunk_log2 < 31) while (((long long)(vv - 1) >> chunk_log2) >> chunk_bits) ++chunk_bits;
  const int invalid_bit = taps + chunk_bits;
  if (v > 0) {
    tile_order_keys_kernel<<<isf_cdiv(v, 256), 256, 0, st>>>(nbr, nbr_stride, taps, v, v_dev, chunk_log2, invalid_bit, tb, k0, r0);
    ISF_LAUNCH_CHECK();
    // isf_set_option("tile_order_drop_bits", d): the d least significant mask bits (near in-plane offsets) are left
    // out of the sort -> fewer radix passes; rows that differ only there keep their storage order (stable sort)
    int begin_bit = ::g_isf_tile_order_drop_bits;
    if (begin_bit > taps - 1) begin_bit = taps - 1;
    ISF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, cb, k0, k1, r0, r1, v, begin_bit, invalid_bit + 1, st));
  }
  tile_order_apply_kernel<<<(int)((nbr_stride + 255) / 256), 256, 0, st>>>(nbr, nbr_stride, taps, v, v_dev, r1, nbr_out, row_perm);
  ISF_LAUNCH_CHECK();
  isf_count_launch(v > 0 ? 4 + (invalid_bit + 1 - (::g_isf_tile_order_drop_bits > taps - 1 ? taps - 1 : ::g_isf_tile_order_drop_bits) + 7) / 8 : 1);   // keys + CUB (histogram, scan, one onesweep per 8 key bits) + apply
  return ISF_OK;
}

// isf_spconv_neighbors_bitmap + isf_spconv_tile_order in one: occupancy masks straight from the bitmap (mode-1 row
// kernel) -> sort -> the table is WRITTEN in tile order (mode-2 row kernel).  The plain table is never materialised
// and nothing is gathered back: two row-kernel passes + one radix sort.  Same nbr_out / row_perm as the two-call path.
extern "C" int isf_spconv_neighbors_bitmap_ordered(const int* out_indices, int v_out, const int* v_out_dev,
                                                   const unsigned int* in_bits, const int* in_prefix, const int* in_shape,
                                                   const int* ksize, const int* stride, const int* padding,
                                                   const int* dilation, int subm, int chunk_log2, int* nbr_out,
                                                   long long nbr_stride, int* row_perm, void* workspace, size_t ws_bytes,
                                                   void* stream) {
  ISF_CHECK_ARG(v_out >= 0 && in_shape && ksize && in_bits && in_prefix && nbr_stride >= v_out && nbr_out && row_perm);
  ISF_CHECK_ARG(subm || (stride && padding));
  ISF_CHECK_ARG(chunk_log2 >= 7 && (v_out == 0 || (out_indices && ((uintptr_t)out_indices & 15) == 0)));
  ConvGeom g;
  fill_geom(&g, in_shape, ksize, stride, padding, dilation, subm);
  ISF_CHECK_ARG(g.kvol >= 1 && g.kvol <= 27);
  cudaStream_t st = (cudaStream_t)stream;
  const int vv = v_out > 0 ? v_out : 1;
  IsfArena A(workspace, ws_bytes);
  unsigned long long *k0, *k1; int *r0, *r1; char* tmp; size_t cb;
  tile_order_carve(A, vv, &k0, &k1, &r0, &r1, &tmp, &cb);
  if (!A.ok || !workspace) { isf_set_error("isf_spconv_neighbors_bitmap_ordered: workspace too small"); return ISF_ERR_WORKSPACE; }
  TapBits tb{};
  tile_order_bits(g.ksize, g.kvol, tb.pos);
  int chunk_bits = 0;
  if (chunk_log2 < 31) while (((long long)(vv - 1) >> chunk_log2) >> chunk_bits) ++chunk_bits;
  const int invalid_bit = g.kvol + chunk_bits;
  int begin_bit = ::g_isf_tile_order_drop_bits;
  if (begin_bit > g.kvol - 1) begin_bit = g.kvol - 1;
  int launches = 1;
  if (v_out > 0) {
    int rb = isf_cdiv(v_out, 256);
    if (rb > ISF_NUM_SMS * 16) rb = ISF_NUM_SMS * 16;
    nbr_bitmap_rows_kernel<1><<<rb, 256, 0, st>>>(out_indices, v_out, v_out_dev, g, in_bits, in_prefix, nullptr, nbr_stride,
                                                  nullptr, nullptr, tb, chunk_log2, invalid_bit, k0, r0);
    ISF_LAUNCH_CHECK();
    ISF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, cb, k0, k1, r0, r1, v_out, begin_bit, invalid_bit + 1, st));
    launches += 3 + (invalid_bit + 1 - begin_bit + 7) / 8;
  }
  int rb2 = (int)((nbr_stride + 255) / 256);
  if (rb2 > ISF_NUM_SMS * 16) rb2 = ISF_NUM_SMS * 16;
  nbr_bitmap_rows_kernel<2><<<rb2, 256, 0, st>>>(out_indices, v_out, v_out_dev, g, in_bits, in_prefix, nbr_out, nbr_stride, r1,
                                                 row_perm, tb, 0, 0, nullptr, nullptr);
  ISF_LAUNCH_CHECK();
  isf_count_launch(launches);
  return ISF_OK;
}

// implicit GEMM over a tile-ordered neighbour table: tile slot i computes storage row row_perm[i]
extern "C" int isf_spconv_implicit_gemm_ordered(const void* feats_in, int cin, const void* w_packed, int kpad, int cout,
                                                const int* nbr_ordered, long long nbr_stride, int taps, int v_out,
                                                const int* v_out_dev, const int* row_perm, const float* scale,
                                                const float* shift, const void* residual, int ldr, int relu,
                                                void* out_f16, int ldo, void* stream) {
  ISF_CHECK_ARG(v_out >= 0 && cin % 8 == 0 && cin >= 8 && taps >= 1 && kpad % 64 == 0 && kpad >= taps * cin);
  ISF_CHECK_ARG(cout == 32 || cout == 64 || cout == 128 || cout == 256);
  ISF_CHECK_ARG((cin & (cin - 1)) == 0);
  if (v_out == 0) return ISF_OK;
  ISF_CHECK_ARG(feats_in && w_packed && nbr_ordered && row_perm && out_f16 && ldo % 8 == 0);
  tc::GemmParams p{};
  p.A = (const __half*)feats_in; p.W = (const __half*)w_packed; p.nbr = nbr_ordered; p.nbr_stride = nbr_stride;
  p.m_dev = v_out_dev; p.m = v_out; p.cin = cin; p.taps = taps;
  p.cin_shift = 0;
  while ((1 << p.cin_shift) < cin) ++p.cin_shift;
  p.kpad = kpad; p.n_total = cout;
  p.scale = scale; p.shift = shift; p.residual = (const __half*)residual; p.ldr = ldr;
  p.res_rows = residual ? row_perm : nullptr;
  p.out_rows = row_perm;
  p.act = relu ? tc::ACT_RELU : tc::ACT_NONE;
  p.out = (__half*)out_f16; p.ldo = ldo;
  p.tile_ctr = isf_take_tile_counter();
  int rc = tc::dispatch_gemm(p, cout, v_out, (cudaStream_t)stream);
  isf_count_launch(1);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// SparseConvTensor.dense()  (structure.py:49-59)  -> [B, C, D, H, W] fp16, zero filled here
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_dense_kernel(const __half* __restrict__ feats, const int* __restrict__ idx,
                                                       int v, int c, int D, int H, int W, __half* __restrict__ out) {
  long long total = (long long)v * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int j = (int)(t / c), ch = (int)(t % c);
    const int* o = idx + (size_t)j * 4;
    out[((((long long)o[0] * c + ch) * D + o[1]) * H + o[2]) * W + o[3]] = feats[t];
  }
}

extern "C" int isf_sparse_to_dense(const void* feats, const int* indices, int v, int c, int batch, int D, int H, int W,
                                   void* out, void* stream) {
  ISF_CHECK_ARG(v >= 0 && c >= 1 && batch >= 1 && out);
  cudaStream_t st = (cudaStream_t)stream;
  ISF_CUDA(cudaMemsetAsync(out, 0, sizeof(__half) * (size_t)batch * c * D * H * W, st));
  if (v == 0) return ISF_OK;
  long long total = (long long)v * c;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  to_dense_kernel<<<blocks, 256, 0, st>>>((const __half*)feats, indices, v, c, D, H, W, (__half*)out);
  isf_count_launch(2);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// BEV view of dense(): [B, C*D, H, W] with channel c' = c*D + z (sparse_encoder.py:133-136), written as NHWC
// fp16 rows of pitch `ldo` starting at channel `ch_offset` (so it can land inside the fused 768-ch buffer).
__global__ void __launch_bounds__(256) to_bev_nhwc_kernel(const __half* __restrict__ feats, const int* __restrict__ idx,
                                                          int v, const int* v_dev, int c, int D, int H, int W,
                                                          __half* __restrict__ out, int ldo, int ch_offset) {
  if (v_dev) v = *v_dev;
  long long total = (long long)v * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int j = (int)(t / c), ch = (int)(t % c);
    const int* o = idx + (size_t)j * 4;
    out[(((long long)o[0] * H + o[2]) * W + o[3]) * ldo + ch_offset + ch * D + o[1]] = feats[t];
  }
}

extern "C" int isf_sparse_to_bev_nhwc(const void* feats, const int* indices, int v, const int* v_dev, int c, int D, int H,
                                      int W, void* out, int ldo, int ch_offset, void* stream) {
  ISF_CHECK_ARG(v >= 0 && c >= 1 && out);
  if (v == 0) return ISF_OK;
  long long total = (long long)v * c;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ISF_NUM_SMS * 16) blocks = ISF_NUM_SMS * 16;
  to_bev_nhwc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)feats, indices, v, v_dev, c, D, H, W,
                                                               (__half*)out, ldo, ch_offset);
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}
