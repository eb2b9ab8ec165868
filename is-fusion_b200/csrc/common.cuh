// Shared helpers for the isfusion_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define ISF_OK 0
#define ISF_ERR_INVALID 1      // bad argument (null pointer, bad size, unsupported shape)
#define ISF_ERR_WORKSPACE 2    // workspace too small
#define ISF_ERR_CUDA 3         // CUDA runtime error (see isf_last_error)
#define ISF_ERR_UNSUPPORTED 4

extern "C" void isf_set_error(const char* fmt, ...);

#define ISF_CHECK_ARG(cond)                                                       \
  do {                                                                            \
    if (!(cond)) {                                                                \
      isf_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);    \
      return ISF_ERR_INVALID;                                                     \
    }                                                                             \
  } while (0)

#define ISF_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      isf_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                    cudaGetErrorString(_e));                                      \
      return ISF_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define ISF_LAUNCH_CHECK() ISF_CUDA(cudaGetLastError())

static inline size_t isf_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace (no hidden allocation in the library).
struct IsfArena {
  char* base;
  size_t cap;
  size_t off;
  bool ok;
  IsfArena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), ok(true) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = isf_align_up(count * sizeof(T));
    if (base == nullptr) {  // sizing pass
      off += bytes;
      return nullptr;
    }
    if (off + bytes > cap) {
      ok = false;
      return nullptr;
    }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

__host__ __device__ inline int isf_cdiv(int a, int b) { return (a + b - 1) / b; }

// SM count of the current device (148 on B200), cached per device ordinal
static inline int isf_num_sms() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}
#define ISF_NUM_SMS isf_num_sms()

extern int g_isf_gemm_cluster;    // isf_set_option("gemm_cluster", 1|2), defined in error.cu
extern int g_isf_tile_order_drop_bits;   // isf_set_option("tile_order_drop_bits", 0..26): low mask bits left out of the tile-order sort
extern int g_isf_gemm_ln_fast;    // isf_set_option("gemm_ln_fast", 0|1): split-row LayerNorm epilogue (default 1; 0 = the round-1 two-pass epilogue)
extern int g_isf_gemm_v4;         // isf_set_option("gemm_v4", 0|1|2): dedicated-epilogue / double-buffered schedule (tc_gemm4.cuh)
extern int g_isf_gemm_wpc;        // isf_set_option("gemm_wpc", 0|1|2): one / two producer warps per K chunk in the v4 schedule (0 = all eight)
extern int g_isf_gemm_rev;        // isf_set_option("gemm_rev", 0|1): v4 walks the tiles from the last one down (default 1)
extern int g_isf_gemm_v4_32;      // isf_set_option("gemm_v4_32", 0|1|2): 32-column layers on the v4 schedule (1 = 2 CTAs / SM, default; 2 = 3 CTAs / SM)
extern int g_isf_gemm_pair;       // isf_set_option("gemm_pair", 0|1|64|128|256): sparse convs on CTA pairs (tcgen05.mma.cta_group::2, tc_gemm5.cuh);
                                  // 1 = every column tile >= 64, N = only column tiles >= N (default 128)
extern int g_isf_gemm_zskip;      // isf_set_option("gemm_zskip", 0|1): zero-skip of the v4 producers
extern int g_isf_gemm_mt;         // isf_set_option("gemm_mt", 1|2): 128-row sub-tiles per CTA tile of the sparse-conv GEMM
extern int g_isf_gemm_debug;      // isf_set_option("gemm_debug", bits): timing experiments (tc_gemm.cuh GemmParams::debug)
extern int g_isf_gemm_trim;       // isf_set_option("gemm_trim", 0|1): fewest CTAs that keep the number of tile rounds (default 0: +1 % throughput with 3 frames in flight, -3.5 % per launch)
extern int g_isf_gemm_mma2;       // isf_set_option("gemm_mma2", 0|1): two MMA-issuing warps in the 256-row CTA-pair kernels (tc_gemm5.cuh)
extern int g_isf_i2s_tc;          // isf_set_option("i2s_tc", 0|1): instance -> scene attention on the tcgen05 kernel (i2s_tc.cu; default 0: measured 3x slower than the mma.sync kernel)
extern int g_isf_nbr_rows;        // isf_set_option("nbr_rows", 0|1): row-per-thread neighbour-table kernel (opt-in)

// ---- 16-bit element type of the dense tensor-core entry points (BASELINE configs[2] names a bf16 path): the `_bf16` entry points
// (isf_gemm_bf16, isf_conv3x3_nhwc_bf16, isf_window_attention_bf16, isf_cam2bev_bf16) run the SAME kernels with bf16 operand /
// output encoding (tcgen05 kind::f16 takes either format at the same rate; accumulation, epilogue and softmax stay fp32).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
int* isf_take_tile_counter();               // error.cu: the counter handed over by isf_spconv_tile_counter (consumed), or nullptr
int isf_bf16_mode();                       // error.cu: 1 inside a `_bf16` entry point of this thread
struct IsfBf16Scope { IsfBf16Scope(); ~IsfBf16Scope(); };
__device__ __forceinline__ float2 isf_unpack2(unsigned int w, bool bf16) {
  if (bf16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
  return __half22float2(*(const __half2*)&w);
}
__device__ __forceinline__ unsigned int isf_pack2(float x, float y, bool bf16) {
  if (bf16) { const __nv_bfloat162 b = __floats2bfloat162_rn(x, y); return *(const unsigned int*)&b; }
  const __half2 h = __floats2half2_rn(x, y);
  return *(const unsigned int*)&h;
}
