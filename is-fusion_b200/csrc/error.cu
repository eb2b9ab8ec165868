#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

extern "C" void isf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* isf_last_error(void) { return g_err; }

extern "C" int isf_version(void) { return 100; }

// Counts kernels launched by this library since the last reset (bench.py "gpu_launches").
static unsigned long long g_launches = 0;
extern "C" void isf_count_launch(int n) { g_launches += (unsigned long long)n; }
extern "C" unsigned long long isf_launch_count(int reset) {
  unsigned long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// ---- runtime options
int g_isf_gemm_cluster = 1;
int g_isf_tile_order_drop_bits = 0;
int g_isf_nbr_rows = 1;
int g_isf_gemm_mt = 3;
static thread_local int* t_isf_tile_ctr = nullptr;
int* isf_take_tile_counter() { int* c = t_isf_tile_ctr; t_isf_tile_ctr = nullptr; return c; }
// see include/isfusion_b200.h
extern "C" int isf_spconv_tile_counter(int* counter2) { t_isf_tile_ctr = counter2; return ISF_OK; }
static thread_local int t_isf_bf16 = 0;
int isf_bf16_mode() { return t_isf_bf16; }
IsfBf16Scope::IsfBf16Scope() { ++t_isf_bf16; }
IsfBf16Scope::~IsfBf16Scope() { --t_isf_bf16; }
int g_isf_gemm_zskip = 0;
int g_isf_gemm_wpc = 0;
int g_isf_gemm_rev = 0;
int g_isf_gemm_pair = 128;
int g_isf_i2s_tc = 0;
int g_isf_gemm_mma2 = 1;
int g_isf_gemm_trim = 0;
int g_isf_gemm_v4_32 = 1;
int g_isf_gemm_v4 = 2;
int g_isf_gemm_ln_fast = 1;
int g_isf_gemm_debug = 0;
extern "C" int isf_set_option(const char* name, int value) {
  if (name && strcmp(name, "gemm_cluster") == 0 && (value == 1 || value == 2)) {
    g_isf_gemm_cluster = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "tile_order_drop_bits") == 0 && value >= 0 && value <= 26) {
    g_isf_tile_order_drop_bits = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_mt") == 0 && value >= 1 && value <= 3) {
    g_isf_gemm_mt = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_ln_fast") == 0 && (value == 0 || value == 1)) {
    g_isf_gemm_ln_fast = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_v4") == 0 && value >= 0 && value <= 2) {
    g_isf_gemm_v4 = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_zskip") == 0 && (value == 0 || value == 1)) {
    g_isf_gemm_zskip = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_pair") == 0 && (value == 0 || value == 1 || value == 64 || value == 128 || value == 256)) {
    g_isf_gemm_pair = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_trim") == 0 && (value == 0 || value == 1)) {
    g_isf_gemm_trim = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_mma2") == 0 && (value == 0 || value == 1)) {
    g_isf_gemm_mma2 = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "i2s_tc") == 0 && (value == 0 || value == 1)) {
    g_isf_i2s_tc = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_wpc") == 0 && value >= 0 && value <= 2) {
    g_isf_gemm_wpc = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_rev") == 0 && (value == 0 || value == 1)) {
    g_isf_gemm_rev = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_v4_32") == 0 && value >= 0 && value <= 2) {
    g_isf_gemm_v4_32 = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "gemm_debug") == 0 && value >= 0 && value <= 7) {
    g_isf_gemm_debug = value;
    return ISF_OK;
  }
  if (name && strcmp(name, "nbr_rows") == 0 && (value == 0 || value == 1)) {
    g_isf_nbr_rows = value;
    return ISF_OK;
  }
  isf_set_error("isf_set_option: unknown option or bad value: %s=%d", name ? name : "(null)", value);
  return ISF_ERR_INVALID;
}
