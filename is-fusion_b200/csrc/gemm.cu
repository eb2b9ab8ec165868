// Dense fp16 GEMM with fused epilogues on tcgen05 (linear layers of the SST / IGF transformers:
// linear0, in_proj, out_proj + residual + LayerNorm, FFN, value_proj ...).
#include "tc_gemm.cuh"

extern "C" void isf_count_launch(int n);

static int pick_bn(int n) {
  if (n == 32 || n == 64 || n == 128 || n == 256) return n;
  if (n % 128 == 0) return 128;
  if (n % 64 == 0) return 64;
  if (n % 32 == 0) return 32;
  return 0;
}

// out[m, n] = epi(A[m, k] . W[n, kpad]^T); see include/isfusion_b200.h
extern "C" int isf_gemm_f16(const void* A, const void* W, int m, int k, int kpad, int n, const float* scale,
                            const float* shift, const void* residual, int ldr, int act, const float* ln_gamma,
                            const float* ln_beta, float ln_eps, void* out_f16, float* out_f32, int ldo,
                            const int* out_rows, const int* m_dev, void* stream) {
  ISF_CHECK_ARG(m >= 0 && k > 0 && k % 8 == 0 && kpad % 64 == 0 && kpad >= k && n > 0);
  if (m == 0) return ISF_OK;
  ISF_CHECK_ARG(A && W && (out_f16 || out_f32));
  ISF_CHECK_ARG(ldo % 8 == 0 && (residual == nullptr || ldr % 8 == 0));
  int bn = pick_bn(n);
  if (!bn) { isf_set_error("isf_gemm_f16: n=%d must be a multiple of 32", n); return ISF_ERR_UNSUPPORTED; }
  if (ln_gamma && bn != n) { isf_set_error("isf_gemm_f16: LayerNorm epilogue needs n in {32,64,128,256}"); return ISF_ERR_UNSUPPORTED; }
  tc::GemmParams p{};
  p.A = (const __half*)A; p.W = (const __half*)W; p.nbr = nullptr; p.nbr_stride = 0; p.m_dev = m_dev; p.m = m;
  p.cin = k; p.cin_shift = 0; p.taps = 1; p.kpad = kpad; p.n_total = n;
  p.scale = scale; p.shift = shift; p.residual = (const __half*)residual; p.ldr = ldr; p.act = act;
  p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.ln_eps = ln_eps;
  p.out = (__half*)out_f16; p.out_f32 = out_f32; p.ldo = ldo; p.out_rows = out_rows;
  int rc = tc::dispatch_gemm(p, bn, m, (cudaStream_t)stream);
  isf_count_launch(1);
  return rc;
}

// bf16 twin (BASELINE configs[2]: "bf16 tensor-core path"): A, W, residual and the 16-bit output are bf16; same kernels, same fp32
// accumulation and epilogue arithmetic (tcgen05 kind::f16 with the BF16 operand formats).
extern "C" int isf_gemm_bf16(const void* A, const void* W, int m, int k, int kpad, int n, const float* scale,
                             const float* shift, const void* residual, int ldr, int act, const float* ln_gamma,
                             const float* ln_beta, float ln_eps, void* out_bf16, float* out_f32, int ldo,
                             const int* out_rows, const int* m_dev, void* stream) {
  IsfBf16Scope scope;
  return isf_gemm_f16(A, W, m, k, kpad, n, scale, shift, residual, ldr, act, ln_gamma, ln_beta, ln_eps, out_bf16, out_f32, ldo,
                      out_rows, m_dev, stream);
}
