// Input contract, camera side: ImageAug3D (resize -> crop -> flip) + ImageNormalize on the device.
//
// Replaces the per-camera PIL / torchvision work of mmdet3d/datasets/pipelines/transforms_3d.py:25-145
// (img.resize(resize_dims) = Pillow's 8-bit BICUBIC resampler, img.crop(crop), FLIP_LEFT_RIGHT, ToTensor, Normalize):
// ~5 ms of host CPU per 1600x900 camera, 6 cameras per frame — more than the whole fusion neck takes on the GPU.
// Here the raw uint8 frames are copied once (pinned host -> HBM, 4.3 MB per camera) and two kernels do the rest:
//   pass 1 (horizontal): for the input rows the vertical pass needs and the output COLUMNS that survive the crop only,
//           sum_j px[xmin + j] * coef[j] in 32-bit integers, + 2^21, >> 22, clip to uint8  (Pillow: Resample.c
//           ImagingResampleHorizontal_8bpc; rounding to uint8 between the passes is part of the result);
//   pass 2 (vertical + ToTensor + Normalize): the same along y for the output ROWS that survive the crop, then
//           (u8 / 255 - mean) / std in float32 with correctly rounded divisions, written CHW as fp32 or fp16.
// Crop windows that leave the resized image read zeros (PIL's crop).  Byte / integer work, HBM-bound: 3 B read per input
// pixel of the needed rows + 3 B write/read of the intermediate + 6-12 B write per output pixel.
// The coefficient tables are Pillow's (precompute_coeffs + normalize_coeffs_8bpc, float64 on the host), computed once per
// geometry by isf_image_plan.  The per-pixel arithmetic is shared host / device source (isf_debug_image_aug_host).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "common.cuh"

extern "C" void isf_count_launch(int n);

namespace {

constexpr int IMG_PRECISION_BITS = 32 - 8 - 2;

inline double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

inline int resample_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(2.0 * filterscale) * 2 + 1;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output position xx of an axis in_size -> out_size
inline void resample_coeffs_at(int in_size, int out_size, int xx, int ksize, int* xmin_out, int* cnt_out, int* coef) {
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const double ss = 1.0 / filterscale;
  const double center = (xx + 0.5) * scale;
  int xmin = (int)(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double k[64];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) { k[x] = bicubic_filter((x + xmin - center + 0.5) * ss); ww += k[x]; }
  for (int x = 0; x < xmax; ++x) if (ww != 0.0) k[x] /= ww;
  for (int x = 0; x < ksize; ++x) {
    if (x >= xmax) coef[x] = 0;
    else if (k[x] < 0) coef[x] = (int)(-0.5 + k[x] * (1 << IMG_PRECISION_BITS));
    else coef[x] = (int)(0.5 + k[x] * (1 << IMG_PRECISION_BITS));
  }
  *xmin_out = xmin;
  *cnt_out = xmax;
}

struct ImagePlanHost {
  int ksx, ksy, row_lo, row_hi;
  std::vector<int> tab;       // xmin[ow] xcnt[ow] xcoef[ow*ksx] ymin[oh] ycnt[oh] ycoef[oh*ksy]
};

int build_plan(int in_w, int in_h, int rs_w, int rs_h, int crop_x, int crop_y, int out_w, int out_h, int flip, ImagePlanHost* P) {
  P->ksx = resample_ksize(in_w, rs_w);
  P->ksy = resample_ksize(in_h, rs_h);
  if (P->ksx > 64 || P->ksy > 64) { isf_set_error("image plan: down-scale factor above 15x is not supported"); return ISF_ERR_UNSUPPORTED; }
  P->tab.assign((size_t)out_w * (2 + P->ksx) + (size_t)out_h * (2 + P->ksy), 0);
  int* xmin = P->tab.data();
  int* xcnt = xmin + out_w;
  int* xcoef = xcnt + out_w;
  int* ymin = xcoef + (size_t)out_w * P->ksx;
  int* ycnt = ymin + out_h;
  int* ycoef = ycnt + out_h;
  for (int ox = 0; ox < out_w; ++ox) {
    const int rx = (flip ? out_w - 1 - ox : ox) + crop_x;
    if (rx < 0 || rx >= rs_w) { xmin[ox] = 0; xcnt[ox] = -1; continue; }          // outside the resized image: zeros
    if (rs_w == in_w) { xmin[ox] = rx; xcnt[ox] = 1; xcoef[(size_t)ox * P->ksx] = 1 << IMG_PRECISION_BITS; continue; }   // no pass: identity
    resample_coeffs_at(in_w, rs_w, rx, P->ksx, &xmin[ox], &xcnt[ox], xcoef + (size_t)ox * P->ksx);
  }
  P->row_lo = in_h; P->row_hi = 0;
  for (int oy = 0; oy < out_h; ++oy) {
    const int ry = oy + crop_y;
    if (ry < 0 || ry >= rs_h) { ymin[oy] = 0; ycnt[oy] = -1; continue; }
    if (rs_h == in_h) { ymin[oy] = ry; ycnt[oy] = 1; ycoef[(size_t)oy * P->ksy] = 1 << IMG_PRECISION_BITS; }
    else resample_coeffs_at(in_h, rs_h, ry, P->ksy, &ymin[oy], &ycnt[oy], ycoef + (size_t)oy * P->ksy);
    if (ymin[oy] < P->row_lo) P->row_lo = ymin[oy];
    if (ymin[oy] + ycnt[oy] > P->row_hi) P->row_hi = ymin[oy] + ycnt[oy];
  }
  if (P->row_hi <= P->row_lo) { P->row_lo = 0; P->row_hi = 1; }                    // every output row is outside: one dummy row
  return ISF_OK;
}

}  // namespace

// one 8-bit resampling tap sum of one channel: Pillow's clip8((2^21 + sum px * coef) >> 22); cnt < 0 = outside = 0
__host__ __device__ __forceinline__ unsigned char img_resample_px(const unsigned char* src, long long stride, const int* coef, int cnt) {
  if (cnt < 0) return 0;
  int ss = 1 << (32 - 8 - 2 - 1);
  for (int j = 0; j < cnt; ++j) ss += (int)src[j * stride] * coef[j];
  ss >>= (32 - 8 - 2);
  return (unsigned char)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

// aligned 32-bit load from a byte buffer (host: through memcpy, no aliasing assumptions)
__host__ __device__ __forceinline__ unsigned int img_load_u32(const unsigned char* p) {
#ifdef __CUDA_ARCH__
  return *(const unsigned int*)p;
#else
  unsigned int w;
  memcpy(&w, p, 4);
  return w;
#endif
}

// the same for FOUR adjacent byte columns of the interleaved intermediate (vertical pass: every byte column is an independent
// signal, the taps are whole rows apart): one aligned 32-bit load per tap instead of four byte loads.  src 4-byte aligned.
__host__ __device__ __forceinline__ void img_resample_px4(const unsigned char* src, long long stride, const int* coef, int cnt,
                                                          unsigned char (&out)[4]) {
  if (cnt < 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  int s0 = 1 << (32 - 8 - 2 - 1), s1 = s0, s2 = s0, s3 = s0;
  for (int j = 0; j < cnt; ++j) {
    const unsigned int w = img_load_u32(src + j * stride);
    const int k = coef[j];
    s0 += (int)(w & 0xFFu) * k;
    s1 += (int)((w >> 8) & 0xFFu) * k;
    s2 += (int)((w >> 16) & 0xFFu) * k;
    s3 += (int)(w >> 24) * k;
  }
  s0 >>= (32 - 8 - 2); s1 >>= (32 - 8 - 2); s2 >>= (32 - 8 - 2); s3 >>= (32 - 8 - 2);
  out[0] = (unsigned char)(s0 < 0 ? 0 : (s0 > 255 ? 255 : s0));
  out[1] = (unsigned char)(s1 < 0 ? 0 : (s1 > 255 ? 255 : s1));
  out[2] = (unsigned char)(s2 < 0 ? 0 : (s2 > 255 ? 255 : s2));
  out[3] = (unsigned char)(s3 < 0 ? 0 : (s3 > 255 ? 255 : s3));
}

// horizontal pass of one output pixel (3 channels, up to 8 taps = 24 contiguous source bytes at any alignment): eight aligned
// 32-bit loads, re-aligned with funnel shifts, instead of 21-24 byte loads.  The caller guarantees that the 32 bytes from
// (src & ~3) lie inside the image buffer and that 0 <= cnt <= 8.
__host__ __device__ __forceinline__ unsigned int img_funnel_r(unsigned int lo, unsigned int hi, unsigned int bytes) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, bytes * 8);
#else
  return bytes == 0 ? lo : (lo >> (bytes * 8)) | (hi << (32 - bytes * 8));
#endif
}

__host__ __device__ __forceinline__ void img_resample_rgb8(const unsigned char* src, const int* coef, int cnt, unsigned char* dst) {
  const uintptr_t a = (uintptr_t)src;
  const unsigned char* wp = (const unsigned char*)(a & ~(uintptr_t)3);
  const unsigned int mis = (unsigned int)(a & 3);
  unsigned int w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = img_load_u32(wp + 4 * i);
  unsigned int al[7];                                  // 28 bytes starting at src
#pragma unroll
  for (int i = 0; i < 7; ++i) al[i] = img_funnel_r(w[i], w[i + 1], mis);
  int s0 = 1 << (32 - 8 - 2 - 1), s1 = s0, s2 = s0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < cnt) {
      const int k = coef[j];
      const int b0 = j * 3, b1 = j * 3 + 1, b2 = j * 3 + 2;
      s0 += (int)((al[b0 >> 2] >> ((b0 & 3) * 8)) & 0xFFu) * k;
      s1 += (int)((al[b1 >> 2] >> ((b1 & 3) * 8)) & 0xFFu) * k;
      s2 += (int)((al[b2 >> 2] >> ((b2 & 3) * 8)) & 0xFFu) * k;
    }
  }
  s0 >>= (32 - 8 - 2); s1 >>= (32 - 8 - 2); s2 >>= (32 - 8 - 2);
  dst[0] = (unsigned char)(s0 < 0 ? 0 : (s0 > 255 ? 255 : s0));
  dst[1] = (unsigned char)(s1 < 0 ? 0 : (s1 > 255 ? 255 : s1));
  dst[2] = (unsigned char)(s2 < 0 ? 0 : (s2 > 255 ? 255 : s2));
}

// one output pixel of the horizontal pass: word path when the window allows it, byte path otherwise (same values)
__host__ __device__ __forceinline__ void img_horizontal_px(const unsigned char* img, long long img_bytes, long long src_off,
                                                           const int* coef, int cnt, unsigned char* dst) {
  const unsigned char* src = img + src_off;
  const long long base = src_off - (long long)((uintptr_t)src & 3);
  if (cnt >= 0 && cnt <= 8 && base >= 0 && base + 32 <= img_bytes && ((uintptr_t)img & 3) == 0) {
    img_resample_rgb8(src, coef, cnt, dst);
  } else {
    dst[0] = img_resample_px(src + 0, 3, coef, cnt);
    dst[1] = img_resample_px(src + 1, 3, coef, cnt);
    dst[2] = img_resample_px(src + 2, 3, coef, cnt);
  }
}

// ToTensor + Normalize of one value, float32 with correctly rounded divisions (torch: x.to(float32).div(255).sub_(mean).div_(std))
__host__ __device__ __forceinline__ float img_normalize_px(unsigned char u, float mean, float stdv) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), mean), stdv);
#else
  volatile float x = (float)u / 255.f;
  volatile float y = x - mean;
  return y / stdv;
#endif
}

struct ImageDims { int n, in_w, in_h, out_w, out_h, ksx, ksy, row_lo, rows; };

__global__ void __launch_bounds__(256) image_horizontal_kernel(const unsigned char* __restrict__ img, ImageDims d,
                                                               const int* __restrict__ xmin, const int* __restrict__ xcnt,
                                                               const int* __restrict__ xcoef, unsigned char* __restrict__ tmp) {
  const long long total = (long long)d.n * d.rows * d.out_w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(t % d.out_w);
    const long long r = t / d.out_w;                         // image * rows + row
    const int row = (int)(r % d.rows), im = (int)(r / d.rows);
    const long long off = (((long long)im * d.in_h + d.row_lo + row) * d.in_w + xmin[ox]) * 3;
    img_horizontal_px(img, (long long)d.n * d.in_h * d.in_w * 3, off, xcoef + (long long)ox * d.ksx, xcnt[ox], tmp + t * 3);
  }
}

template <typename OUT>
__global__ void __launch_bounds__(256) image_vertical_normalize_kernel(const unsigned char* __restrict__ tmp, ImageDims d,
                                                                       const int* __restrict__ ymin, const int* __restrict__ ycnt,
                                                                       const int* __restrict__ ycoef, float m0, float m1, float m2,
                                                                       float s0, float s1, float s2, OUT* __restrict__ out) {
  const long long total = (long long)d.n * d.out_h * d.out_w;
  const long long plane = (long long)d.out_h * d.out_w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(t % d.out_w);
    const long long r = t / d.out_w;
    const int oy = (int)(r % d.out_h), im = (int)(r / d.out_h);
    const int cnt = ycnt[oy];
    const long long stride = (long long)d.out_w * 3;
    const unsigned char* src = tmp + (((long long)im * d.rows + (ymin[oy] - d.row_lo)) * d.out_w + ox) * 3;
    const int* k = ycoef + (long long)oy * d.ksy;
    OUT* o = out + (long long)im * 3 * plane + (long long)oy * d.out_w + ox;
    o[0] = (OUT)img_normalize_px(img_resample_px(src + 0, stride, k, cnt), m0, s0);
    o[plane] = (OUT)img_normalize_px(img_resample_px(src + 1, stride, k, cnt), m1, s1);
    o[2 * plane] = (OUT)img_normalize_px(img_resample_px(src + 2, stride, k, cnt), m2, s2);
  }
}

// vertical pass, four byte columns per thread (row pitch out_w * 3 a multiple of 4): byte column b of a row = pixel b / 3,
// channel b % 3.  7 word loads instead of 28 byte loads per four outputs.
template <typename OUT>
__global__ void __launch_bounds__(256) image_vertical_normalize4_kernel(const unsigned char* __restrict__ tmp, ImageDims d,
                                                                        const int* __restrict__ ymin, const int* __restrict__ ycnt,
                                                                        const int* __restrict__ ycoef, float m0, float m1, float m2,
                                                                        float s0, float s1, float s2, OUT* __restrict__ out) {
  const int quads = d.out_w * 3 / 4;
  const long long total = (long long)d.n * d.out_h * quads;
  const long long plane = (long long)d.out_h * d.out_w;
  const long long stride = (long long)d.out_w * 3;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(t % quads);
    const long long r = t / quads;
    const int oy = (int)(r % d.out_h), im = (int)(r / d.out_h);
    const unsigned char* src = tmp + ((long long)im * d.rows + (ymin[oy] - d.row_lo)) * stride + qd * 4;
    unsigned char u[4];
    img_resample_px4(src, stride, ycoef + (long long)oy * d.ksy, ycnt[oy], u);
    OUT* o = out + (long long)im * 3 * plane + (long long)oy * d.out_w;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int b = qd * 4 + e, px = b / 3, c = b - px * 3;
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
      o[c * plane + px] = (OUT)img_normalize_px(u[e], mean, sd);
    }
  }
}

static int image_check_geometry(int in_w, int in_h, int rs_w, int rs_h, int out_w, int out_h) {
  ISF_CHECK_ARG(in_w >= 1 && in_h >= 1 && rs_w >= 1 && rs_h >= 1 && out_w >= 1 && out_h >= 1);
  ISF_CHECK_ARG((long long)in_w * in_h < (1ll << 28) && (long long)out_w * out_h < (1ll << 28));
  return ISF_OK;
}

// bytes of the device plan of one geometry, and the input rows [rows[0], rows[1]) the vertical pass reads
extern "C" int isf_image_plan_size(int in_w, int in_h, int rs_w, int rs_h, int out_w, int out_h, size_t* plan_bytes) {
  ISF_CHECK_ARG(plan_bytes);
  int rc = image_check_geometry(in_w, in_h, rs_w, rs_h, out_w, out_h);
  if (rc) return rc;
  const int ksx = resample_ksize(in_w, rs_w), ksy = resample_ksize(in_h, rs_h);
  *plan_bytes = ((size_t)out_w * (2 + ksx) + (size_t)out_h * (2 + ksy)) * sizeof(int);
  return ISF_OK;
}

// resize_dims = (rs_w, rs_h), crop = (crop_x, crop_y, crop_x + out_w, crop_y + out_h), flip = FLIP_LEFT_RIGHT after the crop
extern "C" int isf_image_plan(int in_w, int in_h, int rs_w, int rs_h, int crop_x, int crop_y, int out_w, int out_h, int flip,
                              void* plan_dev, size_t plan_bytes, int* rows, void* stream) {
  ISF_CHECK_ARG(plan_dev && rows);
  int rc = image_check_geometry(in_w, in_h, rs_w, rs_h, out_w, out_h);
  if (rc) return rc;
  ImagePlanHost P;
  rc = build_plan(in_w, in_h, rs_w, rs_h, crop_x, crop_y, out_w, out_h, flip, &P);
  if (rc) return rc;
  if (plan_bytes < P.tab.size() * sizeof(int)) { isf_set_error("image plan: buffer too small"); return ISF_ERR_WORKSPACE; }
  // pageable source: the copy is staged before the call returns, so the vector may go out of scope
  ISF_CUDA(cudaMemcpyAsync(plan_dev, P.tab.data(), P.tab.size() * sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  rows[0] = P.row_lo; rows[1] = P.row_hi;
  return ISF_OK;
}

extern "C" int isf_image_aug_workspace_size(int n, int out_w, int row_lo, int row_hi, size_t* bytes) {
  ISF_CHECK_ARG(bytes && n >= 0 && out_w >= 1 && row_hi > row_lo && row_lo >= 0);
  *bytes = (size_t)n * (row_hi - row_lo) * out_w * 3 + 256;
  return ISF_OK;
}

// img [n, in_h, in_w, 3] uint8 (DEVICE) -> out [n, 3, out_h, out_w] fp32 | fp16; mean / std = HOST [3]
extern "C" int isf_image_aug_normalize(const unsigned char* img, int n, int in_w, int in_h, int rs_w, int rs_h, int out_w,
                                       int out_h, int row_lo, int row_hi, const void* plan_dev, const float* mean,
                                       const float* stdv, void* out, int out_is_half, void* workspace, size_t ws_bytes,
                                       void* stream) {
  ISF_CHECK_ARG(n >= 0 && mean && stdv);
  if (n == 0) return ISF_OK;
  ISF_CHECK_ARG(img && plan_dev && out && workspace);
  int rc = image_check_geometry(in_w, in_h, rs_w, rs_h, out_w, out_h);
  if (rc) return rc;
  ISF_CHECK_ARG(row_lo >= 0 && row_hi > row_lo && row_hi <= in_h);
  ImageDims d{n, in_w, in_h, out_w, out_h, resample_ksize(in_w, rs_w), resample_ksize(in_h, rs_h), row_lo, row_hi - row_lo};
  if (ws_bytes < (size_t)n * d.rows * out_w * 3) { isf_set_error("image_aug_normalize: workspace too small"); return ISF_ERR_WORKSPACE; }
  const int* xmin = (const int*)plan_dev;
  const int* xcnt = xmin + out_w;
  const int* xcoef = xcnt + out_w;
  const int* ymin = xcoef + (size_t)out_w * d.ksx;
  const int* ycnt = ymin + out_h;
  const int* ycoef = ycnt + out_h;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* tmp = (unsigned char*)workspace;
  auto grid = [](long long total) { long long b = (total + 255) / 256; return (int)(b > ISF_NUM_SMS * 32 ? ISF_NUM_SMS * 32 : b); };
  image_horizontal_kernel<<<grid((long long)n * d.rows * out_w), 256, 0, st>>>(img, d, xmin, xcnt, xcoef, tmp);
  const bool quad = (out_w * 3) % 4 == 0 && ((uintptr_t)tmp & 3) == 0;      // word loads need a 4-byte row pitch
#define ISF_IMG_V(KERNEL, T, TOTAL)                                                                                  \
  KERNEL<T><<<grid(TOTAL), 256, 0, st>>>(tmp, d, ymin, ycnt, ycoef, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2], (T*)out)
  if (quad) {
    const long long total = (long long)n * out_h * (out_w * 3 / 4);
    if (out_is_half) ISF_IMG_V(image_vertical_normalize4_kernel, __half, total); else ISF_IMG_V(image_vertical_normalize4_kernel, float, total);
  } else {
    const long long total = (long long)n * out_h * out_w;
    if (out_is_half) ISF_IMG_V(image_vertical_normalize_kernel, __half, total); else ISF_IMG_V(image_vertical_normalize_kernel, float, total);
  }
#undef ISF_IMG_V
  isf_count_launch(2);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// TEST HOOK (host memory, no GPU): plan + both passes with the kernels' own per-pixel source; out fp32 [n, 3, out_h, out_w]
extern "C" int isf_debug_image_aug_host(const unsigned char* img, int n, int in_w, int in_h, int rs_w, int rs_h, int crop_x,
                                        int crop_y, int out_w, int out_h, int flip, const float* mean, const float* stdv,
                                        float* out) {
  ISF_CHECK_ARG(img && mean && stdv && out && n >= 1);
  int rc = image_check_geometry(in_w, in_h, rs_w, rs_h, out_w, out_h);
  if (rc) return rc;
  ImagePlanHost P;
  rc = build_plan(in_w, in_h, rs_w, rs_h, crop_x, crop_y, out_w, out_h, flip, &P);
  if (rc) return rc;
  const int* xmin = P.tab.data();
  const int* xcnt = xmin + out_w;
  const int* xcoef = xcnt + out_w;
  const int* ymin = xcoef + (size_t)out_w * P.ksx;
  const int* ycnt = ymin + out_h;
  const int* ycoef = ycnt + out_h;
  const int rows = P.row_hi - P.row_lo;
  std::vector<unsigned char> tmp((size_t)rows * out_w * 3);
  const long long plane = (long long)out_h * out_w;
  for (int im = 0; im < n; ++im) {
    for (int row = 0; row < rows; ++row)
      for (int ox = 0; ox < out_w; ++ox) {
        const long long off = (((long long)im * in_h + P.row_lo + row) * in_w + xmin[ox]) * 3;
        img_horizontal_px(img, (long long)n * in_h * in_w * 3, off, xcoef + (size_t)ox * P.ksx, xcnt[ox], &tmp[((size_t)row * out_w + ox) * 3]);
      }
    const bool quad = (out_w * 3) % 4 == 0 && ((uintptr_t)tmp.data() & 3) == 0;     // the path the device launch takes
    for (int oy = 0; oy < out_h; ++oy) {
      if (quad) {
        for (int qd = 0; qd < out_w * 3 / 4; ++qd) {
          unsigned char u[4];
          img_resample_px4(tmp.data() + ((size_t)(ymin[oy] - P.row_lo)) * out_w * 3 + qd * 4, (long long)out_w * 3,
                           ycoef + (size_t)oy * P.ksy, ycnt[oy], u);
          for (int e = 0; e < 4; ++e) {
            const int b = qd * 4 + e, px = b / 3, c = b - px * 3;
            out[(long long)im * 3 * plane + c * plane + (long long)oy * out_w + px] = img_normalize_px(u[e], mean[c], stdv[c]);
          }
        }
        continue;
      }
      for (int ox = 0; ox < out_w; ++ox) {
        const unsigned char* src = tmp.data() + (((size_t)(ymin[oy] - P.row_lo)) * out_w + ox) * 3;
        for (int c = 0; c < 3; ++c)
          out[(long long)im * 3 * plane + c * plane + (long long)oy * out_w + ox] =
              img_normalize_px(img_resample_px(src + c, (long long)out_w * 3, ycoef + (size_t)oy * P.ksy, ycnt[oy]), mean[c], stdv[c]);
      }
    }
  }
  return ISF_OK;
}
