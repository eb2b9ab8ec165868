// Instance-context decoder of the Instance-Guided Fusion (a11) as ONE kernel.
//
// Reference: InsContextAtt.forward + DeformableTransformerDecoderLayer.forward x 2 (mmdet3d/models/middle_encoders/
// fusion_encoder.py:600-674, 768-830) with MSDeformAttn (:523-598) and ms_deform_attn_core (ops/functions/
// ms_deform_attn_func.py:41-61): per frame 200 instance queries x 128 channels go through, per layer, self-attention
// (8 heads x 16), deformable cross-attention into the 180 x 180 scene map (8 heads x 16 points, bilinear), an FFN and three
// LayerNorms — about 60 small launches (cuBLAS sgemm + elementwise) in a straightforward port.  Here the whole stack, the
// learned query position embedding in front of it and the key/value projection of the instance->scene attention behind it
// run in one launch:
//
//   * a thread-block CLUSTER of 8 CTAs per frame; CTA r owns queries [r*RP, (r+1)*RP), RP = ceil(K/8) <= 32; the fp32 residual
//     stream, the position embedding and every intermediate of those rows stay in shared memory;
//   * every linear layer is a [32 x 128] x [128 x N] product on mma.sync.m16n8k16 (fp16 operands, fp32 accumulate): A from shared
//     memory by ldmatrix, the weights straight from global memory in a host-packed fragment order (one coalesced 2 KB read per
//     8 output columns, L2-resident: 64 CTAs read the same 0.75 MB);
//   * self-attention needs the keys / values of all K queries: each CTA publishes its rows, the cluster exchanges them through
//     distributed shared memory (two cluster barriers per layer), then warp = (head, 16-row tile) runs QK^T / softmax / PV on
//     the tensor cores with the probabilities kept in registers (online softmax over two 112-key blocks);
//   * deformable sampling: warp = one (query, head), lanes = 16 sampling points x 2 channel octets; a lane computes its point's
//     softmax weight (16-lane shuffles), its bilinear corners, and fetches 4 x 16 bytes (8 fp16 channels per corner); the 16
//     partial sums are combined by shuffles.  (The round-1 kernel, like the reference's im2col kernel, ran one thread per output
//     channel that re-derived the same corners 16 times and fetched 2 bytes per load.)
//
// Fixed geometry (what the config instantiates): d_model 128, 8 heads, 1 level, 16 points, d_ffn 128, K <= 224.
#include <cooperative_groups.h>
#include "common.cuh"

extern "C" void isf_count_launch(int n);
namespace cg = cooperative_groups;

namespace dec {

constexpr int C = 128, H = 8, D = 16, P = 16;
constexpr int CL = 8;                 // CTAs per frame
constexpr int THREADS = 512, WARPS = 16;
constexpr int RM = 32;                // padded rows per CTA
constexpr int AP = C + 8;             // pitch of the fp16 A operand (halves)
constexpr int KVP = 2 * C + 8;        // pitch of a (k | v) row (halves)
constexpr int NKP = 224;              // key capacity (two blocks of 112)
constexpr int LAYER_ROWS = 1280, LAYER_F = 2048;
// packed weight rows
constexpr int R_QPE = 0, R_LAYER0 = 128;
constexpr int RQK = 0, RV = 256, RO = 384, RSOAW = 512, ROP = 896, R1 = 1024, R2 = 1152;
// fp32 parameter offsets
constexpr int F_QW1 = 0, F_QS = 256, F_QT = 384, F_QB2 = 512, F_LAYER0 = 640;
constexpr int FBQK = 0, FBV = 256, FBO = 384, FLN2G = 512, FLN2B = 640, FBSOAW = 768, FBOP = 1152, FLN1G = 1280, FLN1B = 1408,
              FB1 = 1536, FB2 = 1664, FLN3G = 1792, FLN3B = 1920;

struct Params {
  const __half* toks;     // [B, S, C] conv_scene output tokens
  const int* top_idx;     // [B, K] cell n = y*G + x
  const __half* value;    // [B*S, ldv]: value_proj(tokens + key_pos) of layer li at columns [li*C, (li+1)*C)
  int ldv;
  const uint4* wp;        // packed fp16 weights (fragment order), see isf_ins_context_decoder
  const float* fp;        // fp32 parameters
  int K, G, layers;
  float* out;             // [B, K, C] fp32 decoder output
  __half* kv_out;         // [B, K, 2C] fp16 = out . Wkv^T + bkv (instance->scene keys | values), may be null
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *(const uint32_t*)&h;
}
__device__ __forceinline__ uint32_t pack_2h(__half lo, __half hi) {
  return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}
// A fragment (16 rows x 16 k) of a row-major fp16 matrix in shared memory
__device__ __forceinline__ void lda(const __half* base, int pitch, int row0, int k0, int lane, uint32_t (&a)[4]) {
  const __half* p = base + (size_t)(row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * pitch + k0 + (lane >> 4) * 8;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

// Y[32][ldy] = A16[32][AP] . W^T for the n_rows output columns whose packed weights start at `wp` (2 KB per 8 columns)
__device__ __forceinline__ void gemm32(const __half* A16, const uint4* __restrict__ wp, int n_rows, float* Y, int ldy, int warp,
                                       int lane) {
  const int gr = lane >> 2, gc = (lane & 3) * 2;
  for (int nt = warp; nt < n_rows / 8; nt += WARPS) {
    const uint4* bp = wp + (size_t)nt * 128 + lane * 4;
    uint32_t bw[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = __ldg(bp + q);
      bw[4 * q] = u.x; bw[4 * q + 1] = u.y; bw[4 * q + 2] = u.z; bw[4 * q + 3] = u.w;
    }
    float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t a[4];
        lda(A16, AP, mt * 16, ks * 16, lane, a);
        mma16816(c[mt], a, bw[2 * ks], bw[2 * ks + 1]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float* y0 = Y + (size_t)(mt * 16 + gr) * ldy + nt * 8 + gc;
      *(float2*)y0 = make_float2(c[mt][0], c[mt][1]);
      *(float2*)(y0 + 8 * ldy) = make_float2(c[mt][2], c[mt][3]);
    }
  }
}

// X[r][:] = LayerNorm(X[r][:] + Y[r][:] + bias) * gamma + beta   (eps 1e-5), 2 rows per warp, 4 columns per lane
__device__ __forceinline__ void add_ln(float* X, const float* Y, int ldy, const float* __restrict__ bias, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int warp, int lane) {
#pragma unroll
  for (int rr = 0; rr < RM / WARPS; ++rr) {
    const int r = warp + rr * WARPS;
    const float4 x = *(const float4*)(X + r * C + lane * 4);
    const float4 y = *(const float4*)(Y + (size_t)r * ldy + lane * 4);
    const float4 b = *(const float4*)(bias + lane * 4);
    float v[4] = {x.x + y.x + b.x, x.y + y.y + b.y, x.z + y.z + b.z, x.w + y.w + b.w};
    float s = v[0] + v[1] + v[2] + v[3];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] -= mean; q += v[j] * v[j]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + 1e-5f);
    const float4 g = *(const float4*)(gamma + lane * 4), be = *(const float4*)(beta + lane * 4);
    *(float4*)(X + r * C + lane * 4) = make_float4(v[0] * rstd * g.x + be.x, v[1] * rstd * g.y + be.y, v[2] * rstd * g.z + be.z,
                                                   v[3] * rstd * g.w + be.w);
  }
}

// A16 = half(X (+ Q))
__device__ __forceinline__ void to_a16(__half* A16, const float* X, const float* Q, int tid) {
  for (int e = tid; e < RM * C / 2; e += THREADS) {
    const int r = e / (C / 2), c = (e % (C / 2)) * 2;
    float2 v = *(const float2*)(X + r * C + c);
    if (Q) { const float2 q = *(const float2*)(Q + r * C + c); v.x += q.x; v.y += q.y; }
    *(__half2*)(A16 + r * AP + c) = __floats2half2_rn(v.x, v.y);
  }
}

__global__ void __launch_bounds__(THREADS, 1) ins_context_decoder_kernel(const Params p) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint8_t smem[];
  float* X = (float*)smem;                                  // [32][128] residual stream
  float* QPE = X + RM * C;                                  // [32][128] query position embedding
  __half* A16 = (__half*)(QPE + RM * C);                    // [32][AP]
  __half* Q16 = A16 + RM * AP;                              // [32][AP]
  __half* KVloc = Q16 + RM * AP;                            // [32][2C]   this CTA's (k | v) rows, read by the whole cluster
  __half* KVall = KVloc + RM * 2 * C;                       // [NKP][KVP] every query's (k | v)
  float* Y = (float*)KVall;                                 // [32][384] GEMM output, aliases KVall (never live together)
  __shared__ float ref0[RM], ref1[RM];
  __shared__ int cell[RM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int rank = (int)cluster.block_rank();
  const int K = p.K, G = p.G, S = G * G;
  const int RP = (K + CL - 1) / CL;
  const int q0 = rank * RP;
  const int nrow = max(0, min(RP, K - q0));                 // valid rows of this CTA
  const float* fp = p.fp;

  // ---- queries: scene tokens at the top-K cells, reference points, position embedding
  if (tid < RM) {
    int n = 0;
    if (tid < nrow) n = p.top_idx[(size_t)b * K + q0 + tid];
    cell[tid] = n;
    ref0[tid] = ((float)(n / G) + 0.5f) / (float)G;          // (y + .5) / G — consumed as the WIDTH coordinate
    ref1[tid] = ((float)(n % G) + 0.5f) / (float)G;
  }
  __syncthreads();
  for (int e = tid; e < RM * C; e += THREADS) {
    const int r = e / C, c = e % C;
    X[e] = r < nrow ? __half2float(p.toks[((size_t)b * S + cell[r]) * C + c]) : 0.f;
    const float h1 = fmaxf((fp[F_QW1 + c * 2] * ref0[r] + fp[F_QW1 + c * 2 + 1] * ref1[r]) * fp[F_QS + c] + fp[F_QT + c], 0.f);
    A16[r * AP + c] = __float2half(h1);
  }
  __syncthreads();
  gemm32(A16, p.wp + (size_t)R_QPE * 16, C, Y, 384, warp, lane);
  __syncthreads();
  for (int e = tid; e < RM * C; e += THREADS) QPE[e] = Y[(e / C) * 384 + (e % C)] + fp[F_QB2 + (e % C)];
  __syncthreads();

  for (int li = 0; li < p.layers; ++li) {
    const uint4* wl = p.wp + (size_t)(R_LAYER0 + li * LAYER_ROWS) * 16;
    const float* fl = fp + F_LAYER0 + li * LAYER_F;
    // ---------------------------------------------------------------- self-attention
    to_a16(A16, X, QPE, tid);
    __syncthreads();
    gemm32(A16, wl + (size_t)RQK * 16, 2 * C, Y, 384, warp, lane);          // q | k of (tgt + pos)
    __syncthreads();
    to_a16(A16, X, nullptr, tid);
    for (int e = tid; e < RM * C; e += THREADS) {
      const int r = e / C, c = e % C;
      Q16[r * AP + c] = __float2half(Y[r * 384 + c] + fl[FBQK + c]);
      KVloc[r * 2 * C + c] = __float2half(Y[r * 384 + C + c] + fl[FBQK + C + c]);
    }
    __syncthreads();
    gemm32(A16, wl + (size_t)RV * 16, C, Y, 384, warp, lane);               // v of tgt
    __syncthreads();
    for (int e = tid; e < RM * C; e += THREADS) {
      const int r = e / C, c = e % C;
      KVloc[r * 2 * C + C + c] = __float2half(Y[r * 384 + c] + fl[FBV + c]);
    }
    cluster.sync();                                                          // every CTA's (k | v) rows are published
    for (int e = tid; e < NKP * (2 * C / 8); e += THREADS) {
      const int j = e / (2 * C / 8), c8 = (e % (2 * C / 8)) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (j < K) {
        const __half* src = cluster.map_shared_rank(KVloc, j / RP) + (size_t)(j % RP) * 2 * C + c8;
        v = *(const uint4*)src;
      }
      *(uint4*)(KVall + (size_t)j * KVP + c8) = v;
    }
    cluster.sync();                                                          // peers have read this CTA's rows (KVloc reusable)
    {
      const int h = warp & 7, mt = warp >> 3;
      const int gr = lane >> 2, gc = (lane & 3) * 2;
      uint32_t qa[4];
      lda(Q16, AP, mt * 16, h * D, lane, qa);
      const float sl2 = 0.25f * 1.44269504088896f;                           // 1/sqrt(16), log2 domain
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 1
      for (int kb = 0; kb < NKP; kb += 112) {
        if (kb >= K) break;
        float s[14][4];
#pragma unroll
        for (int nt = 0; nt < 14; ++nt) {
          s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
          const __half* kr = KVall + (size_t)(kb + nt * 8 + gr) * KVP + h * D;
          mma16816(s[nt], qa, *(const uint32_t*)(kr + gc), *(const uint32_t*)(kr + gc + 8));
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 14; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = kb + nt * 8 + gc + (e & 1);
            s[nt][e] = key < K ? s[nt][e] * sl2 : -INFINITY;
          }
          bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
          bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);               // finite: the block holds at least one valid key
        const float a0 = exp2f(m0 - n0), a1 = exp2f(m1 - n1);               // exp2(-inf) = 0 on the first block
        m0 = n0; m1 = n1;
        float bl0 = 0.f, bl1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 14; ++nt) {
          s[nt][0] = exp2f(s[nt][0] - m0); s[nt][1] = exp2f(s[nt][1] - m0);
          s[nt][2] = exp2f(s[nt][2] - m1); s[nt][3] = exp2f(s[nt][3] - m1);
          bl0 += s[nt][0] + s[nt][1];
          bl1 += s[nt][2] + s[nt][3];
        }
        l0 = l0 * a0 + bl0; l1 = l1 * a1 + bl1;
#pragma unroll
        for (int n = 0; n < 2; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
#pragma unroll
        for (int kt = 0; kt < 7; ++kt) {
          uint32_t pa[4];
          pa[0] = pack_h2(s[2 * kt][0], s[2 * kt][1]);
          pa[1] = pack_h2(s[2 * kt][2], s[2 * kt][3]);
          pa[2] = pack_h2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
          pa[3] = pack_h2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
          const __half* vr = KVall + (size_t)(kb + kt * 16 + gc) * KVP + C + h * D + gr;     // V[key][dv]
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const uint32_t b0 = pack_2h(vr[n * 8], vr[KVP + n * 8]);
            const uint32_t b1 = pack_2h(vr[8 * KVP + n * 8], vr[9 * KVP + n * 8]);
            mma16816(o[n], pa, b0, b1);
          }
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        *(__half2*)(A16 + (size_t)(mt * 16 + gr) * AP + h * D + n * 8 + gc) = __floats2half2_rn(o[n][0] * i0, o[n][1] * i0);
        *(__half2*)(A16 + (size_t)(mt * 16 + gr + 8) * AP + h * D + n * 8 + gc) = __floats2half2_rn(o[n][2] * i1, o[n][3] * i1);
      }
    }
    __syncthreads();                                                         // attention output in A16; KVall dead -> Y free
    gemm32(A16, wl + (size_t)RO * 16, C, Y, 384, warp, lane);
    __syncthreads();
    add_ln(X, Y, 384, fl + FBO, fl + FLN2G, fl + FLN2B, warp, lane);
    __syncthreads();
    // ---------------------------------------------------------------- deformable cross-attention
    to_a16(A16, X, QPE, tid);
    __syncthreads();
    gemm32(A16, wl + (size_t)RSOAW * 16, 384, Y, 384, warp, lane);           // sampling offsets (256) | attention logits (128)
    __syncthreads();
    {
      const int pt = lane & 15, hf = lane >> 4;
      const float fg = (float)G;
      for (int pr = warp; pr < nrow * H; pr += WARPS) {
        const int r = pr / H, h = pr % H;
        const float* yr = Y + (size_t)r * 384;
        float lg = yr[256 + h * P + pt] + fl[FBSOAW + 256 + h * P + pt];
        float mx = lg;
#pragma unroll
        for (int o = 8; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float w = __expf(lg - mx), ws = w;
#pragma unroll
        for (int o = 8; o; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
        w /= ws;
        const float2 so = *(const float2*)(yr + (h * P + pt) * 2);
        const float lw = ref0[r] + (so.x + fl[FBSOAW + (h * P + pt) * 2]) / fg;
        const float lh = ref1[r] + (so.y + fl[FBSOAW + (h * P + pt) * 2 + 1]) / fg;
        const float him = lh * fg - 0.5f, wim = lw * fg - 0.5f;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (him > -1.f && wim > -1.f && him < fg && wim < fg) {
          const int hl = (int)floorf(him), wl_ = (int)floorf(wim);
          const float lhh = him - hl, lww = wim - wl_, hh = 1.f - lhh, hw = 1.f - lww;
          const __half* vb = p.value + (size_t)b * S * p.ldv + li * C + h * D + hf * 8;
          const float cw[4] = {hh * hw * w, hh * lww * w, lhh * hw * w, lhh * lww * w};
          const bool ok[4] = {hl >= 0 && wl_ >= 0, hl >= 0 && wl_ + 1 <= G - 1, hl + 1 <= G - 1 && wl_ >= 0, hl + 1 <= G - 1 && wl_ + 1 <= G - 1};
          const int cy[4] = {hl, hl, hl + 1, hl + 1}, cx[4] = {wl_, wl_ + 1, wl_, wl_ + 1};
          uint4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = ok[k] ? __ldg((const uint4*)(vb + (size_t)(cy[k] * G + cx[k]) * p.ldv)) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __half2* hp = (const __half2*)&v[k];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(hp[e]);
              acc[2 * e] += cw[k] * f.x; acc[2 * e + 1] += cw[k] * f.y;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int o = 8; o; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (pt == 0) {
          uint4 u;
          __half2* hp = (__half2*)&u;
#pragma unroll
          for (int e = 0; e < 4; ++e) hp[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
          *(uint4*)(A16 + (size_t)r * AP + h * D + hf * 8) = u;
        }
      }
    }
    __syncthreads();
    gemm32(A16, wl + (size_t)ROP * 16, C, Y, 384, warp, lane);
    __syncthreads();
    add_ln(X, Y, 384, fl + FBOP, fl + FLN1G, fl + FLN1B, warp, lane);
    __syncthreads();
    // ---------------------------------------------------------------- FFN
    to_a16(A16, X, nullptr, tid);
    __syncthreads();
    gemm32(A16, wl + (size_t)R1 * 16, C, Y, 384, warp, lane);
    __syncthreads();
    for (int e = tid; e < RM * C; e += THREADS) {
      const int r = e / C, c = e % C;
      A16[r * AP + c] = __float2half(fmaxf(Y[r * 384 + c] + fl[FB1 + c], 0.f));
    }
    __syncthreads();
    gemm32(A16, wl + (size_t)R2 * 16, C, Y, 384, warp, lane);
    __syncthreads();
    add_ln(X, Y, 384, fl + FB2, fl + FLN3G, fl + FLN3B, warp, lane);
    __syncthreads();
  }

  // ---- outputs: the decoder state (fp32) and the instance->scene (k | v) projection (fp16)
  for (int e = tid; e < nrow * C; e += THREADS) p.out[((size_t)b * K + q0 + e / C) * C + (e % C)] = X[e];
  if (p.kv_out) {
    to_a16(A16, X, nullptr, tid);
    __syncthreads();
    gemm32(A16, p.wp + (size_t)(R_LAYER0 + p.layers * LAYER_ROWS) * 16, 2 * C, Y, 384, warp, lane);
    __syncthreads();
    const float* bkv = fp + F_LAYER0 + p.layers * LAYER_F;
    for (int e = tid; e < nrow * 2 * C; e += THREADS) {
      const int r = e / (2 * C), c = e % (2 * C);
      p.kv_out[((size_t)b * K + q0 + r) * 2 * C + c] = __float2half(Y[r * 384 + c] + bkv[c]);
    }
  }
  cluster.sync();      // nobody exits while a peer could still address its shared memory
}

constexpr size_t SMEM_BYTES = (size_t)2 * RM * C * 4 + (size_t)2 * RM * AP * 2 + (size_t)RM * 2 * C * 2 + (size_t)NKP * KVP * 2;
static_assert((size_t)RM * 384 * 4 <= (size_t)NKP * KVP * 2, "GEMM output aliases the key/value buffer");

}  // namespace dec

// see include/isfusion_b200.h
extern "C" int isf_ins_context_decoder_sizes(int layers, size_t* packed_weight_bytes, size_t* param_floats) {
  ISF_CHECK_ARG(layers >= 1 && layers <= 8 && packed_weight_bytes && param_floats);
  *packed_weight_bytes = (size_t)(dec::R_LAYER0 + layers * dec::LAYER_ROWS + 2 * dec::C) * dec::C * 2;
  *param_floats = (size_t)dec::F_LAYER0 + (size_t)layers * dec::LAYER_F + 2 * dec::C;
  return ISF_OK;
}

extern "C" int isf_ins_context_decoder(const void* scene_tokens, const int* top_idx, const void* value, int ldv,
                                       const void* packed_weights, const float* params, int batch, int grid, int num_queries,
                                       int channels, int heads, int points, int layers, float* out, void* kv_out, void* stream) {
  ISF_CHECK_ARG(scene_tokens && top_idx && value && packed_weights && params && out);
  ISF_CHECK_ARG(batch >= 1 && grid >= 2 && num_queries >= 1 && layers >= 1 && layers <= 8 && ldv >= layers * channels && ldv % 8 == 0);
  if (channels != dec::C || heads != dec::H || points != dec::P || num_queries > dec::NKP) {
    isf_set_error("ins_context_decoder: built for d_model 128, 8 heads, 16 points, <= 224 queries (got %d, %d, %d, %d)", channels, heads,
                  points, num_queries);
    return ISF_ERR_UNSUPPORTED;
  }
  ISF_CHECK_ARG(((uintptr_t)value & 15) == 0 && ((uintptr_t)packed_weights & 15) == 0 && ((uintptr_t)params & 15) == 0);
  dec::Params p{};
  p.toks = (const __half*)scene_tokens; p.top_idx = top_idx; p.value = (const __half*)value; p.ldv = ldv;
  p.wp = (const uint4*)packed_weights; p.fp = params; p.K = num_queries; p.G = grid; p.layers = layers;
  p.out = out; p.kv_out = (__half*)kv_out;
  int dev = 0;
  ISF_CUDA(cudaGetDevice(&dev));
  static bool configured[64] = {};
  if (dev < 0 || dev >= 64) { isf_set_error("ins_context_decoder: device ordinal %d out of range", dev); return ISF_ERR_INVALID; }
  if (!configured[dev]) {
    ISF_CUDA(cudaFuncSetAttribute(dec::ins_context_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec::SMEM_BYTES));
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(dec::CL, batch);
  cfg.blockDim = dim3(dec::THREADS);
  cfg.dynamicSmemBytes = dec::SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = dec::CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, dec::ins_context_decoder_kernel, p);
  if (e != cudaSuccess) { isf_set_error("ins_context_decoder launch: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
  isf_count_launch(1);
  return ISF_OK;
}
