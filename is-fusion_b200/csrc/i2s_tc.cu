// a12, instance -> scene cross attention (reference mmdet3d/models/middle_encoders/fusion_encoder.py:441-462: nn.MultiheadAttention of
// 32 400 BEV tokens over the 200 instance tokens, 8 heads of 16) on tcgen05 tiles with TMA-staged Q / K.
//
// CTA = one tile of 128 queries x all 8 heads, persistent over the tiles of its frame.  Per head h:
//   S_h [128 x 208] = Q_h [128 x 16] . K_h^T            one tcgen05.mma (M = 128, N = 208, K = 16) into TMEM
//   P_h = exp2((S_h - max) * scale * log2 e)            4 softmax warps, thread = query row: two passes over the TMEM row
//                                                       (max, then exp + sum), P written as fp16 into the 128B-swizzled
//                                                       K-major layout the tensor core reads (un-normalised, <= 1)
//   O_h [128 x 16] = P_h [128 x 208] . V_h [208 x 16]   13 tcgen05.mma (M = 128, N = 16, K = 16) into TMEM; read back, / sum, stored
// Q tile (128 x 128 channels) and K (208 x 128, rows >= num_keys zero-filled by the TMA) arrive by TMA as two 64-channel chunks
// each; a head's 16 channels are a 32-byte step inside the 128-byte swizzle row, i.e. the K-step advance of the GEMM kernels.
// V is needed K-major along the KEYS (B operand [N = 16 channels x K = 208 keys]): it is transposed once per CTA into shared
// memory.  S is double-buffered in TMEM (2 x 208 columns) and O too (2 x 16), so QK^T of head h + 1 and P.V of head h - 1
// overlap the softmax of head h.  Warp 0 = TMA producer, warp 1 = MMA issuer (elected lane of the converged warp, tc_gemm.cuh),
// warps 2-5 = softmax / epilogue.
// MEASURED (B200, profiles/call27_r02.log): correct on the first run (< 1e-3 against fp32 attention of the same operands, ragged
// tiles, < 208 keys, several frames) and 3x SLOWER than the mma.sync kernel of fusion.cu: 143 vs 44 us at 32 400 queries x 200 keys.
// With a head dimension of 16 the tensor-core work per head is one 128 x 208 x 16 instruction + 13 tiny ones, while every score
// has to come back from TMEM (2 passes x 13 tcgen05.ld per row and head), be exponentiated by ONE thread per row and go out again
// through shared memory for the second product: ~9 k cycles per (tile, head) of serial latency.  The mma.sync kernel keeps scores
// and probabilities in the registers of the warp that produced them.  So this kernel is the OPT-IN second implementation
// (isf_set_option("i2s_tc", 1)); both are tested against each other and against fp32.
#include "tc_gemm.cuh"

extern "C" void isf_count_launch(int n);
int isf_i2s_attention_mma(const void* q, const void* kv, int batch, int tokens, int num_keys, int channels, int heads, void* out,
                          void* stream);      // fusion.cu

namespace i2s {
using namespace tc;

constexpr int C = 128, HEADS = 8, D = 16, NKP = 208, TQ = 128;
constexpr int Q_OFF = 0;                          // 2 chunks x [128 rows x 128 B]
constexpr int K_OFF = Q_OFF + 2 * TQ * 128;       // 2 chunks x [208 rows x 128 B]
constexpr int VT_OFF = K_OFF + 2 * NKP * 128;     // 4 key chunks x [128 channel rows x 128 B]
constexpr int P_OFF = VT_OFF + 4 * C * 128;       // 4 key chunks x [128 query rows x 128 B]
constexpr int BAR_OFF = P_OFF + 4 * TQ * 128;
constexpr int NUM_BARS = 16;
constexpr int SMEM = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
static_assert(SMEM <= 227 * 1024, "i2s_tc: shared memory budget exceeded");
static_assert(K_OFF % 1024 == 0 && VT_OFF % 1024 == 0 && P_OFF % 1024 == 0, "swizzle atoms are 1024-byte aligned");
constexpr int I2S_THREADS = 6 * 32;
// TMEM columns: S buffers [0, 208) and [256, 464), O buffers [208, 224) and [464, 480)
__device__ __forceinline__ uint32_t s_col(int i) { return (uint32_t)(i * 256); }
__device__ __forceinline__ uint32_t o_col(int i) { return (uint32_t)(i * 256 + 208); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"((uint64_t)tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(I2S_THREADS, 1)
i2s_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __half* __restrict__ kv,
              int T, int NK, float sl2, __half* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + BAR_OFF);
  uint64_t* q_full = bars + 0;        // TMA: the tile's Q has landed
  uint64_t* q_empty = bars + 1;       // commit: the 8 QK^T instructions of the tile have read it
  uint64_t* k_full = bars + 2;        // TMA: K of the frame (once)
  uint64_t* s_full = bars + 3;        // [2] commit: S buffer holds head g's scores
  uint64_t* s_free = bars + 5;        // [2] 128 softmax threads have read it
  uint64_t* p_full = bars + 7;        // 128 softmax threads have written P
  uint64_t* p_free = bars + 8;        // commit: the P.V instructions have read P
  uint64_t* o_full = bars + 9;        // [2] commit: O buffer holds head g's output
  uint64_t* o_free = bars + 11;       // [2] 128 threads have read it
  uint32_t* tmem_slot = (uint32_t*)(bars + NUM_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int num_tiles = (T + TQ - 1) / TQ;
  const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_tiles * HEADS;      // (tile, head) steps of this CTA

  if (tid == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(k_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128);
      mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 128);
    }
    mbar_init(p_full, 128); mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // V^T: element (key j, channel c) of the frame's V -> row c, K index j of the 128B-swizzled K-major tile (chunks of 64 keys);
  // consecutive threads take consecutive keys of one 8-channel group, so the 2-byte stores of a warp fall into one row
  {
    const __half* vb = kv + (size_t)b * NK * 2 * C + C;
    for (int e = tid; e < NKP * (C / 8); e += I2S_THREADS) {
      const int j = e % NKP, c8 = (e / NKP) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (j < NK) v = *(const uint4*)(vb + (size_t)j * 2 * C + c8);
      const __half* h = (const __half*)&v;
      uint8_t* base = smem + VT_OFF + (j >> 6) * (C * 128) + (j & 7) * 2;
      const int slot = (j & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ch = c8 + i;
        *(__half*)(base + ch * 128 + ((slot ^ (ch & 7)) << 4)) = h[i];
      }
    }
    // keys 208 .. 255 of the last chunk are never read (13 K steps = 208 keys)
  }
  fence_proxy_async();       // generic-proxy writes of V^T -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(k_full, 2 * NKP * 128);
      tma_load_3d(smem_u32(smem + K_OFF), &tmap_k, 0, 0, b, k_full);
      tma_load_3d(smem_u32(smem + K_OFF + NKP * 128), &tmap_k, 64, 0, b, k_full);
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = blockIdx.x + t * gridDim.x;
        if (t > 0) mbar_wait_b(q_empty, (uint32_t)(t - 1) & 1);
        mbar_expect_tx(q_full, 2 * TQ * 128);
        tma_load_3d(smem_u32(smem + Q_OFF), &tmap_q, 0, tile * TQ, b, q_full);
        tma_load_3d(smem_u32(smem + Q_OFF + TQ * 128), &tmap_q, 64, tile * TQ, b, q_full);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp converged; the wrappers elect)
    constexpr uint32_t idesc_s = make_idesc_f16(TQ, NKP);
    constexpr uint32_t idesc_o = make_idesc_f16(TQ, D);
    if (my_tiles > 0) mbar_wait_b(k_full, 0);
    for (int g = 0; g <= G; ++g) {
      if (g < G) {                       // S[g & 1] = Q_h . K_h^T
        const int h = g & 7, t = g >> 3, sb = g & 1;
        if (h == 0) mbar_wait_b(q_full, (uint32_t)t & 1);
        if (g >= 2) mbar_wait_b(&s_free[sb], (uint32_t)((g >> 1) - 1) & 1);
        tc_fence_after();
        const uint64_t a_desc = make_desc_sw128(smem_u32(smem + Q_OFF + (h >> 2) * (TQ * 128))) + (uint64_t)((h & 3) * 2);
        const uint64_t b_desc = make_desc_sw128(smem_u32(smem + K_OFF + (h >> 2) * (NKP * 128))) + (uint64_t)((h & 3) * 2);
        umma_f16(tmem + s_col(sb), a_desc, b_desc, idesc_s, 0u);
        umma_commit(&s_full[sb]);
        if (h == 7) umma_commit(q_empty);
      }
      if (g >= 1) {                      // O[(g-1) & 1] = P . V_h
        const int gp = g - 1, h = gp & 7, ob = gp & 1;
        mbar_wait_b(p_full, (uint32_t)gp & 1);
        if (gp >= 2) mbar_wait_b(&o_free[ob], (uint32_t)((gp >> 1) - 1) & 1);
        fence_proxy_async();
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < NKP / 16; ++ks) {
          const uint64_t a_desc = make_desc_sw128(smem_u32(smem + P_OFF + (ks >> 2) * (TQ * 128))) + (uint64_t)((ks & 3) * 2);
          const uint64_t b_desc = make_desc_sw128(smem_u32(smem + VT_OFF + (ks >> 2) * (C * 128) + h * D * 128)) + (uint64_t)((ks & 3) * 2);
          umma_f16(tmem + o_col(ob), a_desc, b_desc, idesc_o, ks ? 1u : 0u);
        }
        umma_commit(&o_full[ob]);
        umma_commit(p_free);
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax + epilogue (warps 2-5; TMEM lane quarter = warp % 4)
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* prow = smem + P_OFF + r * 128;
    float l_prev = 1.f;
    for (int g = 0; g <= G; ++g) {
      float l = 0.f;
      if (g < G) {
        const int sb = g & 1;
        mbar_wait_b(&s_full[sb], (uint32_t)(g >> 1) & 1);
        tc_fence_after();
        const uint32_t srow = lane_base + s_col(sb);
        float m = -INFINITY;
#pragma unroll 1
        for (int cb = 0; cb < NKP; cb += 16) {
          uint32_t v[16];
          tmem_ld16(srow + cb, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cb + j < NK) m = fmaxf(m, __uint_as_float(v[j]));
        }
        const float ms = m * sl2;
        if (g >= 1) mbar_wait_b(p_free, (uint32_t)(g - 1) & 1);        // P.V of the previous head has read P
#pragma unroll 1
        for (int cb = 0; cb < NKP; cb += 16) {
          uint32_t v[16];
          tmem_ld16(srow + cb, v);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = (cb + 2 * j < NK) ? exp2f(fmaf(__uint_as_float(v[2 * j]), sl2, -ms)) : 0.f;
            const float p1 = (cb + 2 * j + 1 < NK) ? exp2f(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -ms)) : 0.f;
            // the sum runs over the fp16-rounded probabilities the tensor core will multiply
            const __half2 hp = __floats2half2_rn(p0, p1);
            const float2 f = __half22float2(hp);
            l += f.x + f.y;
            pk[j] = *(const uint32_t*)&hp;
          }
          uint8_t* pc = prow + (cb >> 6) * (TQ * 128);
          const int slot = (cb & 63) >> 3;
          *(uint4*)(pc + (((slot) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *(uint4*)(pc + (((slot + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        tc_fence_before();
        mbar_arrive(&s_free[sb]);
        fence_proxy_async();                 // this thread's P row -> visible to the tensor core
        mbar_arrive(p_full);
      }
      if (g >= 1) {                          // output of the previous head
        const int gp = g - 1, h = gp & 7, ob = gp & 1;
        const int tile = blockIdx.x + (gp >> 3) * gridDim.x;
        mbar_wait_b(&o_full[ob], (uint32_t)(gp >> 1) & 1);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(lane_base + o_col(ob), v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_free[ob]);
        const float inv = 1.f / l_prev;
        const int row = tile * TQ + r;
        if (row < T) {
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __half2 hp = __floats2half2_rn(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
            w[j] = *(const uint32_t*)&hp;
          }
          uint4* op = (uint4*)(out + ((size_t)b * T + row) * C + h * D);
          op[0] = make_uint4(w[0], w[1], w[2], w[3]);
          op[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      l_prev = l;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

static int make_tmap3(CUtensorMap* tm, const void* base, int cols, int rows, int batch, long long row_pitch_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { isf_set_error("cuTensorMapEncodeTiled not available"); return ISF_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)row_pitch_elems * 2, (cuuint64_t)row_pitch_elems * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("i2s_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return ISF_ERR_CUDA; }
  return ISF_OK;
}

}  // namespace i2s

// see include/isfusion_b200.h
extern "C" int isf_i2s_attention(const void* q, const void* kv, int batch, int tokens, int num_keys, int channels,
                                 int heads, void* out, void* stream) {
  ISF_CHECK_ARG(q && kv && out && batch >= 1 && tokens >= 1 && num_keys >= 1);
  if (channels != 128 || heads != 8 || num_keys > 208) {
    isf_set_error("i2s_attention: built for d_model 128, 8 heads, <= 208 keys");
    return ISF_ERR_UNSUPPORTED;
  }
  if (!g_isf_i2s_tc || ((uintptr_t)q & 15) || ((uintptr_t)kv & 15) || ((uintptr_t)out & 15))
    return isf_i2s_attention_mma(q, kv, batch, tokens, num_keys, channels, heads, out, stream);
  using namespace i2s;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { isf_set_error("i2s_tc: bad device"); return ISF_ERR_CUDA; }
  static bool configured[64] = {};
  if (!configured[dev]) {
    ISF_CUDA(cudaFuncSetAttribute(i2s_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured[dev] = true;
  }
  CUtensorMap tq, tk;
  memset(&tq, 0, sizeof(tq));
  memset(&tk, 0, sizeof(tk));
  int rc = make_tmap3(&tq, q, C, tokens, batch, C, TQ);
  if (rc) return rc;
  rc = make_tmap3(&tk, kv, C, num_keys, batch, 2 * C, NKP);        // the K half of every [k | v] row
  if (rc) return rc;
  const int tiles = (tokens + TQ - 1) / TQ;
  int gx = isf_num_sms() / batch;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  i2s_tc_kernel<<<dim3(gx, batch), I2S_THREADS, SMEM, (cudaStream_t)stream>>>(tq, tk, (const __half*)kv, tokens, num_keys,
                                                                           (1.0f / sqrtf((float)D)) * 1.44269504088896f, (__half*)out);
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}
