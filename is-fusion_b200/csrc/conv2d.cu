// Dense 3x3 (padding 1, stride 1 | 2) convolution of a channels-last fp16 BEV map on tcgen05, with TMA-tiled halos.
//
//   out[b, oy, ox, n] = act( shift[n] + sum_{dy,dx,c} x[b, oy*s + dy - 1, ox*s + dx - 1, c] * W[n, (dy*3 + dx)*Cin + c] )
//
// The neck's dense convolutions (reference fusion_encoder.py:862-869,915-960, backbones/second.py:127-167: conv_fusion
// 768->128, the heat-map / scene / instance convs, SECONDV2 — 203 GFLOP per frame) as ONE implicit-GEMM kernel with no
// gather at all: a CTA owns an output tile of (4*MT) x 32 pixels (MT sub-tiles of M = 128 rows, one TMEM accumulator each)
// and all BN output channels of its column block.
//
//   * A operand = the input halo, moved by TMA.  stride 1: one 4-D box {64 ch, 32 x, 4*MT + 2 y, 1} per (channel chunk,
//     dx) lands as (4*MT+2)*32 rows of 128 B in the 128B-swizzled K-major layout of the UMMA descriptor; the three dy
//     taps re-use that box from shared memory (a dy shift = 32 rows = 4096 B, so the descriptors stay 1024-B aligned) —
//     3 loads instead of 9 per chunk, and the conv's zero padding is the TMA's out-of-bounds fill (negative coordinates
//     included).  stride 2: the map is viewed as [B, H/2, 2, W/2, 2*C] (row / column parity folded into two tensor
//     dimensions), which turns every tap into a plain 5-D box {64, 32, 1, 4*MT, 1}: no element strides, no gather.
//   * B operand = the packed weights [Cout_pad, 9*Cin] K-major, one 2-D box {64, BN} per (tap, chunk), own ring; one weight
//     box feeds the MT sub-tiles.
//   * warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane: tcgen05.mma M=128 N=BN K=16, tcgen05.commit
//     recycles the rings), warps 2-5 = epilogue: TMEM -> registers -> + folded-BN shift, ReLU -> fp16 -> swizzled
//     staging tile in shared memory -> ONE TMA store per (sub-tile, 64 channels) (cp.async.bulk.tensor .global.shared:
//     UTMASTG; the store clips pixels beyond the map and channels beyond cout).  Column blocks narrower than 64 (the
//     10-class heat-map conv) store straight from registers.
//   * persistent over (batch, tile) with the accumulator handed back through an mbarrier.
#include "tc_gemm.cuh"

extern "C" void isf_count_launch(int n);

namespace dc {
using namespace tc;

constexpr int TW = 32;                      // output pixels per tile row
constexpr int SUB_ROWS = 4;                 // tile rows per M = 128 sub-tile
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 128;
constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS;
constexpr int ROW_BYTES = TW * 128;         // one tile row of one 64-channel chunk in shared memory

template <int BN, int MT, int SA, int SB>
struct Smem {
  static constexpr int A_BYTES = (SUB_ROWS * MT + 2) * ROW_BYTES;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = SA * A_BYTES;
  static constexpr int C_OFF = B_OFF + SB * B_BYTES;              // 2 staging tiles [128 rows][128 B] for the TMA stores
  static constexpr int VEC_OFF = C_OFF + 2 * BM * 128;             // shift[BN] fp32
  static constexpr int BAR_OFF = VEC_OFF + BN * 4;
  static constexpr int NUM_BARS = 2 * SA + 2 * SB + 2;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
  static_assert(TOTAL <= 227 * 1024, "conv3x3: shared memory budget exceeded");
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"((uint64_t)tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"((uint64_t)tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)tmap),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

struct ConvParams {
  int batch, oh, ow;          // output map
  int cin;                    // multiple of 64
  int ldx;                    // input pixel pitch in elements (stride 2: channel offset of the odd pixel of a column pair)
  int n_total;                // padded output channels (rows of W)
  int cout_store;             // channels written (multiple of 8)
  const float* shift;         // [n_total] or nullptr
  int relu;
  __half* out;                // [batch, oh, ow, ldo]
  int ldo;
  int bf16;                   // operands / output are bf16 (isf_conv3x3_nhwc_bf16), else fp16
};

// STRIDE 1: 3 A loads per channel chunk (dx), 3 taps (dy) per load.  STRIDE 2: 9 loads, 1 tap each.
template <int BN, int MT, int SA, int SB, int STRIDE, bool TMA_STORE>
__global__ void __launch_bounds__(THREADS, 1)
conv3x3_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_o) {
  using SL = Smem<BN, MT, SA, SB>;
  constexpr int NL = STRIDE == 1 ? 3 : 9;          // A loads per chunk
  constexpr int TPL = STRIDE == 1 ? 3 : 1;         // taps per load
  constexpr int A_LOAD_BYTES = (STRIDE == 1 ? SUB_ROWS * MT + 2 : SUB_ROWS * MT) * ROW_BYTES;
  constexpr int ACC_COLS = MT * BN;
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  static_assert(ACC_COLS <= 512, "TMEM holds 512 columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = (uint64_t*)(smem + SL::BAR_OFF);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* acc_full = b_empty + SB;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 1);
  float* vec = (float*)(smem + SL::VEC_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_x = (p.ow + TW - 1) / TW;
  const int tiles_y = (p.oh + SUB_ROWS * MT - 1) / (SUB_ROWS * MT);
  const int num_tiles = p.batch * tiles_y * tiles_x;
  const int n0 = blockIdx.y * BN;
  const int chunks = p.cin / BK;

  if (tid == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int c = tid; c < BN; c += THREADS) vec[c] = p.shift ? p.shift[n0 + c] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      uint32_t ca = 0, cb = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / (tiles_y * tiles_x);
        const int ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
        const int y0 = ty * SUB_ROWS * MT, x0 = tx * TW;
        for (int c = 0; c < chunks; ++c) {
#pragma unroll 1
          for (int l = 0; l < NL; ++l, ++ca) {
            const int sa = (int)(ca % SA);
            const uint32_t ua = ca / SA;
            if (ua > 0) mbar_wait_b(&a_empty[sa], (ua - 1) & 1);
            mbar_expect_tx(&a_full[sa], A_LOAD_BYTES);
            const uint32_t a_dst = smem_u32(smem + SL::A_OFF + sa * SL::A_BYTES);
            if constexpr (STRIDE == 1) {
              tma_load_4d(a_dst, &tmap_x, c * BK, x0 + l - 1, y0 - 1, b, &a_full[sa]);
            } else {
              const int dy = l / 3, dx = l % 3;      // input y = 2*oy + dy - 1 -> (pair oy - (dy == 0), parity dy != 1)
              tma_load_5d(a_dst, &tmap_x, (dx != 1 ? p.ldx : 0) + c * BK, x0 - (dx == 0), dy != 1, y0 - (dy == 0), b, &a_full[sa]);
            }
#pragma unroll 1
            for (int t = 0; t < TPL; ++t, ++cb) {
              const int sb = (int)(cb % SB);
              const uint32_t ub = cb / SB;
              if (ub > 0) mbar_wait_b(&b_empty[sb], (ub - 1) & 1);
              const int tap = STRIDE == 1 ? t * 3 + l : l;
              mbar_expect_tx(&b_full[sb], SL::B_BYTES);
              tma_load_2d(smem_u32(smem + SL::B_OFF + sb * SL::B_BYTES), &tmap_w, tap * p.cin + c * BK, n0, &b_full[sb]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc_f16(BM, BN) | (p.bf16 ? tc::IDESC_AB_BF16 : 0u);
    uint32_t ca = 0, cb = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (it > 0) mbar_wait_b(acc_empty, (uint32_t)(it - 1) & 1);      // the epilogue has drained the accumulators
      tc_fence_after();
      for (int c = 0; c < chunks; ++c) {
#pragma unroll 1
        for (int l = 0; l < NL; ++l, ++ca) {
          const int sa = (int)(ca % SA);
          mbar_wait_b(&a_full[sa], (ca / SA) & 1);
          const uint32_t a_addr = smem_u32(smem + SL::A_OFF + sa * SL::A_BYTES);
#pragma unroll 1
          for (int t = 0; t < TPL; ++t, ++cb) {
            const int sb = (int)(cb % SB);
            mbar_wait_b(&b_full[sb], (cb / SB) & 1);
            tc_fence_after();
            {      // whole warp, converged: one election per (tap, chunk) (umma_chunk, tc_gemm.cuh); sub-tiles alternate per K step
              // sub-tile mt, tap row shift t (stride 1: dy rows of the halo box): a multiple of 4096 B
              const uint32_t a0 = a_addr + (uint32_t)((STRIDE == 1 ? t : 0) * ROW_BYTES);
              umma_chunk<1, MT>(tmem_base, tmem_base + (uint32_t)BN, make_desc_sw128(a0), make_desc_sw128(a0 + (uint32_t)(SUB_ROWS * ROW_BYTES)),
                                make_desc_sw128(smem_u32(smem + SL::B_OFF + sb * SL::B_BYTES)), idesc, (c | l | t) ? 1u : 0u, &b_empty[sb]);
            }
            __syncwarp();
          }
          umma_commit(&a_empty[sa]);
          if (c == chunks - 1 && l == NL - 1) umma_commit(acc_full);
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2-5; TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row of the sub-tile = pixel (r / 32, r % 32)
    const int etid = tid - EPI_WARP0 * 32;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    uint32_t sidx = 0;                           // staging tile counter
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / (tiles_y * tiles_x);
      const int ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
      const int y0 = ty * SUB_ROWS * MT, x0 = tx * TW;
      mbar_wait_b(acc_full, (uint32_t)it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        if constexpr (TMA_STORE) {
#pragma unroll 1
          for (int cc = 0; cc < BN / 64; ++cc, ++sidx) {
            uint8_t* stage = smem + SL::C_OFF + (sidx & 1) * (BM * 128);
            if (etid == 0) bulk_wait_read<1>();      // the store issued from this staging tile two rounds ago has read it
            epi_bar();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld32(t_lane + (uint32_t)(mt * BN + cc * 64 + h * 32), v);
              tmem_ld_wait();
              const float* sh = vec + cc * 64 + h * 32;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                uint4 u;
                uint32_t* hp = (uint32_t*)&u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a0 = __uint_as_float(v[j4 * 8 + e * 2]) + sh[j4 * 8 + e * 2];
                  float a1 = __uint_as_float(v[j4 * 8 + e * 2 + 1]) + sh[j4 * 8 + e * 2 + 1];
                  if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                  hp[e] = isf_pack2(a0, a1, p.bf16 != 0);
                }
                // 16-byte piece (h*4 + j4) of row r in the 128B-swizzled tile the store's tensor map describes
                *(uint4*)(stage + r * 128 + (((h * 4 + j4) ^ (r & 7)) << 4)) = u;
              }
            }
            fence_proxy_async();                     // generic-proxy writes -> visible to the TMA (async proxy)
            epi_bar();
            if (etid == 0) {
              tma_store_4d(&tmap_o, smem_u32(stage), n0 + cc * 64, x0, y0 + mt * SUB_ROWS, b);
              bulk_commit();
            }
          }
        } else {
          const int oy = y0 + mt * SUB_ROWS + (r >> 5), ox = x0 + (r & 31);
          const bool ok = oy < p.oh && ox < p.ow;
          __half* orow = p.out + (((long long)b * p.oh + oy) * p.ow + ox) * p.ldo + n0;
#pragma unroll 1
          for (int cb = 0; cb < BN; cb += 32) {
            uint32_t v[32];
            tmem_ld32(t_lane + (uint32_t)(mt * BN + cb), v);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint4 u;
              uint32_t* hp = (uint32_t*)&u;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a0 = __uint_as_float(v[j4 * 8 + e * 2]) + vec[cb + j4 * 8 + e * 2];
                float a1 = __uint_as_float(v[j4 * 8 + e * 2 + 1]) + vec[cb + j4 * 8 + e * 2 + 1];
                if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                hp[e] = isf_pack2(a0, a1, p.bf16 != 0);
              }
              if (ok && n0 + cb + j4 * 8 < p.cout_store) *(uint4*)(orow + cb + j4 * 8) = u;
            }
          }
        }
      }
      // this thread's TMEM reads are complete (tcgen05.wait::ld) and ordered before the hand-back
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
    if (TMA_STORE && etid == 0) bulk_wait_read<0>();   // staging tiles stay valid until the last store has read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

static int encode_tmap(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { isf_set_error("cuTensorMapEncodeTiled not available"); return ISF_ERR_CUDA; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("conv3x3: cuTensorMapEncodeTiled failed (%d, rank %d)", (int)r, rank); return ISF_ERR_CUDA; }
  return ISF_OK;
}

struct ConvArgs {
  const void* x; int batch, h, w, cin, ldx;
  const void* wpk; int n_total;
  int stride;
  ConvParams p;
};

template <int BN, int MT, int SA, int SB, int STRIDE>
static int launch(const ConvArgs& a, cudaStream_t st) {
  using SL = Smem<BN, MT, SA, SB>;
  constexpr bool TMA_STORE = BN >= 64;
  auto kern = conv3x3_kernel<BN, MT, SA, SB, STRIDE, TMA_STORE>;
  int dev = 0;
  ISF_CUDA(cudaGetDevice(&dev));
  static bool configured[64] = {};
  if (dev < 0 || dev >= 64) { isf_set_error("conv3x3: device ordinal %d out of range", dev); return ISF_ERR_INVALID; }
  if (!configured[dev]) {
    ISF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL));
    configured[dev] = true;
  }
  CUtensorMap tx, tw, to;
  memset(&tx, 0, sizeof(tx)); memset(&tw, 0, sizeof(tw)); memset(&to, 0, sizeof(to));
  int rc;
  if (STRIDE == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)a.cin, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.batch};
    cuuint64_t str[3] = {(cuuint64_t)a.ldx * 2, (cuuint64_t)a.w * a.ldx * 2, (cuuint64_t)a.h * a.w * a.ldx * 2};
    cuuint32_t box[4] = {64, TW, SUB_ROWS * MT + 2, 1};
    rc = encode_tmap(&tx, a.x, 4, dims, str, box);
  } else {
    // [B, H/2, 2, W/2, 2*C]: channel' = x-parity * ldx + c (the two pixels of a column pair are adjacent in memory)
    cuuint64_t dims[5] = {(cuuint64_t)a.ldx + a.cin, (cuuint64_t)a.w / 2, 2, (cuuint64_t)a.h / 2, (cuuint64_t)a.batch};
    cuuint64_t str[4] = {(cuuint64_t)a.ldx * 4, (cuuint64_t)a.w * a.ldx * 2, (cuuint64_t)a.w * a.ldx * 4,
                         (cuuint64_t)a.h * a.w * a.ldx * 2};
    cuuint32_t box[5] = {64, TW, 1, SUB_ROWS * MT, 1};
    rc = encode_tmap(&tx, a.x, 5, dims, str, box);
  }
  if (rc) return rc;
  rc = make_weight_tmap(&tw, a.wpk, a.n_total, 9 * a.cin, BN);
  if (rc) return rc;
  if (TMA_STORE) {
    cuuint64_t dims[4] = {(cuuint64_t)a.p.cout_store, (cuuint64_t)a.p.ow, (cuuint64_t)a.p.oh, (cuuint64_t)a.batch};
    cuuint64_t str[3] = {(cuuint64_t)a.p.ldo * 2, (cuuint64_t)a.p.ow * a.p.ldo * 2, (cuuint64_t)a.p.oh * a.p.ow * a.p.ldo * 2};
    cuuint32_t box[4] = {64, TW, SUB_ROWS, 1};
    rc = encode_tmap(&to, a.p.out, 4, dims, str, box);
    if (rc) return rc;
  }
  int sms = 0;
  ISF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles = a.batch * isf_cdiv(a.p.oh, SUB_ROWS * MT) * isf_cdiv(a.p.ow, TW);
  const int n_blocks = a.n_total / BN;
  int gx = sms / n_blocks;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  kern<<<dim3(gx, n_blocks), THREADS, SL::TOTAL, st>>>(a.p, tx, tw, to);
  ISF_LAUNCH_CHECK();
  isf_count_launch(1);
  return ISF_OK;
}

template <int STRIDE>
static int dispatch(const ConvArgs& a, cudaStream_t st) {
  // tile shape: two sub-tiles (8 x 32 pixels, one weight box feeds both) while that still fills the machine,
  // else one (4 x 32); 256 output channels on a small map split into two column blocks of 128
  const int tiles2 = a.batch * isf_cdiv(a.p.oh, 8) * isf_cdiv(a.p.ow, TW);
  const bool two = tiles2 >= 120;
  switch (a.n_total) {
    case 32: return two ? launch<32, 2, 4, 6, STRIDE>(a, st) : launch<32, 1, 4, 6, STRIDE>(a, st);
    case 64: return two ? launch<64, 2, 3, 6, STRIDE>(a, st) : launch<64, 1, 4, 6, STRIDE>(a, st);
    case 128: return two ? launch<128, 2, 3, 4, STRIDE>(a, st) : launch<128, 1, 4, 5, STRIDE>(a, st);
    case 256:
      if (two) return launch<256, 2, 2, 3, STRIDE>(a, st);
      if (tiles2 * 2 < 120) return launch<128, 1, 4, 5, STRIDE>(a, st);      // column blocks of 128 -> twice the CTAs
      return launch<256, 1, 3, 3, STRIDE>(a, st);
    default: isf_set_error("conv3x3: cout_pad=%d not in {32,64,128,256}", a.n_total); return ISF_ERR_UNSUPPORTED;
  }
}

}  // namespace dc

// see include/isfusion_b200.h
extern "C" int isf_conv3x3_nhwc_f16(const void* x, int batch, int h, int w, int cin, int ldx, const void* w_packed, int cout_pad,
                                    const float* shift, int stride, int relu, void* out, int ldo, int cout_store,
                                    void* stream) {
  ISF_CHECK_ARG(batch >= 0 && h > 0 && w > 0 && cin > 0 && cin % 64 == 0 && ldx >= cin && ldx % 8 == 0);
  ISF_CHECK_ARG(stride == 1 || (stride == 2 && h % 2 == 0 && w % 2 == 0));
  ISF_CHECK_ARG(ldo % 8 == 0 && cout_store > 0 && cout_store % 8 == 0 && cout_store <= cout_pad && cout_store <= ldo);
  if (batch == 0) return ISF_OK;
  ISF_CHECK_ARG(x && w_packed && out);
  ISF_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w_packed & 15) == 0);
  dc::ConvArgs a{};
  a.x = x; a.batch = batch; a.h = h; a.w = w; a.cin = cin; a.ldx = ldx; a.wpk = w_packed; a.n_total = cout_pad; a.stride = stride;
  a.p.batch = batch; a.p.oh = (h + 2 - 3) / stride + 1; a.p.ow = (w + 2 - 3) / stride + 1; a.p.cin = cin; a.p.ldx = ldx; a.p.n_total = cout_pad;
  a.p.cout_store = cout_store; a.p.shift = shift; a.p.relu = relu; a.p.out = (__half*)out; a.p.ldo = ldo;
  a.p.bf16 = isf_bf16_mode() ? 1 : 0;
  return stride == 1 ? dc::dispatch<1>(a, (cudaStream_t)stream) : dc::dispatch<2>(a, (cudaStream_t)stream);
}

// bf16 twin of isf_conv3x3_nhwc_f16 (x, w_packed and out are bf16; the TMA boxes move 2-byte elements either way)
extern "C" int isf_conv3x3_nhwc_bf16(const void* x, int batch, int h, int w, int cin, int ldx, const void* w_packed, int cout_pad,
                                     const float* shift, int stride, int relu, void* out, int ldo, int cout_store,
                                     void* stream) {
  IsfBf16Scope scope;
  return isf_conv3x3_nhwc_f16(x, batch, h, w, cin, ldx, w_packed, cout_pad, shift, stride, relu, out, ldo, cout_store, stream);
}
