// Sparse-convolution implicit GEMM with the A rows gathered by the TMA (tile::gather4) — "v3" schedule.
//
// Same contract as tc::gemm_kernel with a neighbour table (tc_gemm.cuh):  D[m, n] = epi(sum_k A[nbr[k / Cin][m], k % Cin] W[n, k]).
// What the round-2 timing experiments showed for the LDGSTS kernel (profiles/call2_r02.log): removing ALL A copies buys 16 %,
// removing the weight loads 5 %, while the issued MMAs account for ~40 % of the time — the rest is the per-chunk handshake of
// 256 producer threads and the tile prologue / epilogue that only co-resident CTAs hide.  Here nothing but the TMA moves data:
//
//   warp 0      producer: per active K chunk (kernel offset, 64 channels) every lane issues ONE
//               cp.async.bulk.tensor.2d.tile::gather4 — four neighbour rows x 128 B straight into the 128B-swizzled UMMA
//               layout (missing neighbours = row index -1 = out of bounds = zero fill) — so one warp instruction brings a whole
//               128-row sub-tile; lane 0 adds the weight box and the expect_tx.  No LSU loads, no address arithmetic per piece.
//   warp 1      MMA issuer (tcgen05.mma M=128 N=BN K=16, MT sub-tiles per weight chunk), tcgen05.commit recycles the stages.
//   warps 2-5   epilogue on a DOUBLE-BUFFERED TMEM accumulator: tile t drains while tile t+1 accumulates.
//   warps 6-9   index warps: neighbour tile of the NEXT tile -> shared memory (double-buffered), active-chunk list.
//
// One persistent CTA per SM; every mbarrier wait is bounded and traps.
#pragma once
#include "tc_gemm.cuh"

namespace tc {

constexpr int V3_EPI_WARP0 = 2;
constexpr int V3_IDX_WARP0 = 6;
constexpr int V3_THREADS = 10 * 32;

template <int BN, int MT, int S>
struct Smem3 {
  static constexpr int TILE_M = MT * BM;
  static constexpr int A_BYTES = TILE_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NBR_OFF = S * STAGE_BYTES;                         // 2 x [MAX_TAPS][TILE_M] int32 row indices
  static constexpr int LIST_OFF = NBR_OFF + 2 * MAX_TAPS * TILE_M * 4;    // 2 x chunk list (uint16)
  static constexpr int FLAG_OFF = LIST_OFF + 2 * MAX_CHUNKS * 2;          // 2 x (32 tap flags + count + pad)
  static constexpr int VEC_OFF = FLAG_OFF + 2 * 64 * 4;                   // scale[BN], shift[BN] fp32
  static constexpr int BAR_OFF = VEC_OFF + 2 * BN * 4;
  static constexpr int NUM_BARS = 2 * S + 8;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
  static_assert(TOTAL <= 227 * 1024, "v3 sparse conv: shared memory budget exceeded");
};

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"((uint64_t)tm), "r"(col), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3), "r"(smem_u32(bar))
               : "memory");
}

// one output row (TMEM lane) of a finished sub-tile: scale / shift (+ residual) (+ activation) -> global (row scatter optional)
template <int BN>
__device__ __forceinline__ void epilogue_row3(const GemmParams& p, uint32_t t_lane, int n0, long long rrow, long long orow,
                                              bool store_ok, const float* __restrict__ vec) {
#pragma unroll 1
  for (int cb = 0; cb < BN; cb += 32) {
    uint32_t v[32];
    tmem_ld32(t_lane + cb, v);
    tmem_ld_wait();
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) * vec[cb + j] + vec[BN + cb + j];
    if (p.residual && rrow >= 0) {
      const uint4* rp = (const uint4*)(p.residual + rrow * p.ldr + n0 + cb);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const uint4 u = rp[j4];
        const __half2* h = (const __half2*)&u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          x[j4 * 8 + e * 2] += f.x; x[j4 * 8 + e * 2 + 1] += f.y;
        }
      }
    }
    if (p.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
    } else if (p.act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
    }
    if (store_ok) {
      if (p.out) {
        uint4* op = (uint4*)(p.out + orow * p.ldo + n0 + cb);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 u;
          __half2* h = (__half2*)&u;
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1]);
          op[j4] = u;
        }
      }
      if (p.out_f32) {
        float4* op = (float4*)(p.out_f32 + orow * p.ldo + n0 + cb);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) op[j4] = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
      }
    }
  }
}

template <int BN, int MT, int S>
__global__ void __launch_bounds__(V3_THREADS, 1)
spconv3_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w) {
  using SL = Smem3<BN, MT, S>;
  constexpr int TILE_M = SL::TILE_M;
  constexpr int ACC_COLS = 2 * MT * BN;                                // double-buffered accumulators
  static_assert(ACC_COLS <= 512, "TMEM holds 512 columns");
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = (uint64_t*)(smem + SL::BAR_OFF);
  uint64_t* empty = full + S;
  uint64_t* acc_full = empty + S;       // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint64_t* list_full = acc_empty + 2;  // [2]
  uint64_t* list_empty = list_full + 2; // [2]
  uint32_t* tmem_slot = (uint32_t*)(list_empty + 2);
  float* vec = (float*)(smem + SL::VEC_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_total = p.m_dev ? *p.m_dev : p.m;
  const int num_tiles = (m_total + TILE_M - 1) / TILE_M;
  const int n0 = blockIdx.y * BN;
  const int kc = p.cin / BK;                       // 64-channel chunks per kernel offset
  const int num_chunks = p.taps * kc;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
      mbar_init(&list_full[b], 1);
      mbar_init(&list_empty[b], 2);                // producer + MMA issuer
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int c = tid; c < BN; c += V3_THREADS) {
    vec[c] = p.scale ? p.scale[n0 + c] : 1.f;
    vec[BN + c] = p.shift ? p.shift[n0 + c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= V3_IDX_WARP0) {
    // ================================================================ index warps: neighbour tile + chunk list per tile
    const int r = tid - V3_IDX_WARP0 * 32;          // 0..127
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      int* nbr_s = (int*)(smem + SL::NBR_OFF) + buf * (MAX_TAPS * TILE_M);
      uint16_t* chunk_list = (uint16_t*)(smem + SL::LIST_OFF) + buf * MAX_CHUNKS;
      int* tap_flag = (int*)(smem + SL::FLAG_OFF) + buf * 64;
      if (ubuf > 0) mbar_wait_b(&list_empty[buf], (ubuf - 1) & 1);
      if (r < 33) tap_flag[r] = 0;
      asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const long long prow = (long long)tile * TILE_M + mt * BM + r;
        const bool pok = prow < m_total;
        int nv[MAX_TAPS];
#pragma unroll
        for (int t = 0; t < MAX_TAPS; ++t) nv[t] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
#pragma unroll
        for (int t = 0; t < MAX_TAPS; ++t) {
          if (t < p.taps) {
            nbr_s[t * TILE_M + mt * BM + r] = nv[t];
            if (nv[t] >= 0) tap_flag[t] = 1;
          }
        }
      }
      asm volatile("bar.sync 3, 128;" ::: "memory");
      if (warp == V3_IDX_WARP0) {
        int count = 0;
        for (int base = 0; base < num_chunks; base += 32) {
          const int q = base + lane;
          const bool act = q < num_chunks && tap_flag[q / kc] != 0;
          const unsigned mask = __ballot_sync(0xffffffffu, act);
          if (act) chunk_list[count + __popc(mask & ((1u << lane) - 1))] = (uint16_t)q;
          count += __popc(mask);
        }
        if (count == 0) { if (lane == 0) chunk_list[0] = 0; count = 1; }    // keeps the accumulator defined
        if (lane == 0) tap_flag[32] = count;
        __syncwarp();
        if (lane == 0) mbar_arrive(&list_full[buf]);
      }
    }
  } else if (warp == 0) {
    // ================================================================ producer: TMA gathers + weight boxes
    uint32_t ca = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      const int* nbr_s = (const int*)(smem + SL::NBR_OFF) + buf * (MAX_TAPS * TILE_M);
      const uint16_t* chunk_list = (const uint16_t*)(smem + SL::LIST_OFF) + buf * MAX_CHUNKS;
      mbar_wait_b(&list_full[buf], ubuf & 1);
      const int n_act = ((const volatile int*)(smem + SL::FLAG_OFF))[buf * 64 + 32];
      for (int i = 0; i < n_act; ++i, ++ca) {
        const int q = chunk_list[i];
        const int tap = q / kc, kchunk = q - tap * kc;
        const int s = (int)(ca % S);
        const uint32_t u = ca / S;
        if (u > 0) mbar_wait_b(&empty[s], (u - 1) & 1);
        const uint32_t a_dst = smem_u32(smem + s * SL::STAGE_BYTES);
        if (lane == 0) {
          mbar_expect_tx(&full[s], SL::STAGE_BYTES);
          tma_load_2d(a_dst + SL::A_BYTES, &tmap_w, q * BK, n0, &full[s]);
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int4 idx = *(const int4*)(nbr_s + tap * TILE_M + mt * BM + 4 * lane);
          tma_gather4(a_dst + mt * (BM * 128) + lane * 512, &tmap_a, kchunk * BK, idx.x, idx.y, idx.z, idx.w, &full[s]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&list_empty[buf]);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(BM, BN);
    uint32_t ca = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      mbar_wait_b(&list_full[buf], ubuf & 1);
      const int n_act = ((const volatile int*)(smem + SL::FLAG_OFF))[buf * 64 + 32];
      __syncwarp();
      if (lane == 0) mbar_arrive(&list_empty[buf]);
      if (ubuf > 0) mbar_wait_b(&acc_empty[buf], (ubuf - 1) & 1);      // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * MT * BN);
      for (int q = 0; q < n_act; ++q, ++ca) {
        const int s = (int)(ca % S);
        mbar_wait_b(&full[s], (ca / S) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES);
          const uint64_t b_desc = make_desc_sw128(a_addr + SL::A_BYTES);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t a_desc = make_desc_sw128(a_addr + (uint32_t)(mt * BM * 128));
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16(d_tmem + (uint32_t)(mt * BN), a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (q | kk) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (q == n_act - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================ epilogue (warps 2-5; TMEM lane quarter = warp % 4)
    const int r = (warp & 3) * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      long long orow[MT], rrow[MT];
      bool sok[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {      // row maps fetched before the wait: their latency overlaps the main loop
        const int grow = tile * TILE_M + mt * BM + r;
        const bool row_ok = grow < m_total;
        orow[mt] = grow;
        if (p.out_rows && row_ok) orow[mt] = p.out_rows[grow];
        sok[mt] = row_ok && orow[mt] >= 0;
        rrow[mt] = row_ok ? (long long)grow : -1;
        if (p.res_rows && row_ok) rrow[mt] = p.res_rows[grow];
      }
      mbar_wait_b(&acc_full[buf], ubuf & 1);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((buf * MT + mt) * BN);
        epilogue_row3<BN>(p, t_lane, n0, rrow[mt], orow[mt], sok[mt], vec);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BN, int MT, int S>
static int launch_spconv3(const GemmParams& p, int m_upper, cudaStream_t st) {
  using SL = Smem3<BN, MT, S>;
  auto kern = spconv3_kernel<BN, MT, S>;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { isf_set_error("spconv v3: bad device"); return ISF_ERR_CUDA; }
  static bool configured[64] = {};
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
    if (e != cudaSuccess) { isf_set_error("cudaFuncSetAttribute (v3): %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
    configured[dev] = true;
  }
  CUtensorMap tw, ta;
  memset(&tw, 0, sizeof(tw));
  memset(&ta, 0, sizeof(ta));
  int rc = make_weight_tmap(&tw, p.W, p.n_total, p.kpad, BN);
  if (rc) return rc;
  {   // feature matrix [rows, Cin] as a 2-D tensor, box = one row x 64 channels (tile::gather4 fetches four such rows).  The row
      // bound is not known to this entry point: any index the neighbour table holds is a valid row, -1 is out of bounds.
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { isf_set_error("cuTensorMapEncodeTiled not available"); return ISF_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)p.cin, (cuuint64_t)1 << 30};
    cuuint64_t strides[1] = {(cuuint64_t)p.cin * 2};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(p.A), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { isf_set_error("spconv v3: cuTensorMapEncodeTiled (A) failed (%d)", (int)r); return ISF_ERR_CUDA; }
  }
  int tiles = (m_upper + SL::TILE_M - 1) / SL::TILE_M;
  if (tiles < 1) tiles = 1;
  const int n_tiles = p.n_total / BN;
  int gx = isf_num_sms() / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  kern<<<dim3(gx, n_tiles), V3_THREADS, SL::TOTAL, st>>>(p, ta, tw);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { isf_set_error("spconv v3 launch: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
  return ISF_OK;
}

// shapes the v3 kernel takes: a neighbour table, 64 | Cin (a gather4 box is 64 channels of one kernel offset), no LayerNorm
static inline bool spconv3_supported(const GemmParams& p, int bn) {
  return p.nbr != nullptr && p.cin % BK == 0 && p.kpad == p.taps * p.cin && p.ln_gamma == nullptr && bn == p.n_total &&
         (p.nbr_stride % 4) == 0 && (((uintptr_t)p.A) & 15) == 0;
}

static inline int dispatch_spconv3(const GemmParams& p, int bn, int m_upper, cudaStream_t st) {
  switch (bn) {
    case 32: return launch_spconv3<32, 2, 4>(p, m_upper, st);       // 4 x 36 KB
    case 64: return launch_spconv3<64, 2, 4>(p, m_upper, st);       // 4 x 40 KB
    case 128: return launch_spconv3<128, 2, 3>(p, m_upper, st);     // 3 x 48 KB, 2 x 2 x 128 TMEM columns
    case 256: return launch_spconv3<256, 1, 4>(p, m_upper, st);     // 4 x 48 KB, 2 x 256 TMEM columns
    default: isf_set_error("spconv v3: unsupported column tile %d", bn); return ISF_ERR_UNSUPPORTED;
  }
}

}  // namespace tc
