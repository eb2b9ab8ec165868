// tcgen05 implicit GEMM, warp-specialised variant ("v2") — OPT-IN (isf_set_option("gemm_v2", 1|2)), not the default:
// written at the end of round 1 from the pacing model in DESIGN.md §3.1, NOT yet run on a GPU.  Same contract and
// epilogue as tc::gemm_kernel (tc_gemm.cuh); what changes is the schedule inside the CTA:
//
//   * one CTA per SM (persistent), 13 warps with fixed roles:
//       warps 0-3   epilogue   (TMEM lane = output row = thread; overlaps the main loop of the NEXT tile)
//       warp  4     MMA issuer (allocates TMEM; one elected lane issues tcgen05.mma / tcgen05.commit)
//       warps 5-12  producers  (neighbour tile + chunk list, A rows by 16-byte cp.async into the 128B-swizzled layout,
//                               weight chunks by TMA)
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns): the MMA warp starts tile t+1 while the epilogue
//     warps drain tile t -> the gather never stops for a prologue / epilogue bubble (v1 hides those only through
//     2-3 co-resident CTAs, which costs smem = gathered bytes in flight);
//   * separate rings for A (SA stages x 16 KB) and the weights (SB stages x BN*128 B): the weight chunks come from L2
//     with a short latency and need fewer stages, so the A ring — the bytes in flight that cover the gather latency —
//     gets the smem (e.g. BN=256: 6 x 16 KB of A in flight per SM instead of 4 x 16 KB);
//   * the neighbour ids of tile t+1 are loaded into registers before the gather loop of tile t (their global latency is
//     hidden), the neighbour tile / chunk list in smem are double-buffered;
//   * 32-byte epilogue accesses (v2_flags bits 1, 2; isf_set_option("gemm_v2_io256", 0|1)): residual rows and fp16 output
//     rows move as LDG/STG.E.ENL2.256 — a full sector per lane — instead of 16-byte halves;
//   * the per-column epilogue vectors (BN scale / shift, LayerNorm gamma / beta) are staged in shared memory once per CTA
//     instead of being fetched by uniform global loads per column and tile;
//   * zero-skip (v2_flags bit 0): one warp-level LDGSTS covers 4 rows x 128 B of a stage; when none of the 4 rows has
//     a neighbour AND the previous use of that stage left zeros in those rows, the copy is skipped (the sparse 32/64-
//     channel stages fill 15-35 % of the slots, and the LSU issue rate of these copies is what bounds them).
//
// Every mbarrier wait is bounded (~2 s) and traps instead of hanging the GPU while the kernel is being brought up.
#pragma once
#include "tc_gemm.cuh"

namespace tc {

constexpr int V2_EPI_WARPS = 4;
constexpr int V2_MMA_WARP = 4;
constexpr int V2_PROD_WARP0 = 5;
constexpr int V2_PRODUCERS = 256;
constexpr int V2_THREADS = (V2_PROD_WARP0 + V2_PRODUCERS / 32) * 32;   // 416

template <int BN, int SA, int SB>
struct Smem2 {
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = SA * A_BYTES;
  static constexpr int NBR_OFF = B_OFF + SB * B_BYTES;                  // 2 x [MAX_TAPS][128] byte offsets
  static constexpr int LIST_OFF = NBR_OFF + 2 * MAX_TAPS * BM * 4;      // 2 x chunk list (uint16)
  static constexpr int FLAG_OFF = LIST_OFF + 2 * MAX_CHUNKS * 2;        // 2 x (32 tap flags + count + pad)
  static constexpr int VEC_OFF = FLAG_OFF + 2 * 64 * 4;                 // [4][BN] fp32: scale, shift, ln_gamma, ln_beta of the CTA's columns
  static constexpr int BAR_OFF = VEC_OFF + 4 * BN * 4;
  static constexpr int NUM_BARS = 2 * SA + 2 * SB + 8;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;      // + tmem slot + alignment slack
  static_assert(TOTAL <= 227 * 1024, "v2 GEMM: shared memory budget exceeded");
};

__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol error must not hang the box
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i)
      if (mbar_try(bar, parity)) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

// 32 accumulator columns of one row with the per-column vectors read from SHARED memory (staged once per CTA): the
// v1 epilogue fetches scale / shift / gamma / beta with one uniform LDG per column and tile (~100-260 global loads per
// thread and tile, ncu: 50-56 us for the LayerNorm GEMMs against 12-20 us without).  Arithmetic and order as epilogue_block.
template <bool LD256>
__device__ __forceinline__ void epilogue_block2(const GemmParams& p, uint32_t taddr, const float* __restrict__ vs,
                                                const float* __restrict__ vh, int n_base, long long rrow, float (&x)[32]) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  tmem_ld_wait();
  const bool has_s = p.scale != nullptr, has_h = p.shift != nullptr;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 s4 = *(const float4*)(vs + j4 * 4), h4 = *(const float4*)(vh + j4 * 4);
    const float sc[4] = {s4.x, s4.y, s4.z, s4.w}, sh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float y = __uint_as_float(v[j4 * 4 + e]);
      if (has_s) y *= sc[e];
      if (has_h) y += sh[e];
      x[j4 * 4 + e] = y;
    }
  }
  if (p.residual && rrow >= 0) {
    if constexpr (LD256) {
#pragma unroll
      for (int j8 = 0; j8 < 2; ++j8) {
        uint32_t w[8];
        ld_global_256(p.residual + rrow * p.ldr + n_base + j8 * 16, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 f = __half22float2(*(const __half2*)&w[e]);
          x[j8 * 16 + e * 2] += f.x; x[j8 * 16 + e * 2 + 1] += f.y;
        }
      }
    } else {
      const uint4* rp = (const uint4*)(p.residual + rrow * p.ldr + n_base);
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
  }
}

// one output row (TMEM lane) of a finished tile: scale/shift (+residual) (+LayerNorm) (+activation) -> global.
// LD256: 32-byte residual loads; (p.v2_flags & 4): 32-byte fp16 stores — one full sector per lane and instruction
// instead of half a sector (the un-coalesced row-per-thread epilogue is 15-25 % of the L1TEX load).
// vec = the CTA's [4][BN] per-column vectors in shared memory.
template <int BN, bool LD256>
__device__ __forceinline__ void epilogue_tile2(const GemmParams& p, uint32_t t_lane, int n0, long long rrow, long long orow,
                                               bool store_ok, const float* __restrict__ vec) {
  float mean = 0.f, rstd = 1.f;
  if (p.ln_gamma) {
    float s1 = 0.f, s2 = 0.f;
    for (int cb = 0; cb < BN; cb += 32) {
      float x[32];
      epilogue_block2<LD256>(p, t_lane + cb, vec + cb, vec + BN + cb, n0 + cb, rrow, x);
#pragma unroll
      for (int j = 0; j < 32; ++j) { s1 += x[j]; s2 += x[j] * x[j]; }
    }
    mean = s1 / BN;
    const float var = s2 / BN - mean * mean;
    rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
  }
  const bool st256 = (p.v2_flags & 4) != 0;
  for (int cb = 0; cb < BN; cb += 32) {
    float x[32];
    epilogue_block2<LD256>(p, t_lane + cb, vec + cb, vec + BN + cb, n0 + cb, rrow, x);
    if (p.ln_gamma) {
      const float* ga = vec + 2 * BN + cb;
      const float* be = vec + 3 * BN + cb;
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = (x[j] - mean) * rstd * ga[j] + be[j];
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
        __half* op = p.out + orow * p.ldo + n0 + cb;
        if (st256) {
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const __half2 h = __floats2half2_rn(x[j8 * 16 + e * 2], x[j8 * 16 + e * 2 + 1]);
              w[e] = *(const uint32_t*)&h;
            }
            st_global_256(op + j8 * 16, w);
          }
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint4 u;
            __half2* h = (__half2*)&u;
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1]);
            ((uint4*)op)[j4] = u;
          }
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

template <int BN, int SA, int SB>
__global__ void __launch_bounds__(V2_THREADS, 1) gemm2_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_w) {
  using SL = Smem2<BN, SA, SB>;
  constexpr int ACC_COLS = 2 * BN;                                   // double-buffered accumulator
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  static_assert(ACC_COLS <= 512, "TMEM holds 512 columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = (uint64_t*)(smem + SL::BAR_OFF);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* acc_full = b_empty + SB;      // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint64_t* list_full = acc_empty + 2;    // [2]
  uint64_t* list_empty = list_full + 2;   // [2]
  uint32_t* tmem_slot = (uint32_t*)(list_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_total = p.m_dev ? *p.m_dev : p.m;
  const int num_tiles = (m_total + BM - 1) / BM;
  const int n0 = blockIdx.y * BN;
  const int num_chunks = p.kpad / BK;
  const bool gather = p.nbr != nullptr;

  if (tid == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], V2_PRODUCERS); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], V2_EPI_WARPS * 32);
      mbar_init(&list_full[b], 1);
      mbar_init(&list_empty[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == V2_MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  {   // per-column epilogue vectors of this CTA's columns -> shared memory (read there by every tile's epilogue)
    float* vec = (float*)(smem + SL::VEC_OFF);
    for (int c = tid; c < BN; c += V2_THREADS) {
      vec[c] = p.scale ? p.scale[n0 + c] : 1.f;
      vec[BN + c] = p.shift ? p.shift[n0 + c] : 0.f;
      vec[2 * BN + c] = p.ln_gamma ? p.ln_gamma[n0 + c] : 1.f;
      vec[3 * BN + c] = p.ln_beta ? p.ln_beta[n0 + c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= V2_PROD_WARP0) {
    // ================================================================ producers
    const int ptid = tid - V2_PROD_WARP0 * 32;       // 0..255
    const int r = ptid & (BM - 1);                   // row of the tile for the neighbour-tile staging
    const int thalf = ptid >> 7;                     // the two thread halves split the kernel offsets
    const int pc = ptid & 7;                         // 16-byte piece of the 128 B chunk row
    const int rbase = ptid >> 3;                     // rows rbase + 32 j
    const uint32_t a_thr_off = (uint32_t)(rbase * 128 + ((pc ^ (rbase & 7)) << 4));
    const char* a_bytes = (const char*)p.A;
    constexpr int NV = (MAX_TAPS + 1) / 2;
    int nv[NV];
    // neighbour ids of the first tile
    {
      const long long prow = (long long)blockIdx.x * BM + r;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int t = thalf + 2 * u;
        nv[u] = (gather && prow < m_total && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
      }
    }
    uint32_t ca = 0;                                  // running chunk counter (A stage = ca % SA, B stage = ca % SB)
    unsigned long long zstate = 0;                    // zero-skip state, 4 bits per A stage (SA <= 16)
    const bool zskip = (p.v2_flags & 1) != 0;
    static_assert(SA <= 16, "zero-skip state holds 16 stages");
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      const int row0 = tile * BM;
      int* nbr_s = (int*)(smem + SL::NBR_OFF) + buf * (MAX_TAPS * BM);
      uint16_t* chunk_list = (uint16_t*)(smem + SL::LIST_OFF) + buf * MAX_CHUNKS;
      int* tap_flag = (int*)(smem + SL::FLAG_OFF) + buf * 64;      // [32] flags, [32] = active chunk count
      int n_act = num_chunks;
      if (gather) {
        if (ubuf > 0) mbar_wait_b(&list_empty[buf], (ubuf - 1) & 1);   // the MMA warp has read this buffer's count
        if (ptid < 33) tap_flag[ptid] = 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          const int t = thalf + 2 * u;
          if (t < p.taps) {
            nbr_s[t * BM + r] = nv[u] >= 0 ? (int)((unsigned)nv[u] * (unsigned)(p.cin * 2)) : -1;
            if (nv[u] >= 0) tap_flag[t] = 1;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == V2_PROD_WARP0) {
          int count = 0;
          for (int base = 0; base < num_chunks; base += 32) {
            const int q = base + lane;
            bool act = q < num_chunks;
            if (act) {
              const int t0 = (q * BK) / p.cin, t1 = (q * BK + BK - 1) / p.cin;
              bool any = false;
              for (int t = t0; t <= t1 && t < p.taps; ++t) any |= tap_flag[t] != 0;
              act = any;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, act);
            if (act) chunk_list[count + __popc(mask & ((1u << lane) - 1))] = (uint16_t)q;
            count += __popc(mask);
          }
          if (count == 0) { if (lane == 0) chunk_list[0] = 0; count = 1; }   // keeps the accumulator defined
          if (lane == 0) tap_flag[32] = count;
          __syncwarp();
          if (lane == 0) mbar_arrive(&list_full[buf]);
        }
        // neighbour ids of this CTA's NEXT tile: in flight during the whole gather loop below
        {
          const long long prow = (long long)(tile + gridDim.x) * BM + r;
          const bool pok = (tile + (int)gridDim.x) < num_tiles && prow < m_total;
#pragma unroll
          for (int u = 0; u < NV; ++u) {
            const int t = thalf + 2 * u;
            nv[u] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        n_act = tap_flag[32];
      }
      unsigned offs[BM * 8 / V2_PRODUCERS];
      int offs_tap = -1;
      unsigned miss = 0;      // bit j: none of the 4 rows of this warp's j-th copy has a neighbour at the current offset
      for (int i = 0; i < n_act; ++i, ++ca) {
        const int q = gather ? (int)chunk_list[i] : i;
        const int sa = (int)(ca % SA), sb = (int)(ca % SB);
        const uint32_t ua = ca / SA, ub = ca / SB;
        if (ua > 0) mbar_wait_b(&a_empty[sa], (ua - 1) & 1);
        const uint32_t a_dst = smem_u32(smem + SL::A_OFF + sa * SL::A_BYTES);
        if (ptid == 0) {
          if (ub > 0) mbar_wait_b(&b_empty[sb], (ub - 1) & 1);
          mbar_expect_tx(&b_full[sb], BN * 128);
          tma_load_2d(smem_u32(smem + SL::B_OFF + sb * SL::B_BYTES), &tmap_w, q * BK, n0, &b_full[sb]);
        }
        const int k = q * BK + pc * 8;
        int tap, ch;
        if (p.taps == 1) { tap = k >= p.cin ? 1 : 0; ch = k; } else { tap = k >> p.cin_shift; ch = k & (p.cin - 1); }
        const char* src_ch = a_bytes + ch * 2;
        if (__any_sync(0xffffffffu, tap != offs_tap)) {        // warp-uniform: the votes below need every lane
          offs_tap = tap;
          const bool tap_ok = tap < p.taps;
          const int* nrow = nbr_s + tap * BM + rbase;
#pragma unroll
          for (int j = 0; j < BM * 8 / V2_PRODUCERS; ++j) {
            unsigned off = 0xFFFFFFFFu;
            if (tap_ok) {
              if (gather) off = (unsigned)nrow[(V2_PRODUCERS / 8) * j];
              else { const int gr = row0 + rbase + (V2_PRODUCERS / 8) * j; if (gr < m_total) off = (unsigned)gr * (unsigned)(p.cin * 2); }
            }
            offs[j] = off;
          }
          miss = 0;
#pragma unroll
          for (int j = 0; j < BM * 8 / V2_PRODUCERS; ++j)
            if (__all_sync(0xffffffffu, offs[j] == 0xFFFFFFFFu)) miss |= 1u << j;
        }
        // zstate bit (4*sa + j): the 4 rows of copy j in stage sa currently hold zeros (left there by an all-missing copy)
        const unsigned zbits = (unsigned)((zstate >> (4 * sa)) & 0xFull);
        const unsigned skip = zskip ? (miss & zbits) : 0u;
#pragma unroll
        for (int j = 0; j < BM * 8 / V2_PRODUCERS; ++j) {
          if (skip & (1u << j)) continue;
          const bool ok = offs[j] != 0xFFFFFFFFu;
          cp_async16(a_dst + a_thr_off + j * (V2_PRODUCERS / 8) * 128, ok ? src_ch + offs[j] : a_bytes, ok ? 16u : 0u);
        }
        zstate = (zstate & ~(0xFull << (4 * sa))) | ((unsigned long long)miss << (4 * sa));
        cp_async_mbar_arrive(&a_full[sa]);
      }
    }
    // every cp.async of this thread has landed before the CTA tears down
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == V2_MMA_WARP) {
    // ================================================================ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(BM, BN);
    uint32_t ca = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      int n_act = num_chunks;
      if (gather) {
        mbar_wait_b(&list_full[buf], ubuf & 1);
        n_act = ((const volatile int*)(smem + SL::FLAG_OFF))[buf * 64 + 32];
        __syncwarp();
        if (lane == 0) mbar_arrive(&list_empty[buf]);
      }
      if (ubuf > 0) mbar_wait_b(&acc_empty[buf], (ubuf - 1) & 1);     // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
      for (int q = 0; q < n_act; ++q, ++ca) {
        const int sa = (int)(ca % SA), sb = (int)(ca % SB);
        mbar_wait_b(&a_full[sa], (ca / SA) & 1);
        mbar_wait_b(&b_full[sb], (ca / SB) & 1);
        fence_proxy_async();     // cp.async (generic proxy) writes of the producers -> visible to the tensor core
        tc_fence_after();
        if (lane == 0) {
          const uint64_t a_desc = make_desc_sw128(smem_u32(smem + SL::A_OFF + sa * SL::A_BYTES));
          const uint64_t b_desc = make_desc_sw128(smem_u32(smem + SL::B_OFF + sb * SL::B_BYTES));
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            umma_f16(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (q | kk) ? 1u : 0u);
          umma_commit(&a_empty[sa]);
          umma_commit(&b_empty[sb]);
          if (q == n_act - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================ epilogue (warps 0-3)
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      const int grow = tile * BM + tid;                 // tid in [0,128): TMEM lane = 32*warp + lane = tid
      const bool row_ok = grow < m_total;
      long long orow = grow;
      if (p.out_rows && row_ok) orow = p.out_rows[grow];
      const bool store_ok = row_ok && orow >= 0;
      long long rrow = row_ok ? (long long)grow : -1;
      if (p.res_rows && row_ok) rrow = p.res_rows[grow];
      mbar_wait_b(&acc_full[buf], ubuf & 1);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN);
      const float* vec = (const float*)(smem + SL::VEC_OFF);
      if (p.v2_flags & 2) epilogue_tile2<BN, true>(p, t_lane, n0, rrow, orow, store_ok, vec);
      else epilogue_tile2<BN, false>(p, t_lane, n0, rrow, orow, store_ok, vec);
      // this thread's TMEM reads of the buffer are complete (tcgen05.wait::ld inside epilogue_block) and ordered
      // before the arrival that lets the MMA warp overwrite it
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == V2_MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BN, int SA, int SB>
static int launch_gemm2(const GemmParams& p, int m_upper, cudaStream_t st) {
  using SL = Smem2<BN, SA, SB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<BN, SA, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
    if (e != cudaSuccess) { isf_set_error("cudaFuncSetAttribute (v2): %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
    configured = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int rc = make_weight_tmap(&tm, p.W, p.n_total, p.kpad, BN);
  if (rc) return rc;
  int tiles = (m_upper + BM - 1) / BM;
  if (tiles < 1) tiles = 1;
  const int n_tiles = p.n_total / BN;
  int gx = ISF_NUM_SMS / n_tiles;               // one CTA per SM
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  gemm2_kernel<BN, SA, SB><<<dim3(gx, n_tiles), V2_THREADS, SL::TOTAL, st>>>(p, tm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { isf_set_error("gemm v2 launch: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
  return ISF_OK;
}

// ring depths: the A ring takes the shared memory the weight ring does not need
static inline int dispatch_gemm2(const GemmParams& p0, int bn, int m_upper, cudaStream_t st) {
  GemmParams p = p0;
  p.v2_flags = ::g_isf_gemm_v2_zskip ? 1 : 0;
  if (::g_isf_gemm_v2_io256) {       // 32-byte epilogue accesses where every row start is 32-byte aligned
    if (p.residual && ((uintptr_t)p.residual & 31) == 0 && p.ldr % 16 == 0) p.v2_flags |= 2;
    if (p.out && ((uintptr_t)p.out & 31) == 0 && p.ldo % 16 == 0) p.v2_flags |= 4;
  }
  switch (bn) {
    case 32: return launch_gemm2<32, 10, 3>(p, m_upper, st);      // 160 + 12 KB
    case 64: return launch_gemm2<64, 9, 3>(p, m_upper, st);       // 144 + 24 KB
    case 128: return launch_gemm2<128, 8, 3>(p, m_upper, st);     // 128 + 48 KB
    case 256: return launch_gemm2<256, 6, 3>(p, m_upper, st);     //  96 + 96 KB
    default: isf_set_error("gemm v2: unsupported column tile %d", bn); return ISF_ERR_UNSUPPORTED;
  }
}

}  // namespace tc
