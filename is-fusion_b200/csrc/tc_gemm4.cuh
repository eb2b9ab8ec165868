// tcgen05 implicit GEMM, schedule "v4": the producers of tc::gemm_kernel (tc_gemm.cuh) with the epilogue on its own warps and a
// DOUBLE-BUFFERED TMEM accumulator.  Same contract, same arithmetic order, same bits as tc::gemm_kernel (no LayerNorm).
//
// Why (round-2 source-level ncu of the stage-3 launch, profiles/ncu_gemm_r02.md): ~21 % of the warp-stall samples sat in the
// epilogue (its global loads), ~11 % at the __syncthreads that separates the epilogue of tile t from the main loop of tile
// t+1, and the tensor pipe idles while the eight producer warps — which are also the epilogue warps — drain a tile.  Here
//   warps 0-7    producers: tile prologue (neighbour tile -> shared memory, active-chunk list) and the A gathers (16-byte
//                cp.async into the 128B-swizzled UMMA layout, hardware-signalled full barrier), weight chunk by TMA;
//                they start tile t+1 as soon as their last chunk of tile t is issued;
//   warp  8      MMA issuer: tcgen05.mma into accumulator (t & 1), tcgen05.commit frees smem stages / publishes the accumulator;
//   warps 9-12   epilogue: accumulator (t & 1) -> BN fold, residual, activation -> global, row maps and the first residual
//                block fetched before the accumulator is complete.
// The CTA-wide barrier per tile is gone; the only CTA-wide synchronisation is at start-up and tear-down.  Two CTAs per SM for
// column tiles <= 128 (2 x 2 x 128 TMEM columns), one for 256.  A failed TMA-gather variant of this skeleton ("v3",
// tile::gather4 for the A rows) is documented in DESIGN.md 3.1: the TMA serves ~1 gather4 per 40-85 cycles per SM.
#pragma once
#include "tc_gemm.cuh"

namespace tc {

constexpr int V4_MMA_WARP = 8;
constexpr int V4_EPI_WARP0 = 9;
constexpr int V4_THREADS = 13 * 32;

template <int BN, int STAGES, int MT = 1>
struct Smem4 {
  static constexpr int TILE_M = MT * BM;
  static constexpr int A_BYTES = TILE_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NBR_OFF = STAGES * STAGE_BYTES;                  // [MAX_TAPS][128] byte offsets of the neighbour rows
  static constexpr int LIST_OFF = NBR_OFF + MAX_TAPS * TILE_M * 4;      // active chunk list (uint16)
  static constexpr int FLAG_OFF = LIST_OFF + MAX_CHUNKS * 2;            // 32 tap flags + count, + n_act slots of the two buffers
  static constexpr int VEC_OFF = FLAG_OFF + 64 * 4;                     // scale[BN] | shift[BN] fp32
  static constexpr int BAR_OFF = VEC_OFF + 2 * BN * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 8;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};

// MT = 128-row sub-tiles per CTA tile: with MT = 2 one weight chunk feeds 256 output rows (half the weight bytes per FLOP from L2)
// WPC ("warp per chunk", round 2 calls 14-16; opt-in, isf_set_option("gemm_wpc", 1|2)): a K chunk is gathered by one or two
// producer warps that own its smem stage instead of by all eight.  Hypothesis: the 257 mbarrier arrivals + 256 try_wait pollers
// per chunk pace the MIO queue.  Measured: bit-identical and SLOWER (21 launches 2.83 -> 3.55 ms with two warps per chunk, 4.80 ms
// with one): a single warp cannot keep 32 dependent LDS -> LDGSTS pairs per chunk in flight; with every load removed the floor
// only moved from 2.18 to 1.98 ms, i.e. the arrivals were never the cost.  Kept as the second producer mapping of the kernel.
template <int BN, int STAGES, int CTAS, int MT, bool WPC>
__global__ void __launch_bounds__(V4_THREADS, CTAS) gemm4_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_w) {
  using SL = Smem4<BN, STAGES, MT>;
  constexpr int TILE_M = SL::TILE_M;
  constexpr int TSPLIT = PRODUCERS / TILE_M;          // thread groups that split the kernel offsets in the tile prologue
  constexpr int RPT = TILE_M * 8 / PRODUCERS;         // A rows per producer thread and chunk
  constexpr int ACC_COLS = 2 * MT * BN;
  static_assert(ACC_COLS * CTAS <= 512, "TMEM holds 512 columns per SM");
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  constexpr int NVR = (MAX_TAPS + TSPLIT - 1) / TSPLIT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = (uint64_t*)(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* list_full = acc_empty + 2;         // [2]  n_act slot of the buffer is valid
  uint64_t* list_empty = list_full + 2;        // [2]  the MMA warp has read it
  uint32_t* tmem_slot = (uint32_t*)(list_empty + 2);
  int* nbr_s = (int*)(smem + SL::NBR_OFF);
  uint16_t* chunk_list = (uint16_t*)(smem + SL::LIST_OFF);
  int* tap_flag = (int*)(smem + SL::FLAG_OFF);             // [0,32) flags, [32] count (producers), [40 + buf] n_act slot (MMA warp)
  float* vec = (float*)(smem + SL::VEC_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_total = p.m_dev ? *p.m_dev : p.m;
  const int num_tiles = (m_total + TILE_M - 1) / TILE_M;
  const int n0 = blockIdx.y * BN;
  const int num_chunks = p.kpad / BK;
  const bool gather = p.nbr != nullptr;
  // tile walk: CTA b takes positions b, b + grid, ... of the tile sequence; rev = from the last tile down (under the neighbour-
  // pattern order the densest tiles sit at the end: walking them first leaves the cheap ones for the partial last round)
  const bool rev = p.rev != 0;
#define V4_TILE(idx) (rev ? num_tiles - 1 - (idx) : (idx))
  // Dynamic tile scheduling (p.tile_ctr, isf_spconv_tile_counter): the CTAs draw positions of the tile sequence from a global counter
  // instead of taking every gridDim.x-th one, and walk it from the LAST tile down (under the neighbour-pattern order the densest
  // tiles sit at the end: the cheap ones then fill the tail).  Static round-robin leaves the longest CTA 1.42x / 1.16x / 1.19x
  // above the mean on stages 2 / 3 / 4 of the bench frame (tile cost = 1 ... 27 kernel offsets), drawing 1.11x / 1.05x / 1.13x
  // (tests/analysis model).  Producer thread 0 draws two tiles ahead (the atomic's latency hides behind a tile's
  // gather loop) into a ring of 8 positions in shared memory that the three roles read; position >= num_tiles = stop.
  const bool dyn = p.tile_ctr != nullptr && gridDim.y == 1;
  volatile int* ids = (volatile int*)(tap_flag + 48);        // [8] drawn positions, indexed by (iteration & 7)
#define V4_POS(it) (dyn ? ids[(it) & 7] : (int)blockIdx.x + (it) * (int)gridDim.x)
#define V4_TILE_OF(pos) (dyn ? num_tiles - 1 - (pos) : V4_TILE(pos))

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], (WPC ? 32 * p.wpc : PRODUCERS) + 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
      mbar_init(&list_full[b], 1);
      mbar_init(&list_empty[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == V4_MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (dyn && tid == 32) {
    const int base = atomicAdd(p.tile_ctr, 2);
    ids[0] = base;
    ids[1] = base + 1;
  }
  for (int c = tid; c < BN; c += V4_THREADS) {
    vec[c] = p.scale ? p.scale[n0 + c] : 1.f;
    vec[BN + c] = p.shift ? p.shift[n0 + c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PWARPS) {
    // ================================================================ producers
    const int r = tid % TILE_M;
    const int thalf = tid / TILE_M;
    const int pc = tid & 7;
    const int rbase = tid >> 3;
    const uint32_t a_thr_off = (uint32_t)(rbase * 128 + ((pc ^ (rbase & 7)) << 4));
    const char* a_bytes = (const char*)p.A;
    uint32_t ca = 0;
    // zero-skip: one warp-level LDGSTS covers 4 rows x 128 B of a stage; when none of the 4 rows has a neighbour at the current
    // kernel offset AND the previous use of the stage left zeros in those rows, the copy is not issued (the 64 / 128-column
    // stages fill 40-60 % of their slots, and the LSU queue is what throttles their producers).  Bit (s * RPT + j) of zst.
    static_assert(STAGES * RPT <= 32, "zero-skip state");
    unsigned zst = 0;
    const bool zskip = p.zskip != 0;
    // neighbour ids of the tile AFTER the one being gathered live in registers: their global-load latency hides behind the
    // gather loop instead of sitting between two tiles (where only STAGES chunks of work are queued for the tensor core)
    int nv[NVR];
    if (gather) {
      const int pos0 = V4_POS(0);
      const long long prow = (long long)V4_TILE_OF(pos0) * TILE_M + r;
      const bool pok = pos0 < num_tiles && prow < m_total;
#pragma unroll
      for (int u = 0; u < NVR; ++u) {
        const int t = thalf + TSPLIT * u;
        nv[u] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
      }
    }
    int it = 0;
    for (;; ++it) {
      const int tidx = V4_POS(it);
      if (tidx >= num_tiles) break;
      const int tile = V4_TILE_OF(tidx);
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      const int row0 = tile * TILE_M;
      // every producer warp has issued its last chunk of the previous tile: the neighbour tile / chunk list can be replaced
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (gather) {
        if (tid < 33) tap_flag[tid] = 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int u = 0; u < NVR; ++u) {
          const int t = thalf + TSPLIT * u;
          if (t < p.taps) {
            nbr_s[t * TILE_M + r] = nv[u] >= 0 ? (int)((unsigned)nv[u] * (unsigned)(p.cin * 2)) : -1;
            if (nv[u] >= 0) tap_flag[t] = 1;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (warp == 0) {
        int count = 0;
        for (int base = 0; base < num_chunks; base += 32) {
          const int q = base + lane;
          bool act = q < num_chunks;
          if (act && gather) {
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
        if (ubuf > 0) mbar_wait_b(&list_empty[buf], (ubuf - 1) & 1);       // the MMA warp has read this slot's previous value
        if (lane == 0) { tap_flag[32] = count; tap_flag[40 + buf] = count; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&list_full[buf]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n_act = tap_flag[32];
      if (gather) {      // prefetch the next tile's neighbour ids (consumed after this tile's gather loop)
        const int nidx = V4_POS(it + 1);
        const long long prow = (long long)V4_TILE_OF(nidx) * TILE_M + r;
        const bool pok = nidx < num_tiles && prow < m_total;
#pragma unroll
        for (int u = 0; u < NVR; ++u) {
          const int t = thalf + TSPLIT * u;
          nv[u] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
        }
      }
      int drawn = 0;
      if (dyn && tid == 0) drawn = atomicAdd(p.tile_ctr, 1);       // position of iteration it + 2 (consumed after the gather loop)
      if constexpr (WPC) {
        // a chunk is gathered by G = p.wpc (1 or 2) producer warps that OWN its smem stage: warps [G s, G s + G) serve stage s and
        // nothing else, so each of them sees every phase of that stage's empty barrier in order (a warp that skipped phases would
        // alias the parity test: call 14's fault); warps >= G * STAGES only take part in the tile prologue.
        // lane = (row & 3, 16-byte piece): 4 rows x 128 B per instruction, TILE_M / (4 G) instructions per lane and chunk
        const int G = p.wpc;
        const int s = warp / G, part = warp - s * G;
        if (s < STAGES) {
          const int lrow = lane >> 3, lpc = lane & 7;
          const uint32_t sw0 = (uint32_t)(lrow * 128 + ((lpc ^ lrow) << 4));           // rows 8m + lrow
          const uint32_t sw1 = (uint32_t)(lrow * 128 + ((lpc ^ (lrow | 4)) << 4));     // rows 8m + 4 + lrow
          const uint32_t sa = smem_u32(smem + s * SL::STAGE_BYTES);
          const int jn = TILE_M / 4 / G, jb = part * jn;
          for (int i = (s + STAGES - (int)(ca % STAGES)) % STAGES; i < n_act; i += STAGES) {
            const uint32_t use = (ca + (uint32_t)i) / STAGES;
            const int q = chunk_list[i];
            if (use > 0) mbar_wait_b(&empty_bar[s], (use - 1) & 1);
            if (part == 0 && lane == 0) {
              if ((p.debug & 1) && i > 0) mbar_arrive(&full_bar[s]);      // timing experiment: weight chunk only once per tile
              else {
                mbar_expect_tx(&full_bar[s], BN * 128);
                tma_load_2d(sa + SL::A_BYTES, &tmap_w, q * BK, n0, &full_bar[s]);
              }
            }
            const int k = q * BK + lpc * 8;
            int tap, ch;
            if (p.taps == 1) { tap = k >= p.cin ? 1 : 0; ch = k; } else { tap = k >> p.cin_shift; ch = k & (p.cin - 1); }
            const bool tap_ok = tap < p.taps;
            const char* src_ch = a_bytes + ch * 2;
            const int* nrow = nbr_s + (tap_ok ? tap : 0) * TILE_M + lrow;
            if (!(p.debug & 4)) {
#pragma unroll 8
              for (int j = jb; j < jb + jn; ++j) {
                unsigned off = 0xFFFFFFFFu;
                if (tap_ok) {
                  if (gather) off = (unsigned)nrow[4 * j];
                  else { const int gr = row0 + 4 * j + lrow; if (gr < m_total) off = (unsigned)gr * (unsigned)(p.cin * 2); }
                }
                const bool ok = off != 0xFFFFFFFFu;
                cp_async16(sa + (uint32_t)(j * 512) + ((j & 1) ? sw1 : sw0), ok ? src_ch + off : a_bytes, ok ? 16u : 0u);
              }
            }
            cp_async_mbar_arrive(&full_bar[s]);
          }
        }
        ca += (uint32_t)n_act;
      } else {
        unsigned offs[RPT];
        int offs_tap = -1;
        for (int i = 0; i < n_act; ++i, ++ca) {
          const int q = chunk_list[i];
          const int s = (int)(ca % STAGES);
          const uint32_t use = ca / STAGES;
          if (use > 0) mbar_wait_b(&empty_bar[s], (use - 1) & 1);
          const uint32_t sa = smem_u32(smem + s * SL::STAGE_BYTES);
          if (tid == 0) {
            if ((p.debug & 1) && i > 0) mbar_arrive(&full_bar[s]);      // timing experiment: weight chunk only once per tile
            else {
              mbar_expect_tx(&full_bar[s], BN * 128);
              tma_load_2d(sa + SL::A_BYTES, &tmap_w, q * BK, n0, &full_bar[s]);
            }
          }
          const int k = q * BK + pc * 8;
          int tap, ch;
          if (p.taps == 1) { tap = k >= p.cin ? 1 : 0; ch = k; } else { tap = k >> p.cin_shift; ch = k & (p.cin - 1); }
          const bool tap_ok = tap < p.taps;
          const char* src_ch = a_bytes + ch * 2;
          if (tap != offs_tap) {
            offs_tap = tap;
            const int* nrow = nbr_s + tap * TILE_M + rbase;
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              unsigned off = 0xFFFFFFFFu;
              if (tap_ok) {
                if (gather) off = (unsigned)nrow[32 * j];
                else { const int gr = row0 + rbase + 32 * j; if (gr < m_total) off = (unsigned)gr * (unsigned)(p.cin * 2); }
              }
              offs[j] = off;
            }
          }
          unsigned skip = 0;
          if (zskip) {
            unsigned miss = 0;
#pragma unroll
            for (int j = 0; j < RPT; ++j)
              if (__all_sync(0xffffffffu, offs[j] == 0xFFFFFFFFu)) miss |= 1u << j;
            skip = miss & (zst >> (s * RPT));
            zst = (zst & ~(((1u << RPT) - 1u) << (s * RPT))) | (miss << (s * RPT));
          }
          if (!(p.debug & 4))
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            if (skip & (1u << j)) continue;
            const bool ok = offs[j] != 0xFFFFFFFFu;
            cp_async16(sa + a_thr_off + j * 32 * 128, ok ? src_ch + offs[j] : a_bytes, ok ? 16u : 0u);
          }
          cp_async_mbar_arrive(&full_bar[s]);
        }
      }
      if (dyn && tid == 0) ids[(it + 2) & 7] = drawn;      // read by the roles after the next tile's first producer barrier at the earliest
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (dyn && tid == 0) {      // every CTA has stopped drawing when it gets here: the last one re-arms the counter for the next launch
      __threadfence();
      if (atomicAdd(p.tile_ctr + 1, 1) == (int)gridDim.x - 1) { p.tile_ctr[0] = 0; p.tile_ctr[1] = 0; }
    }
  } else if (warp == V4_MMA_WARP) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc_f16(BM, BN) | (p.bf16 ? IDESC_AB_BF16 : 0u);
    uint32_t ca = 0;
    int it = 0;
    for (;; ++it) {
      if (V4_POS(it) >= num_tiles) break;
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      mbar_wait_b(&list_full[buf], ubuf & 1);
      const int n_act = ((const volatile int*)tap_flag)[40 + buf];
      __syncwarp();
      if (lane == 0) mbar_arrive(&list_empty[buf]);
      if (ubuf > 0) mbar_wait_b(&acc_empty[buf], (ubuf - 1) & 1);      // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * MT * BN);
      for (int q = 0; q < n_act; ++q, ++ca) {
        const int s = (int)(ca % STAGES);
        mbar_wait_b(&full_bar[s], (ca / STAGES) & 1);
        fence_proxy_async();     // cp.async (generic proxy) writes of the producers -> visible to the tensor core
        tc_fence_after();
        {      // whole warp, converged: one election per chunk (umma_chunk, tc_gemm.cuh); sub-tiles alternate per K step
          const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES);
          umma_chunk<1, MT>(d_tmem, d_tmem + (uint32_t)BN, make_desc_sw128(a_addr), make_desc_sw128(a_addr + (uint32_t)(BM * 128)),
                            make_desc_sw128(a_addr + SL::A_BYTES), idesc, q ? 1u : 0u, &empty_bar[s]);
          if (q == n_act - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================ epilogue (warps 9-12; TMEM lane quarter = warp % 4)
    const int r = (warp & 3) * 32 + lane;
    int it = 0;
    for (;; ++it) {
      const int tidx = V4_POS(it);
      if (tidx >= num_tiles) break;
      const int tile = V4_TILE_OF(tidx);
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int grow = tile * TILE_M + mt * BM + r;
      const bool row_ok = grow < m_total;
      long long orow = grow;
      if (p.out_rows && row_ok) orow = p.out_rows[grow];
      long long rrow = row_ok ? (long long)grow : -1;
      if (p.res_rows && row_ok) rrow = p.res_rows[grow];
      const bool store_ok = row_ok && orow >= 0;
      const bool has_res = p.residual != nullptr && rrow >= 0;
      const uint4* rp = has_res ? (const uint4*)(p.residual + rrow * p.ldr + n0) : nullptr;
      uint4 rnext[2];
      if (has_res) { rnext[0] = rp[0]; rnext[1] = rp[1]; }
      if (mt == 0) {
        mbar_wait_b(&acc_full[buf], ubuf & 1);
        tc_fence_after();
      }
      const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((buf * MT + mt) * BN);
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 16) {
        uint4 rcur[2];
        if (has_res) {
          rcur[0] = rnext[0]; rcur[1] = rnext[1];
          if (cb + 16 < BN) { rnext[0] = rp[(cb + 16) / 8]; rnext[1] = rp[(cb + 16) / 8 + 1]; }
        }
        uint32_t v[16];
        tmem_ld16(t_lane + cb, v);
        tmem_ld_wait();
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float y = __uint_as_float(v[j]);
          if (p.scale) y = __fmul_rn(y, vec[cb + j]);
          if (p.shift) y = __fadd_rn(y, vec[BN + cb + j]);
          x[j] = y;
        }
        if (has_res) {
#pragma unroll
          for (int j4 = 0; j4 < 2; ++j4) {
            const uint32_t* h = (const uint32_t*)&rcur[j4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = isf_unpack2(h[e], p.bf16 != 0);
              x[j4 * 8 + e * 2] += f.x; x[j4 * 8 + e * 2 + 1] += f.y;
            }
          }
        }
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = gelu_erf(x[j]);
        }
        if (store_ok) {
          if (p.out) {
            uint4* op = (uint4*)(p.out + orow * p.ldo + n0 + cb);
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              uint4 u;
              uint32_t* h = (uint32_t*)&u;
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = isf_pack2(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1], p.bf16 != 0);
              op[j4] = u;
            }
          }
          if (p.out_f32) {
            float4* op = (float4*)(p.out_f32 + orow * p.ldo + n0 + cb);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) op[j4] = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
          }
        }
      }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == V4_MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
#undef V4_TILE
#undef V4_POS
#undef V4_TILE_OF
}

template <int BN, int STAGES, int CTAS, int MT, bool WPC>
static int launch_gemm4_impl(const GemmParams& p, int m_upper, cudaStream_t st) {
  using SL = Smem4<BN, STAGES, MT>;
  static_assert((SL::TOTAL + 1024) * CTAS <= 227 * 1024, "v4 GEMM: shared memory budget exceeded");
  auto kern = gemm4_kernel<BN, STAGES, CTAS, MT, WPC>;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { isf_set_error("gemm v4: bad device"); return ISF_ERR_CUDA; }
  static bool configured[64] = {};
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
    if (e != cudaSuccess) { isf_set_error("cudaFuncSetAttribute (v4): %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
    configured[dev] = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int rc = make_weight_tmap(&tm, p.W, p.n_total, p.kpad, BN);
  if (rc) return rc;
  int tiles = (m_upper + SL::TILE_M - 1) / SL::TILE_M;
  if (tiles < 1) tiles = 1;
  const int n_tiles = p.n_total / BN;
  int gx = isf_num_sms() * CTAS / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  // the fewest CTAs that finish in the same number of tile rounds: 780 tiles on 148 CTAs are 6 rounds, 130 CTAs do them in 6 as
  // well - the SMs left over run the small kernels of the other frames in flight (isf_set_option("gemm_trim", 0) = all CTAs)
  if (::g_isf_gemm_trim && p.nbr != nullptr) {
    const int rounds = (tiles + gx - 1) / gx;
    gx = (tiles + rounds - 1) / rounds;
  }
  kern<<<dim3(gx, n_tiles), V4_THREADS, SL::TOTAL, st>>>(p, tm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { isf_set_error("gemm v4 launch: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
  return ISF_OK;
}

template <int BN, int STAGES, int CTAS, int MT = 1>
static int launch_gemm4(const GemmParams& p, int m_upper, cudaStream_t st) {
  if (::g_isf_gemm_wpc) {
    GemmParams q = p;
    q.wpc = (::g_isf_gemm_wpc == 2 && STAGES * 2 <= PWARPS) ? 2 : 1;
    return launch_gemm4_impl<BN, STAGES, CTAS, MT, true>(q, m_upper, st);
  }
  return launch_gemm4_impl<BN, STAGES, CTAS, MT, false>(p, m_upper, st);
}

static inline bool gemm4_supported(const GemmParams& p, int bn) { return p.ln_gamma == nullptr && (bn >= 64 || (bn == 32 && ::g_isf_gemm_v4_32)); }

static inline int dispatch_gemm4(const GemmParams& p, int bn, int m_upper, cudaStream_t st) {
  // 256-row tiles (one weight chunk per 256 rows): measured on v4 (round 2, call 8) -8 % on the 64-column layers, +7 % on the
  // 128-column ones -> default (gemm_mt = 3) only for 64 columns; gemm_mt = 2 forces them for 128 as well, 1 disables them
  if (p.nbr != nullptr && m_upper >= 2 * 148 * 256 && ::g_isf_gemm_mt >= 2) {
    if (bn == 64) return launch_gemm4<64, 2, 2, 2>(p, m_upper, st);                              // 2 x 40 KB + 27 KB, 2 CTAs / SM
    if (bn == 128 && ::g_isf_gemm_mt == 2) return launch_gemm4<128, 4, 1, 2>(p, m_upper, st);    // 4 x 48 KB + 27 KB, 512 TMEM columns
  }
  switch (bn) {
    case 32:                                                       // stage 1: 3-5 chunks per tile, latency-bound
      if (::g_isf_gemm_v4_32 == 2) return launch_gemm4<32, 2, 3>(p, m_upper, st);      // 2 x 20 KB + 14 KB, 3 CTAs / SM (48 registers)
      return launch_gemm4<32, 4, 2>(p, m_upper, st);                                   // 4 x 20 KB + 14 KB, 2 CTAs / SM
    case 64: return launch_gemm4<64, 4, 2>(p, m_upper, st);        // 4 x 24 KB, 2 x 128 TMEM columns per CTA
    case 128: return launch_gemm4<128, 3, 2>(p, m_upper, st);      // 3 x 32 KB
    case 256: return launch_gemm4<256, 4, 1>(p, m_upper, st);      // 4 x 48 KB, all 512 TMEM columns
    default: isf_set_error("gemm v4: unsupported column tile %d", bn); return ISF_ERR_UNSUPPORTED;
  }
}

}  // namespace tc
