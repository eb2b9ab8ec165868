// tcgen05 implicit GEMM, schedule "v5": the v4 schedule (tc_gemm4.cuh) on a CTA PAIR with tcgen05.mma.cta_group::2.
//
// Why (round 2, calls 13-16): with every global load removed (timing experiments) the v4 kernel still needs
// ~126 SM-cycles per M = 128 MMA instruction whatever its N (64 / 128), ~200 at N = 256; with the loads the cost per 128 x 64
// block is that floor PLUS the bytes the producers write: the tensor core fetches its smem operands at ~64 B / clk
// ((128 + N) x 32 B per K = 16 step) through the same port the LDGSTS / TMA writes use.  An SS-mode cta_group::1 instruction
// of shape 128 x N therefore cannot exceed N / (128 + N) of the issue rate... and the same arithmetic (8 KB of operand reads per
// 128 cycles + the TMA writes) reproduces the measured cuBLAS peak (68 % of nominal).  cta_group::2 is the hardware's answer:
// the pair computes M = 256 rows x N columns per instruction, each CTA supplying ITS 128 rows of A and HALF of B, so the operand
// bytes read per CTA and FLOP drop by (128 + N) / (128 + N / 2) and the weight bytes each SM pulls from L2 halve.
//
// Differences to v4:
//   * cluster of 2 CTAs; super-tile = 2 x TILE_M rows, CTA r gathers rows [r TILE_M, (r + 1) TILE_M) and TMA-loads weight rows
//     [r BN / 2, (r + 1) BN / 2) of the chunk into its own stage;
//   * both CTAs walk the SAME chunk list: the per-tile kernel-offset masks are exchanged through distributed shared memory
//     (st.async with complete_tx on the peer's barrier) and OR-ed;
//   * the leader's MMA thread waits for its own full barrier and for `peer_full` (one remote arrival per chunk from the peer's
//     otherwise idle MMA warp, which relays the peer's full barrier: cp.async cannot signal a remote barrier), issues
//     tcgen05.mma.cta_group::2 and commits with multicast to the empty / accumulator barriers of both CTAs;
//   * each CTA's epilogue warps drain their own 128 TMEM lanes; the peer's arrive on the leader's acc_empty remotely.
// Same arithmetic per output row (same chunks in the same order into an fp32 accumulator) -> same bits as v4.
#pragma once
#include "tc_gemm4.cuh"

namespace tc {

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Plain (cta-scope release) remote arrive, as cutlass::arch::ClusterBarrier::arrive(cta_id): the .release.cluster form costs a
// MEMBAR.ALL.GPU per arrival and the .acquire.cluster try_wait a CCTL.IVALL per success (call 17: the relay thread of the peer CTA
// then serves one chunk per ~900 cycles and paces the whole pair).  What the barriers order here is consumed by the tensor core
// (behind fence.proxy.async) or carried by st.async, not by generic loads.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// remote 4-byte store that completes 4 transaction bytes on a barrier of the destination CTA when the data has landed
__device__ __forceinline__ void st_async_u32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(v), "r"(cluster_bar)
               : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// whole warp, converged; the elected lane issues (see umma_f16 in tc_gemm.cuh)
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit2_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int BN, int STAGES, int MT = 1>
struct Smem5 {
  static constexpr int TILE_M = MT * BM;                                // rows per CTA (the pair covers 2 TILE_M)
  static constexpr int A_BYTES = TILE_M * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;                        // this CTA's half of the weight chunk
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NBR_OFF = STAGES * STAGE_BYTES;
  static constexpr int LIST_OFF = NBR_OFF + MAX_TAPS * TILE_M * 4;
  static constexpr int FLAG_OFF = LIST_OFF + MAX_CHUNKS * 2;            // 32 tap flags + count, n_act slots, peer masks
  static constexpr int VEC_OFF = FLAG_OFF + 64 * 4;
  static constexpr int BAR_OFF = VEC_OFF + 2 * BN * 4;
  static constexpr int NUM_BARS = 3 * STAGES + 10;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte alignment of the 128B swizzle");
};

// MMAW = MMA-issuing warps of the leader (1, or 2 with MT = 2: warp 8 issues sub-tile 0, warp 13 sub-tile 1).  One thread issues a
// tcgen05.mma per ~90-100 cycles (bench_mma_rate.py); a 256 x 128 x 16 pair instruction occupies the two tensor pipes for 64, so a
// single issuer leaves a third of the pipe idle on the 128-column layers.  Each warp commits its own instructions: the stage /
// accumulator barriers then expect MMAW arrivals.
template <int BN, int STAGES, int CTAS, int MT, int MMAW = 1>
__global__ void __launch_bounds__(V4_THREADS + (MMAW - 1) * 32, CTAS) gemm5_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_w) {
  static_assert(MMAW == 1 || (MMAW == 2 && MT == 2), "two MMA warps split the two sub-tiles");
  constexpr int NTHREADS = V4_THREADS + (MMAW - 1) * 32;
  constexpr int MMA_WARP1 = V4_THREADS / 32;          // the second MMA warp sits behind the epilogue warps
  using SL = Smem5<BN, STAGES, MT>;
  constexpr int TILE_M = SL::TILE_M;
  constexpr int TSPLIT = PRODUCERS / TILE_M;
  constexpr int RPT = TILE_M * 8 / PRODUCERS;
  constexpr int ACC_COLS = 2 * MT * BN;
  static_assert(ACC_COLS * CTAS <= 512, "TMEM holds 512 columns per SM");
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  constexpr int NVR = (MAX_TAPS + TSPLIT - 1) / TSPLIT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = (uint64_t*)(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* peer_full = empty_bar + STAGES;    // leader only: the peer's stage is full (one remote arrival per chunk)
  uint64_t* acc_full = peer_full + STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]  leader only: 128 local + 128 remote epilogue threads
  uint64_t* list_full = acc_empty + 2;         // [2]
  uint64_t* list_empty = list_full + 2;        // [2]
  uint64_t* xbar = list_empty + 2;             // [2]  the peer's kernel-offset mask of the tile has arrived
  uint32_t* tmem_slot = (uint32_t*)(xbar + 2);
  int* nbr_s = (int*)(smem + SL::NBR_OFF);
  uint16_t* chunk_list = (uint16_t*)(smem + SL::LIST_OFF);
  int* tap_flag = (int*)(smem + SL::FLAG_OFF);             // [0,32) flags, [32] count, [40 + buf] n_act slot, [48 + buf] peer mask
  float* vec = (float*)(smem + SL::VEC_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader (issues the MMAs)
  const int m_total = p.m_dev ? *p.m_dev : p.m;
  const int num_super = (m_total + 2 * TILE_M - 1) / (2 * TILE_M);
  const int pair0 = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int num_chunks = p.kpad / BK;
  const bool rev = p.rev != 0;
#define V5_SUPER(idx) (rev ? num_super - 1 - (idx) : (idx))

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], PRODUCERS + 1);
      mbar_init(&empty_bar[s], MMAW);
      mbar_init(&peer_full[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], MMAW);
      mbar_init(&acc_empty[b], 256);
      mbar_init(&list_full[b], 1);
      mbar_init(&list_empty[b], cluster_ctarank() == 0 ? MMAW : 1);      // leader: its MMA warps; peer: the relay warp
      mbar_init(&xbar[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == V4_MMA_WARP) tmem_alloc2<TMEM_COLS>(tmem_slot);
  for (int c = tid; c < BN; c += NTHREADS) {
    vec[c] = p.scale ? p.scale[c] : 1.f;
    vec[BN + c] = p.shift ? p.shift[c] : 0.f;
  }
  tc_fence_before();
  cluster_sync_all();        // barriers of both CTAs initialised before any remote arrival; TMEM address published
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PWARPS) {
    // ================================================================ producers (as v4, rows of this CTA's half)
    const int r = tid % TILE_M;
    const int thalf = tid / TILE_M;
    const int pc = tid & 7;
    const int rbase = tid >> 3;
    const uint32_t a_thr_off = (uint32_t)(rbase * 128 + ((pc ^ (rbase & 7)) << 4));
    const char* a_bytes = (const char*)p.A;
    uint32_t ca = 0;
    int nv[NVR];
    {
      const long long prow = ((long long)V5_SUPER(pair0) * 2 + rank) * TILE_M + r;
      const bool pok = pair0 < num_super && prow < m_total;
#pragma unroll
      for (int u = 0; u < NVR; ++u) {
        const int t = thalf + TSPLIT * u;
        nv[u] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
      }
    }
    int it = 0;
    for (int sidx = pair0; sidx < num_super; sidx += npairs, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
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
      if (warp == 0) {
        // kernel offsets present in this CTA's rows -> peer; union of both = the chunk list both CTAs walk
        const unsigned own = __ballot_sync(0xffffffffu, lane < p.taps && tap_flag[lane] != 0);
        if (lane == 0) {
          mbar_expect_tx(&xbar[buf], 4);      // this CTA's arrival of the phase + the 4 bytes the peer will send
          st_async_u32(mapa_u32(smem_u32(&tap_flag[48 + buf]), rank ^ 1u), own, mapa_u32(smem_u32(&xbar[buf]), rank ^ 1u));
        }
        mbar_wait_b(&xbar[buf], ubuf & 1);
        const unsigned both = own | (unsigned)((const volatile int*)tap_flag)[48 + buf];
        int count = 0;
        for (int base = 0; base < num_chunks; base += 32) {
          const int q = base + lane;
          bool act = q < num_chunks;
          if (act) {
            const int t0 = (q * BK) / p.cin, t1 = (q * BK + BK - 1) / p.cin;
            bool any = false;
            for (int t = t0; t <= t1 && t < p.taps; ++t) any |= ((both >> t) & 1u) != 0;
            act = any;
          }
          const unsigned mask = __ballot_sync(0xffffffffu, act);
          if (act) chunk_list[count + __popc(mask & ((1u << lane) - 1))] = (uint16_t)q;
          count += __popc(mask);
        }
        if (count == 0) { if (lane == 0) chunk_list[0] = 0; count = 1; }   // keeps the accumulator defined
        if (ubuf > 0) mbar_wait_b(&list_empty[buf], (ubuf - 1) & 1);
        if (lane == 0) { tap_flag[32] = count; tap_flag[40 + buf] = count; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&list_full[buf]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n_act = tap_flag[32];
      {      // prefetch the next tile's neighbour ids
        const int nidx = sidx + npairs;
        const long long prow = ((long long)V5_SUPER(nidx) * 2 + rank) * TILE_M + r;
        const bool pok = nidx < num_super && prow < m_total;
#pragma unroll
        for (int u = 0; u < NVR; ++u) {
          const int t = thalf + TSPLIT * u;
          nv[u] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
        }
      }
      unsigned offs[RPT];
      int offs_tap = -1;
      for (int i = 0; i < n_act; ++i, ++ca) {
        const int q = chunk_list[i];
        const int s = (int)(ca % STAGES);
        const uint32_t use = ca / STAGES;
        if (use > 0) mbar_wait_b(&empty_bar[s], (use - 1) & 1);
        const uint32_t sa = smem_u32(smem + s * SL::STAGE_BYTES);
        if (tid == 0) {
          if ((p.debug & 1) && i > 0) mbar_arrive(&full_bar[s]);
          else {
            mbar_expect_tx(&full_bar[s], (BN / 2) * 128);
            tma_load_2d(sa + SL::A_BYTES, &tmap_w, q * BK, (int)rank * (BN / 2), &full_bar[s]);
          }
        }
        const int k = q * BK + pc * 8;
        const int tap = k >> p.cin_shift, ch = k & (p.cin - 1);
        const bool tap_ok = tap < p.taps;
        const char* src_ch = a_bytes + ch * 2;
        if (tap != offs_tap) {
          offs_tap = tap;
          const int* nrow = nbr_s + tap * TILE_M + rbase;
#pragma unroll
          for (int j = 0; j < RPT; ++j) offs[j] = tap_ok ? (unsigned)nrow[32 * j] : 0xFFFFFFFFu;
        }
        if (!(p.debug & 4))
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const bool ok = offs[j] != 0xFFFFFFFFu;
            cp_async16(sa + a_thr_off + j * 32 * 128, ok ? src_ch + offs[j] : a_bytes, ok ? 16u : 0u);
          }
        cp_async_mbar_arrive(&full_bar[s]);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == V4_MMA_WARP || (MMAW == 2 && warp == MMA_WARP1)) {
    uint32_t ca = 0;
    int it = 0;
    const int sub = warp == V4_MMA_WARP ? 0 : 1;      // MMAW == 2: the sub-tile this warp issues
    if (rank != 0 && sub == 1) {
      // the peer's second MMA warp has nothing to do
    } else if (rank == 0) {
      // ================================================================ leader: MMA issuer for the pair
      constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN);
      for (int sidx = pair0; sidx < num_super; sidx += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t ubuf = (uint32_t)(it >> 1);
        mbar_wait_b(&list_full[buf], ubuf & 1);
        const int n_act = ((const volatile int*)tap_flag)[40 + buf];
        __syncwarp();
        if (lane == 0) mbar_arrive(&list_empty[buf]);
        if (ubuf > 0) mbar_wait_b(&acc_empty[buf], (ubuf - 1) & 1);      // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * MT * BN);
        for (int q = 0; q < n_act; ++q, ++ca) {
          const int s = (int)(ca % STAGES);
          const uint32_t par = (ca / STAGES) & 1;
          mbar_wait_b(&full_bar[s], par);
          mbar_wait_b(&peer_full[s], par);
          fence_proxy_async();
          tc_fence_after();
          {      // whole warp, converged: one election per chunk (umma_chunk, tc_gemm.cuh)
            const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES);
            if constexpr (MMAW == 2) {
              const uint32_t d_sub = d_tmem + (uint32_t)(sub * BN);
              const uint64_t a_sub = make_desc_sw128(a_addr + (uint32_t)(sub * BM * 128));
              umma_chunk<2, 1>(d_sub, d_sub, a_sub, a_sub, make_desc_sw128(a_addr + SL::A_BYTES), idesc, q ? 1u : 0u, &empty_bar[s], 3);
            } else {
              umma_chunk<2, MT>(d_tmem, d_tmem + (uint32_t)BN, make_desc_sw128(a_addr), make_desc_sw128(a_addr + (uint32_t)(BM * 128)),
                                make_desc_sw128(a_addr + SL::A_BYTES), idesc, q ? 1u : 0u, &empty_bar[s], 3);
            }
            if (q == n_act - 1) umma_commit2_mcast(&acc_full[buf], 3);
          }
          __syncwarp();
        }
      }
    } else {
      // ================================================================ peer: relays its full barriers to the leader
      for (int sidx = pair0; sidx < num_super; sidx += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t ubuf = (uint32_t)(it >> 1);
        mbar_wait_b(&list_full[buf], ubuf & 1);
        const int n_act = ((const volatile int*)tap_flag)[40 + buf];
        __syncwarp();
        if (lane == 0) mbar_arrive(&list_empty[buf]);
        for (int q = 0; q < n_act; ++q, ++ca) {
          const int s = (int)(ca % STAGES);
          mbar_wait_b(&full_bar[s], (ca / STAGES) & 1);
          fence_proxy_async();     // this CTA's cp.async writes -> visible to the tensor core before the leader is told
          if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&peer_full[s]), 0));
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================ epilogue (this CTA's 128 x MT rows)
    const int r = (warp & 3) * 32 + lane;
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), 0);
    int it = 0;
    for (int sidx = pair0; sidx < num_super; sidx += npairs, ++it) {
      const int buf = it & 1;
      const uint32_t ubuf = (uint32_t)(it >> 1);
      const long long row_base = ((long long)V5_SUPER(sidx) * 2 + rank) * TILE_M;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const long long grow = row_base + mt * BM + r;
        const bool row_ok = grow < m_total;
        long long orow = grow;
        if (p.out_rows && row_ok) orow = p.out_rows[grow];
        long long rrow = row_ok ? grow : -1;
        if (p.res_rows && row_ok) rrow = p.res_rows[grow];
        const bool store_ok = row_ok && orow >= 0;
        const bool has_res = p.residual != nullptr && rrow >= 0;
        const uint4* rp = has_res ? (const uint4*)(p.residual + rrow * p.ldr) : nullptr;
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
              const __half2* h = (const __half2*)&rcur[j4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
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
              uint4* op = (uint4*)(p.out + orow * p.ldo + cb);
#pragma unroll
              for (int j4 = 0; j4 < 2; ++j4) {
                uint4 u;
                __half2* h = (__half2*)&u;
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1]);
                op[j4] = u;
              }
            }
            if (p.out_f32) {
              float4* op = (float4*)(p.out_f32 + orow * p.ldo + cb);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) op[j4] = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
            }
          }
        }
      }
      tc_fence_before();
      if (rank == 0) mbar_arrive(&acc_empty[buf]);
      else mbar_arrive_remote(acc_empty_leader + (uint32_t)(buf * 8));
    }
  }

  tc_fence_before();
  cluster_sync_all();        // no CTA of the pair exits (or frees TMEM) while the other may still address it
  if (warp == V4_MMA_WARP) tmem_dealloc2<TMEM_COLS>(tmem_base);
#undef V5_SUPER
}

template <int BN, int STAGES, int CTAS, int MT = 1, int MMAW = 1>
static int launch_gemm5(const GemmParams& p, int m_upper, cudaStream_t st) {
  using SL = Smem5<BN, STAGES, MT>;
  static_assert((SL::TOTAL + 1024) * CTAS <= 227 * 1024, "v5 GEMM: shared memory budget exceeded");
  auto kern = gemm5_kernel<BN, STAGES, CTAS, MT, MMAW>;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { isf_set_error("gemm v5: bad device"); return ISF_ERR_CUDA; }
  static bool configured[64] = {};
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
    if (e != cudaSuccess) { isf_set_error("cudaFuncSetAttribute (v5): %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
    configured[dev] = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int rc = make_weight_tmap(&tm, p.W, p.n_total, p.kpad, BN / 2);
  if (rc) return rc;
  int supers = (m_upper + 2 * SL::TILE_M - 1) / (2 * SL::TILE_M);
  if (supers < 1) supers = 1;
  int pairs = isf_num_sms() * CTAS / 2;
  if (pairs < 1) pairs = 1;
  if (pairs > supers) pairs = supers;
  if (::g_isf_gemm_trim) {      // the fewest pairs that finish in the same number of tile rounds (tc_gemm4.cuh)
    const int rounds = (supers + pairs - 1) / pairs;
    pairs = (supers + rounds - 1) / rounds;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs, 1);
  cfg.blockDim = dim3(V4_THREADS + (MMAW - 1) * 32);
  cfg.dynamicSmemBytes = SL::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, tm);
  if (e != cudaSuccess) { isf_set_error("gemm v5 launch: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
  return ISF_OK;
}

// sparse-conv launches whose column tile is the whole Cout (one column block) and that fill the machine with super-tiles
static inline bool gemm5_supported(const GemmParams& p, int bn, int m_upper) {
  return p.nbr != nullptr && p.ln_gamma == nullptr && p.taps > 1 && bn == p.n_total && bn >= 64 && m_upper >= 16 * BM;
}

static inline int dispatch_gemm5(const GemmParams& p, int bn, int m_upper, cudaStream_t st) {
  // ONE CTA per SM in every configuration (256-row CTA tiles for the 64- / 128-column layers): two pairs per TPC each allocating
  // tensor memory with cta_group::2 is a situation no vendor kernel creates, and the one hung bench run of call 21 had it
  // (gemm_mt = 1 selects those two-CTAs-per-SM configurations for A/B runs).
  const bool two = ::g_isf_gemm_mt == 1;
  switch (bn) {
    case 64:
      if (two) return launch_gemm5<64, 4, 2, 1>(p, m_upper, st);
      return ::g_isf_gemm_mma2 ? launch_gemm5<64, 4, 1, 2, 2>(p, m_upper, st) : launch_gemm5<64, 4, 1, 2>(p, m_upper, st);      // 4 x 36 KB + 27 KB
    case 128:
      if (two) return launch_gemm5<128, 4, 2, 1>(p, m_upper, st);
      return ::g_isf_gemm_mma2 ? launch_gemm5<128, 4, 1, 2, 2>(p, m_upper, st) : launch_gemm5<128, 4, 1, 2>(p, m_upper, st);   // 4 x 40 KB + 27 KB
    case 256: return launch_gemm5<256, 6, 1, 1>(p, m_upper, st);   // 6 x 32 KB + 14 KB, all 512 TMEM columns
    default: isf_set_error("gemm v5: unsupported column tile %d", bn); return ISF_ERR_UNSUPPORTED;
  }
}

}  // namespace tc
