// tcgen05 implicit-GEMM building block (sm_100a).
//
//   D[m, n] = epilogue( sum_k A[row(m, k), k % Cin] * W[n, k] )        k in [0, taps*Cin)
//
// A is an fp16 feature matrix [rows, Cin]; `row(m,k)` is either m (plain GEMM, taps = 1) or
// nbr[(k / Cin) * nbr_stride + m] (gathered rows of a sparse-convolution rulebook, -1 = no neighbour
// -> zero).  W is pre-packed K-major [N, Kpad] fp16 (Kpad = taps*Cin rounded up to 64).
//
// One CTA = 128 output rows x BN output columns.  Warps 0-3 are the producers (one A row per thread,
// 16-byte cp.async pieces written straight into the 128B-swizzled K-major layout the UMMA smem
// descriptor expects) and, after the main loop, the epilogue (TMEM lane = output row = thread).
// Warp 4 allocates TMEM and its elected lane issues tcgen05.mma (M=128, N=BN, K=16) and the
// tcgen05.commit that recycles smem stages through mbarriers.  Accumulators live in TMEM (fp32).
#pragma once
#include <cuda.h>
#include <string.h>
#include "common.cuh"

namespace tc {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
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
// bounded wait (~2 s): a protocol error traps instead of hanging the box (used by kernels under bring-up)
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the barrier receives one arrival when all cp.async issued so far by this thread have landed (no pending-count
// increment: the arrival is part of the barrier's expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16, fp32 accumulate).
// ISSUE FORM (round 2, scripts/bench_mma_rate.py, profiles/call20_r02.log): these wrappers are called by ALL 32 lanes of a
// converged warp and predicate the instruction with elect.sync.  Inside `if (lane == 0) { ... }` the compiler wraps every UTCHMMA
// in an ELECT / BRA.U.ANY loop and a thread then issues ONE instruction per ~179 cycles whatever its shape (128 x 32 ... 256 x 256);
// in the elected form the same thread issues one per 86 cycles and the tensor pipe runs at its nominal 128 N / 256 cycles per
// instruction (N = 256: 128.0 cycles measured).  elect.sync with a full mask always elects the same lane, so tcgen05.commit (which
// tracks the MMAs of the executing thread) sees the instructions it must wait for.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

// One K chunk (BK = 64 = four K = 16 steps) of MT (1 | 2) sub-tiles behind ONE elect.sync: the 4 MT instructions (K step outermost,
// sub-tiles alternating) and the tcgen05.commit that releases the shared-memory stage.  Per-instruction election costs ~86 cycles
// of issue per instruction (bench_mma_rate.py), more than the 64 cycles a 128 x 128 instruction occupies the tensor pipe.
// CG = cta_group (2: the commit is multicast to the CTAs of `mask`).  Whole warp, converged.
template <int CG, int MT>
__device__ __forceinline__ void umma_chunk(uint32_t d0, uint32_t d1, uint64_t a0, uint64_t a1, uint64_t b, uint32_t idesc,
                                           uint32_t acc_first, uint64_t* commit_bar, uint16_t mask = 3) {
  static_assert((CG == 1 || CG == 2) && (MT == 1 || MT == 2), "umma_chunk");
#define ISF_MMA1 "@pe tcgen05.mma.cta_group::1.kind::f16 "
#define ISF_MMA2 "@pe tcgen05.mma.cta_group::2.kind::f16 "
#define ISF_CHUNK_HEAD                                                                              \
  "{\n\t.reg .pred pe, p0;\n\t.reg .b64 x1, x2, x3, y1, y2, y3, b1, b2, b3;\n\t"                  \
  "elect.sync _|pe, 0xffffffff;\n\t"                                                                \
  "setp.ne.b32 p0, %6, 0;\n\t"                                                                      \
  "add.u64 x1, %2, 2;\n\tadd.u64 x2, %2, 4;\n\tadd.u64 x3, %2, 6;\n\t"                             \
  "add.u64 y1, %3, 2;\n\tadd.u64 y2, %3, 4;\n\tadd.u64 y3, %3, 6;\n\t"                             \
  "add.u64 b1, %4, 2;\n\tadd.u64 b2, %4, 4;\n\tadd.u64 b3, %4, 6;\n\t"
#define ISF_CHUNK_ARGS                                                                              \
  ::"r"(d0), "r"(d1), "l"(a0), "l"(a1), "l"(b), "r"(idesc), "r"(acc_first), "r"(smem_u32(commit_bar)), "h"(mask) : "memory"
  if constexpr (CG == 1 && MT == 1)
    asm volatile(ISF_CHUNK_HEAD
                 ISF_MMA1 "[%0], %2, %4, %5, p0;\n\t" ISF_MMA1 "[%0], x1, b1, %5, 1;\n\t"
                 ISF_MMA1 "[%0], x2, b2, %5, 1;\n\t" ISF_MMA1 "[%0], x3, b3, %5, 1;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}" ISF_CHUNK_ARGS);
  else if constexpr (CG == 1 && MT == 2)
    asm volatile(ISF_CHUNK_HEAD
                 ISF_MMA1 "[%0], %2, %4, %5, p0;\n\t" ISF_MMA1 "[%1], %3, %4, %5, p0;\n\t"
                 ISF_MMA1 "[%0], x1, b1, %5, 1;\n\t" ISF_MMA1 "[%1], y1, b1, %5, 1;\n\t"
                 ISF_MMA1 "[%0], x2, b2, %5, 1;\n\t" ISF_MMA1 "[%1], y2, b2, %5, 1;\n\t"
                 ISF_MMA1 "[%0], x3, b3, %5, 1;\n\t" ISF_MMA1 "[%1], y3, b3, %5, 1;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}" ISF_CHUNK_ARGS);
  else if constexpr (CG == 2 && MT == 1)
    asm volatile(ISF_CHUNK_HEAD
                 ISF_MMA2 "[%0], %2, %4, %5, p0;\n\t" ISF_MMA2 "[%0], x1, b1, %5, 1;\n\t"
                 ISF_MMA2 "[%0], x2, b2, %5, 1;\n\t" ISF_MMA2 "[%0], x3, b3, %5, 1;\n\t"
                 "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], %8;\n\t}"
                 ISF_CHUNK_ARGS);
  else
    asm volatile(ISF_CHUNK_HEAD
                 ISF_MMA2 "[%0], %2, %4, %5, p0;\n\t" ISF_MMA2 "[%1], %3, %4, %5, p0;\n\t"
                 ISF_MMA2 "[%0], x1, b1, %5, 1;\n\t" ISF_MMA2 "[%1], y1, b1, %5, 1;\n\t"
                 ISF_MMA2 "[%0], x2, b2, %5, 1;\n\t" ISF_MMA2 "[%1], y2, b2, %5, 1;\n\t"
                 ISF_MMA2 "[%0], x3, b3, %5, 1;\n\t" ISF_MMA2 "[%1], y3, b3, %5, 1;\n\t"
                 "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], %8;\n\t}"
                 ISF_CHUNK_ARGS);
#undef ISF_MMA1
#undef ISF_MMA2
#undef ISF_CHUNK_HEAD
#undef ISF_CHUNK_ARGS
}

// ---- thread-block-cluster helpers (weight-chunk multicast)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D TMA tile load, multicast to every CTA of `mask`: data and complete_tx land at the same CTA-relative
// offsets (dst, bar) in each destination CTA.
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"((uint64_t)tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"((uint64_t)tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {      // whole warp, elected lane (see umma_f16)
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (base_lane + i), columns col..col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns of fp32
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B smem matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format F16 (0) @7/@10, K-major A and B,
// n_dim = N>>3 @17, m_dim = M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t IDESC_AB_BF16 = (1u << 7) | (1u << 10);      // a_format / b_format = BF16 (1) instead of F16 (0)

// ---------------------------------------------------------------- kernel
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct GemmParams {
  const __half* A;        // [rowsA, Cin]
  const __half* W;        // [N_total, Kpad] K-major
  const int* nbr;         // [taps, nbr_stride] or nullptr (identity rows)
  long long nbr_stride;
  const int* m_dev;       // optional device-side row count (overrides m when non-null)
  int m;                  // output rows
  int cin;                // channels per tap (multiple of 8)
  int cin_shift;          // log2(cin) when taps > 1 (cin is then a power of two), else unused
  int taps;
  int kpad;               // padded K (multiple of 64)
  int n_total;            // total output columns (B rows); the CTA covers [blockIdx.y*BN, +BN)
  // epilogue
  const float* scale;     // [N_total] or nullptr
  const float* shift;     // [N_total] or nullptr
  const __half* residual; // [m, ldr] or nullptr (added before the activation)
  int ldr;
  const int* res_rows;    // optional gather of the residual row: residual[res_rows[m]] (-1 = none)
  int act;
  const float* ln_gamma;  // LayerNorm over the BN columns after residual (requires BN == n_total)
  const float* ln_beta;
  float ln_eps;
  __half* out;            // [m, ldo] fp16 (may be nullptr)
  float* out_f32;         // [m, ldo] fp32 (may be nullptr)
  int ldo;
  const int* out_rows;    // optional scatter: output row index per m (e.g. window -> flat), -1 = skip
  int* tile_ctr;          // tc_gemm4 / tc_gemm5: int[2] zeroed device buffer for dynamic tile scheduling (isf_spconv_tile_counter), or nullptr
  int bf16;               // 16-bit operands / residual / output are bf16 (else fp16); dense GEMMs only (isf_gemm_bf16)
  int wpc;                // tc_gemm4.cuh: producer warps per K chunk (1 or 2) of the warp-per-chunk mapping
  int rev;                // tc_gemm4.cuh: tiles walked from the last one down (isf_set_option("gemm_rev"))
  int zskip;              // tc_gemm4.cuh: skip all-missing 4-row copies onto rows that already hold zeros (isf_set_option("gemm_zskip"))
  int debug;              // timing experiments only (isf_set_option("gemm_debug")): 1 = weight chunk loaded for the first chunk of a
                          // tile only, 2 = every gathered row reads row 0, 4 = no A copies at all.  Results are then meaningless.
};

constexpr int BM = 128;
constexpr int BK = 64;          // 64 fp16 = 128 B = one swizzle row
constexpr int PRODUCERS = 256;  // 8 warps gather A; all of them also run the epilogue (two column halves)
constexpr int MAX_TAPS = 27;
constexpr int MAX_CHUNKS = 256;     // K <= 16384
constexpr int THREADS = PRODUCERS + 32;
constexpr int PWARPS = PRODUCERS / 32;

template <int BN, int STAGES, int MT = 1>
struct SmemLayout {
  static constexpr int A_BYTES = MT * BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NBR_OFF = STAGES * STAGE_BYTES;            // [MAX_TAPS][MT*128] int32 neighbour tile
  static constexpr int LIST_OFF = NBR_OFF + MAX_TAPS * MT * BM * 4;     // active chunk list (uint16) + flags
  static constexpr int VEC_OFF = LIST_OFF + MAX_CHUNKS * 2 + 64 * 4;     // epilogue scale[BN] | shift[BN] fp32 (column tiles <= 128: the
  static constexpr int VEC_BYTES = BN <= 128 ? 2 * BN * 4 : 0;           // 256-column tile has no room left at 2 CTAs / SM)
  static constexpr int BAR_OFF = VEC_OFF + VEC_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // barriers + tmem slot, + alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// one 32-column block of the accumulator row owned by this thread: TMEM -> registers, BN fold (scale / shift) and the
// residual row (four 16-byte loads).  Shared by the LayerNorm statistics pass and the output pass, so both see
// bit-identical values.
// 256-bit global accesses (LDG/STG.E.ENL2.256 on sm_100): one whole 32-byte sector per lane and instruction.  The
// address must be 32-byte aligned.
__device__ __forceinline__ void ld_global_256(const void* ptr, uint32_t (&w)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(ptr));
}
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// IO256 (v2 kernel only, when the residual rows are 32-byte aligned): the residual row comes in two 32-byte loads
// instead of four 16-byte ones.  Same values, same order of the additions.
template <bool IO256 = false>
__device__ __forceinline__ void epilogue_block(const GemmParams& p, uint32_t taddr, int n_base, long long rrow, float (&x)[32]) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float y = __uint_as_float(v[j]);
    if (p.scale) y = __fmul_rn(y, p.scale[n_base + j]);      // explicit roundings: every epilogue variant of the library
    if (p.shift) y = __fadd_rn(y, p.shift[n_base + j]);      // (fast / LayerNorm / v3) produces the same bits
    x[j] = y;
  }
  if (p.residual && rrow >= 0) {
    if constexpr (IO256) {
#pragma unroll
      for (int j8 = 0; j8 < 2; ++j8) {
        uint32_t w[8];
        ld_global_256(p.residual + rrow * p.ldr + n_base + j8 * 16, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 f = isf_unpack2(w[e], p.bf16 != 0);
          x[j8 * 16 + e * 2] += f.x; x[j8 * 16 + e * 2 + 1] += f.y;
        }
      }
    } else {
      const uint4* rp = (const uint4*)(p.residual + rrow * p.ldr + n_base);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const uint4 u = rp[j4];
        const uint32_t* h = (const uint32_t*)&u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = isf_unpack2(h[e], p.bf16 != 0);
          x[j4 * 8 + e * 2] += f.x; x[j4 * 8 + e * 2 + 1] += f.y;
        }
      }
    }
  }
}

// CS = thread-block-cluster size.  CS > 1: the CTAs of a cluster walk the same K-chunk sequence on CS consecutive
// row tiles; each CTA TMA-loads 1/CS of the weight chunk and multicasts it to the whole cluster, so weight
// traffic from L2 drops by CS; smem stages are recycled cluster-wide (tcgen05.commit multicast on the empty barrier).
// MT = 128-row sub-tiles per CTA tile (one TMEM accumulator each): with MT = 2 one weight chunk feeds 256 output rows, which
// halves the weight bytes an SM pulls from L2 per FLOP (the 128/256-column layers are paced by the SM's L2 read port).
// LN = LayerNorm-epilogue instantiation (dense GEMMs only: the neighbour-tile region of shared memory holds the per-column
// vectors and the row-statistics exchange): both warp halves take half of the row's columns, ONE pass over TMEM with the
// half row kept in registers, partial (sum, sum of squares) exchanged through shared memory.
template <int BN, int STAGES, int CS, int MT, bool LN>
__global__ void __launch_bounds__(THREADS, (LN || MT > 1) ? 1 : (BN <= 64 ? 3 : 2)) gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_w) {
  using SL = SmemLayout<BN, STAGES, MT>;
  static_assert(CS == 1 || MT == 1, "clusters only with single sub-tile CTAs");
  constexpr int TILE_M = BM * MT;
  constexpr int TSPLIT = PRODUCERS / TILE_M;        // thread groups that split the kernel offsets in the tile prologue
  constexpr int NVR = (MAX_TAPS + TSPLIT - 1) / TSPLIT;
  constexpr int RPT = TILE_M * 8 / PRODUCERS;       // A rows per producer thread and chunk
  constexpr uint16_t CMASK = (uint16_t)((1u << CS) - 1);
  constexpr int SLICE_ROWS = BN / CS;
  constexpr int ACC_COLS = BN * MT;
  static_assert(ACC_COLS <= 512, "TMEM holds 512 columns");
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offsetting the __shared__ array itself (keeps the shared address space visible to the
  // compiler: LDS/STS instead of generic LD/ST on every smem access)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = (uint64_t*)(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_total = p.m_dev ? *p.m_dev : p.m;
  const int num_tiles = (m_total + TILE_M - 1) / TILE_M;
  const int n0 = blockIdx.y * BN;
  const int num_chunks = p.kpad / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], PRODUCERS + 1);                  // + the expect_tx arrive of the weight-TMA issuer
      mbar_init(&empty_bar[s], CS);                             // one tcgen05.commit per CTA of the cluster
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == PWARPS) tmem_alloc<TMEM_COLS>(tmem_slot);
  if constexpr (!LN && SL::VEC_BYTES > 0) {      // BN scale / shift of this CTA's columns -> shared memory, once per CTA
    float* vec = (float*)(smem + SL::VEC_OFF);
    for (int c = tid; c < BN; c += THREADS) {
      vec[c] = p.scale ? p.scale[n0 + c] : 1.f;
      vec[BN + c] = p.shift ? p.shift[n0 + c] : 0.f;
    }
  }
  if constexpr (LN) {       // per-column vectors of this CTA's columns -> shared memory (read there by every tile's epilogue)
    float* vec = (float*)(smem + SL::NBR_OFF);
    for (int c = tid; c < BN; c += THREADS) {
      vec[c] = p.scale ? p.scale[n0 + c] : 1.f;
      vec[BN + c] = p.shift ? p.shift[n0 + c] : 0.f;
      vec[2 * BN + c] = p.ln_gamma[n0 + c];
      vec[3 * BN + c] = p.ln_beta[n0 + c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();       // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
  const int num_groups = (num_tiles + CS - 1) / CS;

  uint32_t chunk_ctr = 0;   // running chunk counter across tiles (stage = ctr % STAGES, use = ctr / STAGES)
  uint32_t tile_ctr = 0;
  int* nbr_s = (int*)(smem + SL::NBR_OFF);
  uint16_t* chunk_list = (uint16_t*)(smem + SL::LIST_OFF);
  int* tap_flag = (int*)(smem + SL::LIST_OFF + MAX_CHUNKS * 2);   // [32] flags + [1] active count at index 32
  const bool gather = p.nbr != nullptr;

  for (int grp = cluster_id; grp < num_groups; grp += num_clusters, ++tile_ctr) {
    const int tile = grp * CS + crank;     // may lie beyond the last tile: the CTA then only feeds the multicast
    const int row0 = tile * TILE_M;
    // ---------------------------------------------------------------- tile prologue: neighbour tile -> smem,
    // list of K-chunks that have at least one neighbour in this tile (empty kernel offsets are skipped)
    if (warp < PWARPS) {
      if (gather) {
        const int r = tid % TILE_M;            // row of the tile; the thread groups split the kernel offsets
        const int thalf = tid / TILE_M;
        if (tid < 33) tap_flag[tid] = 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll 1
        for (int cs = 0; cs < CS; ++cs) {       // flags = union over the cluster's tiles (same chunk list everywhere)
          const long long prow = (long long)(grp * CS + cs) * TILE_M + r;
          const bool pok = prow < m_total;
          // all neighbour loads of this row are issued before any is consumed (one exposed latency, not 14)
          int nv[NVR];
#pragma unroll
          for (int u = 0; u < NVR; ++u) {
            const int t = thalf + TSPLIT * u;
            nv[u] = (pok && t < p.taps) ? p.nbr[(long long)t * p.nbr_stride + prow] : -1;
          }
#pragma unroll
          for (int u = 0; u < NVR; ++u) {
            const int t = thalf + TSPLIT * u;
            if (t < p.taps) {
              // staged as the BYTE offset of the neighbour's feature row (0xFFFFFFFF = none): one add per piece later
              if (cs == crank) nbr_s[t * TILE_M + r] = nv[u] >= 0 ? (int)((unsigned)nv[u] * (unsigned)(p.cin * 2)) : -1;
              if (nv[u] >= 0) tap_flag[t] = 1;
            }
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
        if (lane == 0) tap_flag[32] = count;
      }
    }
    __syncthreads();
    const int n_act = tap_flag[32];
    if (warp < PWARPS) {
      // ------------------------------------------------ producers
      const int r = tid & (BM - 1);            // epilogue row of this thread inside a sub-tile (TMEM lane = 32*(warp%4) + lane)
      // A tile = 4 warp-coalesced sweeps: thread -> 16-byte piece c = tid%8 of rows rr_j = tid/8 + 32 j, so one warp
      // instruction fetches 4 whole 128 B rows; (rr_j & 7) is the same for every j -> the swizzled offset is a constant
      const int pc = tid & 7;
      const int rbase = tid >> 3;
      const uint32_t a_thr_off = (uint32_t)(rbase * 128 + ((pc ^ (rbase & 7)) << 4));
      const char* a_bytes = (const char*)p.A;
      uint32_t c0 = chunk_ctr;
      int s = (int)(c0 % STAGES);
      uint32_t use = c0 / STAGES;
      // byte offsets of this thread's 4 rows for the kernel offset it is currently gathering: re-read from the smem
      // neighbour tile only when the offset changes (Cin = 128 / 256 spend 2 / 4 chunks on one offset)
      unsigned offs[RPT];
      int offs_tap = -1;
      for (int i = 0; i < n_act; ++i) {
        const int q = chunk_list[i];
        if (use > 0) mbar_wait_b(&empty_bar[s], (use - 1) & 1);
        const uint32_t sa = smem_u32(smem + s * SL::STAGE_BYTES);
        const uint32_t b_base = sa + SL::A_BYTES;
        // weights: one TMA box (128B swizzle = the UMMA layout); with clusters each CTA loads 1/CS and multicasts
        if (tid == 0) {
          if ((p.debug & 1) && i > 0) mbar_arrive(&full_bar[s]);
          else {
          mbar_expect_tx(&full_bar[s], BN * 128);
          if constexpr (CS == 1) tma_load_2d(b_base, &tmap_w, q * BK, n0, &full_bar[s]);
          else tma_load_2d_mcast(b_base + crank * SLICE_ROWS * 128, &tmap_w, q * BK, n0 + crank * SLICE_ROWS, &full_bar[s], CMASK);
          }
        }
        {
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
                if (gather) off = (unsigned)nrow[(PRODUCERS / 8) * j];
                else { const int gr = row0 + rbase + (PRODUCERS / 8) * j; if (gr < m_total) off = (unsigned)gr * (unsigned)(p.cin * 2); }
              }
              if ((p.debug & 2) && off != 0xFFFFFFFFu) off = 0;
              offs[j] = off;
            }
          }
          if (!(p.debug & 4))
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const bool ok = offs[j] != 0xFFFFFFFFu;
            cp_async16(sa + a_thr_off + j * (PRODUCERS / 8) * 128, ok ? src_ch + offs[j] : a_bytes, ok ? 16u : 0u);
          }
        }
        // asynchronous arrival: the stage's full barrier is signalled by the hardware when this thread's pieces have
        // landed, so the thread runs ahead to the next chunk instead of waiting for its own loads
        cp_async_mbar_arrive(&full_bar[s]);
        if (++s == STAGES) { s = 0; ++use; }
      }
      // ------------------------------------------------ epilogue
      // warps 0-3 own columns [0, BN/2), warps 4-7 [BN/2, BN) of the same rows; the row-local LayerNorm needs the
      // whole row in one thread and stays on warps 0-3
      if constexpr (LN) {
        static_assert(!LN || (BN >= 64 && MT == 1), "LayerNorm epilogue: 64..256 columns, single sub-tile");
        mbar_wait_b(accum_bar, tile_ctr & 1);
        tc_fence_after();
        constexpr int HC = BN / 2;                       // columns per thread
        const int half = warp >> 2;
        const float* vec = (const float*)(smem + SL::NBR_OFF);
        float2* part = (float2*)(smem + SL::NBR_OFF + 4 * BN * 4);     // [2][128] partial (sum, sum of squares)
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * HC);
        const int grow = row0 + r;
        const bool row_ok = grow < m_total;
        long long orow = grow;
        if (p.out_rows && row_ok) orow = p.out_rows[grow];
        const bool store_ok = row_ok && orow >= 0;
        long long rrow = row_ok ? (long long)grow : -1;
        if (p.res_rows && row_ok) rrow = p.res_rows[grow];
        float x[HC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int cb = 0; cb < HC; cb += 32) {
          uint32_t v[32];
          tmem_ld32(t_lane + cb, v);
          tmem_ld_wait();
          const float* vs = vec + half * HC + cb;
          const float* vh = vec + BN + half * HC + cb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = __uint_as_float(v[j]);
            if (p.scale) y = __fmul_rn(y, vs[j]);
            if (p.shift) y = __fadd_rn(y, vh[j]);
            x[cb + j] = y;
          }
          if (p.residual && rrow >= 0) {
            const uint4* rp = (const uint4*)(p.residual + rrow * p.ldr + n0 + half * HC + cb);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 u = rp[j4];
              const uint32_t* h = (const uint32_t*)&u;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = isf_unpack2(h[e], p.bf16 != 0);
                x[cb + j4 * 8 + e * 2] += f.x; x[cb + j4 * 8 + e * 2 + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1 += x[cb + j]; s2 += x[cb + j] * x[cb + j]; }
        }
        part[half * BM + r] = make_float2(s1, s2);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float2 o = part[(half ^ 1) * BM + r];
        // both halves add the two partials in the same order (low half first): identical statistics on both sides
        const float t1 = half ? o.x + s1 : s1 + o.x, t2 = half ? o.y + s2 : s2 + o.y;
        const float mean = t1 / BN;
        const float var = t2 / BN - mean * mean;
        const float rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
        const float* ga = vec + 2 * BN + half * HC;
        const float* be = vec + 3 * BN + half * HC;
#pragma unroll
        for (int cb = 0; cb < HC; cb += 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = (x[cb + j] - mean) * rstd * ga[cb + j] + be[cb + j];
            if (p.act == ACT_RELU) y = fmaxf(y, 0.f);
            else if (p.act == ACT_GELU) y = gelu_erf(y);
            x[cb + j] = y;
          }
          if (store_ok) {
            if (p.out) {
              uint4* op = (uint4*)(p.out + orow * p.ldo + n0 + half * HC + cb);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                uint4 u;
                uint32_t* h = (uint32_t*)&u;
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = isf_pack2(x[cb + j4 * 8 + e * 2], x[cb + j4 * 8 + e * 2 + 1], p.bf16 != 0);
                op[j4] = u;
              }
            }
            if (p.out_f32) {
              float4* op = (float4*)(p.out_f32 + orow * p.ldo + n0 + half * HC + cb);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) op[j4] = make_float4(x[cb + j4 * 4], x[cb + j4 * 4 + 1], x[cb + j4 * 4 + 2], x[cb + j4 * 4 + 3]);
            }
          }
        }
        tc_fence_before();
      } else {
      const bool ln = p.ln_gamma != nullptr;
      const int cb_lo = (ln || BN < 64) ? 0 : (warp >> 2) * (BN / 2);
      const int cb_hi = (ln || BN < 64) ? ((warp >> 2) ? 0 : BN) : cb_lo + BN / 2;
      if (!ln) {
        // pipelined epilogue (no LayerNorm): the row maps and the first residual block are fetched BEFORE the wait for the
        // accumulator, the residual of block i+1 while block i is processed, scale / shift come from shared memory (<= 128
        // columns) — the round-2 source-level profile had 20 % of the kernel's stall samples on these dependent global loads
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * BN);
        const int grow = row0 + mt * BM + r;
        const bool row_ok = grow < m_total;
        long long orow = grow;
        if (p.out_rows && row_ok) orow = p.out_rows[grow];
        long long rrow = row_ok ? (long long)grow : -1;
        if (p.res_rows && row_ok) rrow = p.res_rows[grow];
        const bool store_ok = row_ok && orow >= 0;
        const bool has_res = p.residual != nullptr && rrow >= 0 && cb_lo < cb_hi;
        const uint4* rp = has_res ? (const uint4*)(p.residual + rrow * p.ldr + n0 + cb_lo) : nullptr;
        uint4 rnext[2];                                   // 16-column blocks keep the live set small (2-3 CTAs / SM)
        if (has_res) { rnext[0] = rp[0]; rnext[1] = rp[1]; }
        if (mt == 0) {
          mbar_wait_b(accum_bar, tile_ctr & 1);
          tc_fence_after();
        }
#pragma unroll 1
        for (int cb = cb_lo; cb < cb_hi; cb += 16) {
          uint4 rcur[2];
          if (has_res) {
            rcur[0] = rnext[0]; rcur[1] = rnext[1];
            if (cb + 16 < cb_hi) { rnext[0] = rp[(cb + 16 - cb_lo) / 8]; rnext[1] = rp[(cb + 16 - cb_lo) / 8 + 1]; }
          }
          uint32_t v[16];
          tmem_ld16(t_lane + cb, v);
          tmem_ld_wait();
          float x[16];
          if constexpr (SL::VEC_BYTES > 0) {
            const float* vs = (const float*)(smem + SL::VEC_OFF) + cb;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float y = __uint_as_float(v[j]);
              if (p.scale) y = __fmul_rn(y, vs[j]);
              if (p.shift) y = __fadd_rn(y, vs[BN + j]);
              x[j] = y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float y = __uint_as_float(v[j]);
              if (p.scale) y = __fmul_rn(y, p.scale[n0 + cb + j]);
              if (p.shift) y = __fadd_rn(y, p.shift[n0 + cb + j]);
              x[j] = y;
            }
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
      } else {
      mbar_wait_b(accum_bar, tile_ctr & 1);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * BN);
      const int grow = row0 + mt * BM + r;
      const bool row_ok = grow < m_total;
      long long orow = grow;
      if (p.out_rows && row_ok) orow = p.out_rows[grow];
      const bool store_ok = row_ok && orow >= 0;
      long long rrow = row_ok ? (long long)grow : -1;
      if (p.res_rows && row_ok) rrow = p.res_rows[grow];
      float mean = 0.f, rstd = 1.f;
      if (ln && warp < 4) {
        // pass 1 over TMEM: row statistics of (acc*scale+shift+residual)
        float s1 = 0.f, s2 = 0.f;
        for (int cb = 0; cb < BN; cb += 32) {
          float x[32];
          epilogue_block(p, t_lane + cb, n0 + cb, rrow, x);
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1 += x[j]; s2 += x[j] * x[j]; }
        }
        mean = s1 / BN;
        float var = s2 / BN - mean * mean;
        rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
      }
      for (int cb = cb_lo; cb < cb_hi; cb += 32) {
        float x[32];
        epilogue_block(p, t_lane + cb, n0 + cb, rrow, x);
        if (p.ln_gamma) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = (x[j] - mean) * rstd * p.ln_gamma[n0 + cb + j] + p.ln_beta[n0 + cb + j];
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
              uint32_t* h = (uint32_t*)&u;
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = isf_pack2(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1], p.bf16 != 0);
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
      }
      // all TMEM reads of this tile are ordered before the __syncthreads below, after which the MMA
      // warp may overwrite the accumulator with the next tile
      tc_fence_before();
      }
    } else {
      // ------------------------------------------------ MMA issuer (warp PWARPS)
      const uint32_t idesc = make_idesc_f16(BM, BN) | (p.bf16 ? IDESC_AB_BF16 : 0u);
      uint32_t c0 = chunk_ctr;
      for (int q = 0; q < n_act; ++q) {
        const uint32_t cc = c0 + q;
        const int s = cc % STAGES;
        mbar_wait_b(&full_bar[s], (cc / STAGES) & 1);
        fence_proxy_async();     // cp.async (generic proxy) writes of the producers -> visible to the tensor core
        tc_fence_after();
        {      // whole warp, converged: the wrappers elect the issuing lane
          const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES);
          const uint32_t b_addr = a_addr + SL::A_BYTES;
          const uint64_t b_desc = make_desc_sw128(b_addr);
          if constexpr (CS == 1) {
            umma_chunk<1, MT>(tmem_base, tmem_base + (uint32_t)BN, make_desc_sw128(a_addr), make_desc_sw128(a_addr + (uint32_t)(BM * 128)),
                              b_desc, idesc, q ? 1u : 0u, &empty_bar[s]);
          } else {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the >>4 address field
                umma_f16(tmem_base + (uint32_t)(mt * BN), make_desc_sw128(a_addr + (uint32_t)(mt * BM * 128)) + (uint64_t)(kk * 2),
                         b_desc + (uint64_t)(kk * 2), idesc, (q | kk) ? 1u : 0u);
              }
            }
            umma_commit_mcast(&empty_bar[s], CMASK);
          }
          if (q == n_act - 1) umma_commit(accum_bar);
        }
        __syncwarp();
      }
    }
    chunk_ctr += n_act;
    // The epilogue of tile t must finish reading TMEM before the MMA warp starts tile t+1.
    __syncthreads();
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();       // nobody leaves while a peer may still multicast into / signal this CTA
  if (warp == PWARPS) tmem_dealloc<TMEM_COLS>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// W [n_total, kpad] fp16 K-major; box = 64 K-elements (128 B, one swizzle row) x slice_rows
static int make_weight_tmap(CUtensorMap* tm, const void* w, int n_total, int kpad, int slice_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { isf_set_error("cuTensorMapEncodeTiled not available"); return ISF_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)n_total};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)slice_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return ISF_ERR_CUDA; }
  return ISF_OK;
}


template <int BN, int STAGES, int CS, int MT = 1, bool LN = false>
static int launch_gemm(const GemmParams& p, int m_upper, cudaStream_t st) {
  using SL = SmemLayout<BN, STAGES, MT>;
  static_assert(SL::TOTAL <= 227 * 1024, "gemm: shared memory budget exceeded");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { isf_set_error("gemm: bad device"); return ISF_ERR_CUDA; }
  static bool configured[64] = {};          // cudaFuncSetAttribute is per-device state
  static int sm_count[64] = {};
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, STAGES, CS, MT, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
    if (e != cudaSuccess) { isf_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
    e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || sm_count[dev] < 1) { isf_set_error("gemm: cannot query the SM count"); return ISF_ERR_CUDA; }
    configured[dev] = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  {
    int rc = make_weight_tmap(&tm, p.W, p.n_total, p.kpad, BN / CS);
    if (rc) return rc;
  }
  int tiles = (m_upper + BM * MT - 1) / (BM * MT);
  if (tiles < 1) tiles = 1;
  int ctas_per_sm = (227 * 1024) / (SL::TOTAL + 1024);
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  if (ctas_per_sm * BN * MT > 512) ctas_per_sm = 512 / (BN * MT);      // TMEM columns
  int n_tiles = p.n_total / BN;
  int groups = (tiles + CS - 1) / CS;
  int gx = sm_count[dev] * ctas_per_sm / n_tiles / CS;     // clusters
  if (gx < 1) gx = 1;
  if (gx > groups) gx = groups;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(gx * CS, n_tiles);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SL::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<BN, STAGES, CS, MT, LN>, p, tm);
  if (e != cudaSuccess) { isf_set_error("gemm launch: %s", cudaGetErrorString(e)); return ISF_ERR_CUDA; }
  return ISF_OK;
}

template <int CS>
static inline int dispatch_gemm_cs(const GemmParams& p, int bn, int m_upper, cudaStream_t st) {
  switch (bn) {
    case 32: return launch_gemm<32, 4, CS>(p, m_upper, st);
    case 64: return launch_gemm<64, 2, CS>(p, m_upper, st);     // 3 CTAs/SM beat a deeper ring (measured)
    case 128: return launch_gemm<128, 3, CS>(p, m_upper, st);
    case 256: return launch_gemm<256, 2, CS>(p, m_upper, st);
    default: isf_set_error("gemm: unsupported column tile %d", bn); return ISF_ERR_UNSUPPORTED;
  }
}

// 256-row CTA tiles (two accumulators, one weight chunk per 256 rows): isf_set_option("gemm_mt", 2), sparse-conv launches
static inline int dispatch_gemm_mt2(const GemmParams& p, int bn, int m_upper, cudaStream_t st) {
  switch (bn) {
    case 32: return launch_gemm<32, 4, 1, 2>(p, m_upper, st);      // 4 x 36 KB + 27 KB
    case 64: return launch_gemm<64, 3, 1, 2>(p, m_upper, st);      // 3 x 40 KB + 27 KB: 1 CTA/SM
    case 128: return launch_gemm<128, 3, 1, 2>(p, m_upper, st);    // 3 x 48 KB + 27 KB
    case 256: return launch_gemm<256, 3, 1, 2>(p, m_upper, st);    // 3 x 64 KB + 27 KB = 221 KB, all 512 TMEM columns
    default: return dispatch_gemm_cs<1>(p, bn, m_upper, st);
  }
}

static inline bool gemm4_supported(const GemmParams& p, int bn);                               // tc_gemm4.cuh
static inline int dispatch_gemm4(const GemmParams& p, int bn, int m_upper, cudaStream_t st);
static inline bool gemm5_supported(const GemmParams& p, int bn, int m_upper);                  // tc_gemm5.cuh
static inline int dispatch_gemm5(const GemmParams& p, int bn, int m_upper, cudaStream_t st);

static inline int dispatch_gemm(const GemmParams& p0, int bn, int m_upper, cudaStream_t st) {
  if (p0.taps > MAX_TAPS || p0.kpad / BK > MAX_CHUNKS) {
    isf_set_error("gemm: taps=%d (max %d) / K chunks=%d (max %d) out of range", p0.taps, MAX_TAPS, p0.kpad / BK, MAX_CHUNKS);
    return ISF_ERR_UNSUPPORTED;
  }
  GemmParams p = p0;
  p.debug = ::g_isf_gemm_debug;
  p.zskip = ::g_isf_gemm_zskip;
  p.rev = ::g_isf_gemm_rev;
  p.wpc = 1;
  p.bf16 = (::isf_bf16_mode() && p.nbr == nullptr) ? 1 : 0;
  if (p.ln_gamma != nullptr && p.nbr == nullptr && ::g_isf_gemm_ln_fast) {      // LayerNorm epilogue: dense GEMMs, one column block
    switch (bn) {
      case 64: return launch_gemm<64, 2, 1, 1, true>(p, m_upper, st);
      case 128: return launch_gemm<128, 3, 1, 1, true>(p, m_upper, st);
      case 256: return launch_gemm<256, 2, 1, 1, true>(p, m_upper, st);
      default: break;
    }
  }
  // CTA pairs on tcgen05.mma.cta_group::2 (tc_gemm5.cuh): sparse-conv launches with >= 64 columns
  if (::g_isf_gemm_pair && gemm5_supported(p, bn, m_upper) && (::g_isf_gemm_pair == 1 || bn >= ::g_isf_gemm_pair))
    return dispatch_gemm5(p, bn, m_upper, st);
  // dedicated epilogue warps + double-buffered accumulator (tc_gemm4.cuh): 1 = sparse-conv launches, 2 = dense GEMMs too
  if (::g_isf_gemm_v4 && gemm4_supported(p, bn) && m_upper >= 8 * BM && (p.nbr != nullptr || ::g_isf_gemm_v4 == 2))
    return dispatch_gemm4(p, bn, m_upper, st);
  // 256-row CTA tiles (one weight chunk per 256 rows): measured -5..-7 % on the 256-column layers of stage 4 (round 2, with the
  // pipelined epilogue), neutral or slower elsewhere and slower below two waves of tiles -> the default (3) takes them only there
  if (p.nbr != nullptr && m_upper >= 148 * 256 &&
      (::g_isf_gemm_mt == 2 || (::g_isf_gemm_mt == 3 && bn == 256 && m_upper >= 2 * 148 * 256)))
    return dispatch_gemm_mt2(p, bn, m_upper, st);
  switch (::g_isf_gemm_cluster) {
    case 2: return dispatch_gemm_cs<2>(p, bn, m_upper, st);
    default: break;
  }
  // measured: sharing the weight chunk in a CTA pair pays only for the 32-column tiles
  if (bn == 32 && m_upper > 4 * BM) return dispatch_gemm_cs<2>(p, bn, m_upper, st);
  return dispatch_gemm_cs<1>(p, bn, m_upper, st);
}

}  // namespace tc

#include "tc_gemm4.cuh"
#include "tc_gemm5.cuh"
