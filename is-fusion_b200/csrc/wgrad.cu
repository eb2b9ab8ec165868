// Training: weight gradient of a sparse convolution,
//     dW[k][ci][co] = sum over the pairs p of kernel offset k of  X[in[p]][ci] * G[out[p]][co]
// (reference: spconv's indice_conv_backward: per-offset gather of the input and output-gradient rows followed by a GEMM,
// bevfusion-ops/spconv/include/spconv/spconv_ops.h:364-470) as ONE kernel: no gathered copies of X / G in global memory.
//
// Grid = (kernel offsets x B x B blocks of the Cin x Cout matrix, pair chunks of a kernel offset), B = 32 / 64 / 128 by the layer's
// width.  A CTA walks its chunk of pairs in stages: the X rows and G rows of a stage are gathered by 16-byte cp.async into shared
// memory (4 stages, prefetch distance 3: a gather is two dependent global loads, index then row), and 8 warps run
// mma.sync.m16n8k16 with the PAIR index as K.  A = X^T and B = G are both stored pair-major, so their fragments come from
// ldmatrix.x4.trans (the pitch of B + 8 halfs puts the 8 rows of a matrix on 8 different 16-byte bank groups).  Narrow layers
// split the pairs of a stage over warp groups instead of the channels (B = 32: 8 groups of 16 pairs, B = 64: 2 x 32, B = 128: all
// warps share the 32 pairs).  The fp32 accumulators are added to dW with red.global.add (dW is zero on entry; the chunks and the
// warp groups of a kernel offset and nothing else contend for an address).
// fp16 operands (the gradient arrives scaled into fp16's range by the caller), fp32 accumulation.
#include <cuda_fp16.h>
#include <stdint.h>

#include "common.cuh"

extern "C" void isf_count_launch(int n);

namespace wg {

constexpr int THREADS = 256;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four 8x8 fp16 matrices, transposed on the way: thread (g = lane / 4, t = lane % 4) receives {M[2t][g], M[2t+1][g]} of each
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], const __half* row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cp16(void* dst, const void* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16u : 0u) : "memory");
}

// B x B block of the Cin x Cout matrix per CTA; 8 warps = (B / WM) x (B / WN) warp tiles x KG groups that split the pairs of a stage
// (KPG pairs per group and stage), STAGES shared-memory stages of KG * KPG pairs.
template <int B, int WM, int WN, int KG, int KPG, int STAGES>
struct Cfg {
  static constexpr int PITCH = B + 8;                 // halfs: rows 16 bytes apart modulo 128 -> conflict-free ldmatrix
  static constexpr int KST = KG * KPG;                // pairs per stage
  static constexpr int PIECES = B / 8;                // 16-byte pieces per row
  static constexpr int SMEM = STAGES * 2 * KST * PITCH * 2;
  static_assert((B / WM) * (B / WN) * KG * 32 == THREADS, "8 warps");
  static_assert(KPG % 16 == 0 && WM % 16 == 0 && WN % 16 == 0, "mma tiles");
};

template <int B, int WM, int WN, int KG, int KPG, int STAGES>
__global__ void __launch_bounds__(THREADS) wgrad_kernel(const __half* __restrict__ X, int cin, const __half* __restrict__ G, int cout,
                                                        const int* __restrict__ in_rows, const int* __restrict__ out_rows,
                                                        const int* __restrict__ tap_off, int taps, int chunk, float* __restrict__ dW) {
  using C = Cfg<B, WM, WN, KG, KPG, STAGES>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __half* xs = reinterpret_cast<__half*>(smem_raw);                       // [STAGES][KST][PITCH]
  __half* gs = xs + STAGES * C::KST * C::PITCH;
  // blockIdx.x = kernel offset + taps * channel block (fastest), blockIdx.y = pair chunk: the CTAs in flight together work on the
  // same stretch of output sites under every kernel offset, i.e. on the same G rows and on neighbouring X rows -> they meet in L2
  // (the first version walked chunk-fastest and read 4.2 GB from DRAM for 275 MB of operands, profiles/ncu_raw_wgrad_1_r02.csv)
  const int tap = blockIdx.x % taps, blk = blockIdx.x / taps;
  // every kernel offset is cut into gridDim.y chunks of ITS OWN pair count (chunk = the host's bound for the longest list): chunk c
  // of every offset then covers about the same stretch of output sites, and no CTA of a short list is empty
  const int t0 = tap_off[tap], n_k = tap_off[tap + 1] - t0;
  int ck = min(chunk, (n_k + (int)gridDim.y - 1) / (int)gridDim.y);
  ck = (ck + Cfg<B, WM, WN, KG, KPG, STAGES>::KST - 1) / Cfg<B, WM, WN, KG, KPG, STAGES>::KST * Cfg<B, WM, WN, KG, KPG, STAGES>::KST;
  const int p_begin = t0 + blockIdx.y * ck;
  const int p_end = min(t0 + n_k, p_begin + ck);
  if (p_begin >= p_end) return;
  const int nb = (cout + B - 1) / B;
  const int m0 = (blk / nb) * B, n0 = (blk % nb) * B;                     // channel block of this CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int WMN = (B / WM) * (B / WN);
  const int kg = warp / WMN, wmn = warp % WMN;
  const int wm = (wmn / (B / WN)) * WM, wn = (wmn % (B / WN)) * WN;       // warp tile inside the block
  const bool active = m0 + wm < cin && n0 + wn < cout;                    // warp-uniform: tiles past the matrix do nothing
  float acc[WM / 16][WN / 8][4];
#pragma unroll
  for (int i = 0; i < WM / 16; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  // a gather is two dependent global loads (pair index, then the row): the indices of a stage are loaded one iteration before its
  // rows are requested, so only the prologue waits for them
  constexpr int NL = C::KST * C::PIECES / THREADS;
  static_assert(NL * THREADS == C::KST * C::PIECES, "whole passes of the block over a stage");
  int ri[NL], ro[NL];
  auto load_idx = [&](int p0) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int p = p0 + (tid + l * THREADS) / C::PIECES;
      const bool okp = p < p_end;
      ri[l] = okp ? __ldg(in_rows + p) : -1;
      ro[l] = okp ? __ldg(out_rows + p) : -1;
    }
  };
  auto issue = [&](int buf) {
    __half* xb = xs + buf * C::KST * C::PITCH;
    __half* gb = gs + buf * C::KST * C::PITCH;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int i = tid + l * THREADS;
      const int pr = i / C::PIECES, lc = (i % C::PIECES) * 8;
      const bool okp = ri[l] >= 0;
      cp16(xb + pr * C::PITCH + lc, X + (size_t)(okp ? ri[l] : 0) * cin + m0 + lc, okp && m0 + lc < cin);
      cp16(gb + pr * C::PITCH + lc, G + (size_t)(okp ? ro[l] : 0) * cout + n0 + lc, okp && n0 + lc < cout);
    }
  };

  const int stages = (p_end - p_begin + C::KST - 1) / C::KST;
#pragma unroll
  for (int d = 0; d < STAGES - 1; ++d) {
    if (d < stages) { load_idx(p_begin + d * C::KST); issue(d); }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  load_idx(p_begin + (STAGES - 1) * C::KST);       // past the chunk: -1s
  // ldmatrix row of this lane: matrix q = lane / 8, row r = lane % 8
  const int q = lane >> 3, r = lane & 7;
  for (int s = 0; s < stages; ++s) {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    __syncthreads();                       // stage s has landed for every thread; the buffer of stage s - 1 is free
    if (s + STAGES - 1 < stages) issue((s + STAGES - 1) % STAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    load_idx(p_begin + (s + STAGES) * C::KST);     // consumed by the next iteration's issue: the latency hides behind the MMAs below
    if (!active) continue;
    const __half* xb = xs + (s % STAGES) * C::KST * C::PITCH + kg * KPG * C::PITCH;
    const __half* gb = gs + (s % STAGES) * C::KST * C::PITCH + kg * KPG * C::PITCH;
#pragma unroll
    for (int k0 = 0; k0 < KPG; k0 += 16) {
      uint32_t a[WM / 16][4];
#pragma unroll
      for (int i = 0; i < WM / 16; ++i)     // matrices: (k 0-7, m), (k 0-7, m+8), (k 8-15, m), (k 8-15, m+8) = a0..a3
        ldsm4_t(a[i], xb + (k0 + r + (q >> 1) * 8) * C::PITCH + wm + i * 16 + (q & 1) * 8);
#pragma unroll
      for (int jj = 0; jj < WN / 16; ++jj) {
        uint32_t b[4];                      // matrices: (k 0-7, n), (k 8-15, n), (k 0-7, n+8), (k 8-15, n+8) = b0,b1 of two column tiles
        ldsm4_t(b, gb + (k0 + r + (q & 1) * 8) * C::PITCH + wn + jj * 16 + (q >> 1) * 8);
#pragma unroll
        for (int i = 0; i < WM / 16; ++i) {
          mma16816(acc[i][2 * jj], a[i], b[0], b[1]);
          mma16816(acc[i][2 * jj + 1], a[i], b[2], b[3]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!active) return;
  const int g = lane >> 2, t = lane & 3;
  float* dw = dW + (size_t)tap * cin * cout;
#pragma unroll
  for (int i = 0; i < WM / 16; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) {
      const int m = m0 + wm + i * 16 + g, n = n0 + wn + j * 8 + 2 * t;
      if (n < cout) {       // cout is a multiple of 8: n and n + 1 are valid together
        if (m < cin) { atomicAdd(dw + (size_t)m * cout + n, acc[i][j][0]); atomicAdd(dw + (size_t)m * cout + n + 1, acc[i][j][1]); }
        if (m + 8 < cin) { atomicAdd(dw + (size_t)(m + 8) * cout + n, acc[i][j][2]); atomicAdd(dw + (size_t)(m + 8) * cout + n + 1, acc[i][j][3]); }
      }
    }
}

template <int B, int WM, int WN, int KG, int KPG, int STAGES>
static int launch(const void* x, int cin, const void* g, int cout, const int* in_rows, const int* out_rows, const int* tap_offsets,
                  int taps, int max_pairs, float* dw, cudaStream_t stream) {
  using C = Cfg<B, WM, WN, KG, KPG, STAGES>;
  auto kern = wgrad_kernel<B, WM, WN, KG, KPG, STAGES>;
  ISF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  // chunks of pairs per kernel offset: enough CTAs to fill the machine, at least 512 pairs each (every chunk ends in one atomic add
  // per accumulator)
  const int blocks = ((cin + B - 1) / B) * ((cout + B - 1) / B);
  int chunks = (4 * isf_num_sms() + taps * blocks - 1) / (taps * blocks);
  if (chunks < 1) chunks = 1;
  int chunk = (max_pairs + chunks - 1) / chunks;
  if (chunk < 512) chunk = 512;
  chunk = (chunk + C::KST - 1) / C::KST * C::KST;
  chunks = (max_pairs + chunk - 1) / chunk;
  kern<<<dim3(taps * blocks, chunks), THREADS, C::SMEM, stream>>>((const __half*)x, cin, (const __half*)g, cout, in_rows, out_rows,
                                                                  tap_offsets, taps, chunk, dw);
  isf_count_launch(1);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

}  // namespace wg

// see include/isfusion_b200.h
extern "C" int isf_spconv_wgrad(const void* x, int cin, const void* g, int cout, const int* in_rows, const int* out_rows,
                                const int* tap_offsets, int taps, int max_pairs_per_tap, float* dw, void* stream) {
  ISF_CHECK_ARG(cin >= 8 && cin % 8 == 0 && cout >= 8 && cout % 8 == 0 && taps >= 1 && taps <= 27 && max_pairs_per_tap >= 0);
  if (max_pairs_per_tap == 0) return ISF_OK;
  ISF_CHECK_ARG(x && g && in_rows && out_rows && tap_offsets && dw);
  ISF_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)g & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int c = cin > cout ? cin : cout;
  if (c <= 32) return wg::launch<32, 32, 32, 8, 16, 4>(x, cin, g, cout, in_rows, out_rows, tap_offsets, taps, max_pairs_per_tap, dw, st);
  if (c <= 64) return wg::launch<64, 32, 32, 2, 32, 4>(x, cin, g, cout, in_rows, out_rows, tap_offsets, taps, max_pairs_per_tap, dw, st);
  return wg::launch<128, 32, 64, 1, 32, 4>(x, cin, g, cout, in_rows, out_rows, tap_offsets, taps, max_pairs_per_tap, dw, st);
}
