// Bring-up probe for TMA tile::gather4 (NOT on the product path): gathers 128 rows x 64 fp16 columns of a row-major matrix
// into one smem tile the way the implicit-GEMM producer would, and dumps the raw shared memory.  `variant` bits:
//   1 = ".shared::cta ... .cta_group::1" instruction form (else ".shared::cluster"), 2 = every 4-row gather lands on its own
//   1024-byte block (else consecutive 512-byte blocks), 4 = un-swizzled tensor map (else SWIZZLE_128B).
#include "tc_gemm.cuh"

template <int FORM>
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  if constexpr (FORM == 0)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"((uint64_t)tm), "r"(col), "r"(r0), "r"(r1), "r"(r2),
                 "r"(r3), "r"(tc::smem_u32(bar))
                 : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1"
                 " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"((uint64_t)tm), "r"(col), "r"(r0), "r"(r1), "r"(r2),
                 "r"(r3), "r"(tc::smem_u32(bar))
                 : "memory");
}

template <int FORM>
__global__ void __launch_bounds__(128) gather4_probe_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                            int col0, int dst_stride, uint4* __restrict__ out, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bar = (uint64_t*)(smem + 32768);
  for (int e = threadIdx.x; e < 32768 / 16; e += blockDim.x) ((uint4*)smem)[e] = make_uint4(0x47004700u, 0x47004700u, 0x47004700u, 0x47004700u);
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(bar, 128 * 128);
    for (int g = 0; g < 32; ++g)
      gather4<FORM>(tc::smem_u32(smem + g * dst_stride), &tm, col0, idx[4 * g], idx[4 * g + 1], idx[4 * g + 2], idx[4 * g + 3], bar);
  }
  int ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it) ok = tc::mbar_try(bar, 0) ? 1 : 0;
  if (threadIdx.x == 0) *status = ok;
  __syncthreads();
  for (int e = threadIdx.x; e < 32768 / 16; e += blockDim.x) out[e] = ((const uint4*)smem)[e];
}

extern "C" int isf_debug_tma_gather4(const void* a, int rows, int cols, const int* idx_dev, int col0, int variant, void* out32k,
                                     int* status_dev, void* stream) {
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn) { isf_set_error("no cuTensorMapEncodeTiled"); return ISF_ERR_CUDA; }
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, (variant & 4) ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("encode failed %d", (int)r); return ISF_ERR_CUDA; }
  const int dst_stride = (variant & 2) ? 1024 : 512;
  if (variant & 1) {
    ISF_CUDA(cudaFuncSetAttribute(gather4_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000));
    gather4_probe_kernel<1><<<1, 128, 36000, (cudaStream_t)stream>>>(tm, idx_dev, col0, dst_stride, (uint4*)out32k, status_dev);
  } else {
    ISF_CUDA(cudaFuncSetAttribute(gather4_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000));
    gather4_probe_kernel<0><<<1, 128, 36000, (cudaStream_t)stream>>>(tm, idx_dev, col0, dst_stride, (uint4*)out32k, status_dev);
  }
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark (NOT on the product path): how fast can one SM pull 128 gathered rows x 128 B into shared memory?
//   mode 0: tile::gather4, one warp, every lane issues its own instruction (32 per chunk)
//   mode 1: tile::gather4, one elected lane issues all 32
//   mode 2: tile::gather4, four warps, every lane of the first 8 lanes of each... (8 instructions per warp)
//   mode 3: plain 2-D TMA tile load of 128 CONSECUTIVE rows (one instruction per chunk) — the no-gather reference
//   mode 4: LDGSTS (16-byte cp.async, 256 threads, 4 pieces each — the implicit-GEMM producer's pattern)
// A ring of STAGES 16 KB stages; a stage is re-issued as soon as its previous load has landed (no consumer): the result is
// the sustained load rate with STAGES chunks in flight.  out[blockIdx.x] = cycles per chunk.
// ------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(256) gather_bench_kernel(const __grid_constant__ CUtensorMap tm_row,
                                                           const __grid_constant__ CUtensorMap tm_tile, const __half* __restrict__ a,
                                                           int cols, const int* __restrict__ idx, int n_idx, int mode, int iters,
                                                           float* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bar = (uint64_t*)(smem + STAGES * 16384);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) tc::mbar_init(&bar[s], mode == 4 ? 256 : 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int s = it % STAGES;
    const uint32_t u = it / STAGES;
    const uint32_t dst = tc::smem_u32(smem + s * 16384);
    const int* ix = idx + ((size_t)(blockIdx.x * 977 + it) * 128) % (size_t)(n_idx - 128);
    if (mode == 4) {
      if (u > 0) tc::mbar_wait_b(&bar[s], (u - 1) & 1);
      const int pc = tid & 7, rbase = tid >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ix[rbase + 32 * j];
        tc::cp_async16(dst + (rbase + 32 * j) * 128 + ((pc ^ (rbase & 7)) << 4), (const char*)a + ((size_t)(r < 0 ? 0 : r) * cols + pc * 8) * 2,
                       r < 0 ? 0u : 16u);
      }
      tc::cp_async_mbar_arrive(&bar[s]);
    } else if (mode == 0) {
      if (warp == 0) {
        if (u > 0) tc::mbar_wait_b(&bar[s], (u - 1) & 1);
        if (lane == 0) tc::mbar_expect_tx(&bar[s], 16384);
        __syncwarp();
        const int4 r = *(const int4*)(ix + 4 * lane);
        gather4<0>(dst + lane * 512, &tm_row, 0, r.x, r.y, r.z, r.w, &bar[s]);
      }
    } else if (mode == 1) {
      if (tid == 0) {
        if (u > 0) tc::mbar_wait_b(&bar[s], (u - 1) & 1);
        tc::mbar_expect_tx(&bar[s], 16384);
        for (int g = 0; g < 32; ++g) {
          const int4 r = *(const int4*)(ix + 4 * g);
          gather4<0>(dst + g * 512, &tm_row, 0, r.x, r.y, r.z, r.w, &bar[s]);
        }
      }
    } else if (mode == 2) {
      if (warp < 4) {
        if (u > 0) tc::mbar_wait_b(&bar[s], (u - 1) & 1);
        if (tid == 0) tc::mbar_expect_tx(&bar[s], 16384);
        __syncwarp();
        if (lane < 8) {
          const int g = warp * 8 + lane;
          const int4 r = *(const int4*)(ix + 4 * g);
          gather4<0>(dst + g * 512, &tm_row, 0, r.x, r.y, r.z, r.w, &bar[s]);
        }
      }
    } else {
      if (tid == 0) {
        if (u > 0) tc::mbar_wait_b(&bar[s], (u - 1) & 1);
        tc::mbar_expect_tx(&bar[s], 16384);
        int r0 = ix[0];
        if (r0 < 0) r0 = 0;
        tc::tma_load_2d(dst, &tm_tile, 0, r0, &bar[s]);
      }
    }
  }
  // drain: wait for the last use of every stage
  if (mode == 4 || tid == 0 || (mode == 0 && warp == 0) || (mode == 2 && warp < 4)) {
    for (int s = 0; s < STAGES; ++s) {
      const int uses = (iters - s + STAGES - 1) / STAGES;
      if (uses > 0) tc::mbar_wait_b(&bar[s], (uint32_t)(uses - 1) & 1);
    }
  }
  if (mode == 4) asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)iters;
}

extern "C" int isf_debug_gather_bench(const void* a, int rows, int cols, const int* idx_dev, int n_idx, int mode, int stages,
                                      int iters, int blocks, float* out_cycles, void* stream) {
  ISF_CHECK_ARG(a && idx_dev && out_cycles && rows >= 128 && cols >= 64 && n_idx > 256 && mode >= 0 && mode <= 4 && iters >= 1);
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn) { isf_set_error("no cuTensorMapEncodeTiled"); return ISF_ERR_CUDA; }
  CUtensorMap tr, tt;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t box1[2] = {64, 1}, box128[2] = {64, 128};
  CUresult r = fn(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a), dims, strides, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("encode failed %d", (int)r); return ISF_ERR_CUDA; }
  r = fn(&tt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a), dims, strides, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("encode failed %d", (int)r); return ISF_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
#define ISF_GB(S)                                                                                                         \
  {                                                                                                                       \
    ISF_CUDA(cudaFuncSetAttribute(gather_bench_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * 16384 + 2048)); \
    gather_bench_kernel<S><<<blocks, 256, S * 16384 + 2048, st>>>(tr, tt, (const __half*)a, cols, idx_dev, n_idx, mode, iters,  \
                                                                  out_cycles);                                           \
  }
  if (stages == 4) ISF_GB(4) else if (stages == 8) ISF_GB(8) else if (stages == 12) ISF_GB(12) else {
    isf_set_error("gather bench: stages in {4, 8, 12}");
    return ISF_ERR_INVALID;
  }
#undef ISF_GB
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark (NOT on the product path): issue rate of tcgen05.mma kind::f16 by instruction shape and operand source.
// One thread per CTA (per CTA pair for cta_group::2) issues `iters` groups of four K = 16 instructions of shape m x n on constant
// (zero) shared-memory operands, four groups in flight (tcgen05.commit ring); nothing else runs.  out[block] = cycles per
// instruction as seen by that CTA.  a_tmem = 1: the A operand comes from tensor memory (TS mode) instead of shared memory;
// chains = independent accumulators the instructions of one issuing thread alternate between; issuers = warps of the CTA that
// each issue their own stream (own accumulators, own commit ring).  TMEM columns allocated = what that needs (128 / 256 / 512).
// With probe != 0 (one CTA, m x 32 x 16, cta_group::1) it runs ONE instruction on A[i][0] = i + 1, B[j][0] = 1 and dumps TMEM
// column 0 of all 128 lanes: the accumulator layout of the shape (out_probe[lane] = D value held by that lane).
// ------------------------------------------------------------------------------------------------
#include "tc_gemm5.cuh"

template <int CG, int COLS>
__global__ void __launch_bounds__(128) mma_rate_kernel(int m, int n, int a_tmem, int chains, int issuers, int iters,
                                                       float* __restrict__ out, float* __restrict__ out_probe) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;                  // 256 rows x 128 B
  uint8_t* sb = smem + 32768;          // 256 rows x 128 B
  uint64_t* bar = (uint64_t*)(smem + 65536);      // [4 issuers][4]
  uint32_t* slot = (uint32_t*)(bar + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool probe = out_probe != nullptr;
  for (int e = tid; e < 65536 / 16; e += 128) ((uint4*)smem)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (probe) {
    // row r, k = 0 of the 128B-swizzled K-major layout sits at r * 128 + ((r & 7) << 4)
    for (int r = tid; r < 256; r += 128) {
      *(__half*)(sa + r * 128 + ((r & 7) << 4)) = __float2half((float)(r + 1));
      *(__half*)(sb + r * 128 + ((r & 7) << 4)) = __float2half(1.f);
    }
  }
  if (tid == 0) {
    for (int s = 0; s < 16; ++s) tc::mbar_init(&bar[s], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (CG == 2) tc::tmem_alloc2<COLS>(slot); else tc::tmem_alloc<COLS>(slot);
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  if constexpr (CG == 2) tc::cluster_sync_all(); else __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  const bool leader = CG == 1 || tc::cluster_ctarank() == 0;
  // warp w < issuers issues its own instruction stream into its own accumulators (columns [(w chains + c) n, + n)): lane 0 inside
  // `if (lane == 0)` (the library's MMA warps), or - a_tmem bit 1 - the whole warp converged with the instruction predicated by
  // elect.sync (no per-instruction ELECT / BRA.U.ANY loop in the SASS)
  const bool elect_mode = (a_tmem & 6) != 0, group_mode = (a_tmem & 4) != 0;
  a_tmem &= 1;
  if (elect_mode && warp < issuers && leader) {
    uint64_t* mybar = bar + 4 * warp;
    const uint32_t idesc = tc::make_idesc_f16(CG * m, n);
    const uint64_t a_desc = tc::make_desc_sw128(tc::smem_u32(sa)), b_desc = tc::make_desc_sw128(tc::smem_u32(sb));
    const uint32_t d_t = tmem + (uint32_t)(warp * n);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i & 3;
      if (i >= 4) tc::mbar_wait_b(&mybar[s], (uint32_t)((i - 4) >> 2) & 1);
      if (group_mode) {      // one election per group of four instructions + commit (tc::umma_chunk: what the library issues)
        tc::umma_chunk<CG, 1>(d_t, d_t, a_desc, a_desc, b_desc, idesc, 1u, &mybar[s], 1);
        continue;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if constexpr (CG == 2)
          asm volatile(
              "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
              "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, 1;\n\t}" ::"r"(d_t),
              "l"(a_desc + (uint64_t)(kk * 2)), "l"(b_desc + (uint64_t)(kk * 2)), "r"(idesc)
              : "memory");
        else
          asm volatile(
              "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
              "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t}" ::"r"(d_t),
              "l"(a_desc + (uint64_t)(kk * 2)), "l"(b_desc + (uint64_t)(kk * 2)), "r"(idesc)
              : "memory");
      }
      if constexpr (CG == 2)
        asm volatile(
            "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
                tc::smem_u32(&mybar[s])),
            "h"((uint16_t)1)
            : "memory");
      else
        asm volatile(
            "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(tc::smem_u32(&mybar[s]))
            : "memory");
    }
    for (int i = (iters > 4 ? iters - 4 : 0); i < iters; ++i) tc::mbar_wait_b(&mybar[i & 3], (uint32_t)(i >> 2) & 1);
    const long long t1 = clock64();
    if (warp == 0 && lane == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)(iters * 4);
  } else if (lane == 0 && warp < issuers && leader) {
    uint64_t* mybar = bar + 4 * warp;
    const uint32_t idesc = tc::make_idesc_f16(CG * m, n);
    const uint64_t a_desc = tc::make_desc_sw128(tc::smem_u32(sa)), b_desc = tc::make_desc_sw128(tc::smem_u32(sb));
    const uint32_t a_t = tmem + (uint32_t)(COLS - 32);      // TS mode: A in the last 32 columns
    const int groups = probe ? 1 : iters;
    const long long t0 = clock64();
    for (int i = 0; i < groups; ++i) {
      const int s = i & 3;
      if (i >= 4) tc::mbar_wait_b(&mybar[s], (uint32_t)((i - 4) >> 2) & 1);
      for (int kk = 0; kk < (probe ? 1 : 4); ++kk)
        for (int c = 0; c < chains; ++c) {
          const uint32_t d_t = tmem + (uint32_t)((warp * chains + c) * n);
          const uint32_t acc = (probe && kk == 0) ? 0u : 1u;
          if (a_tmem) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_t),
                "r"(a_t + (uint32_t)(kk * 8)), "l"(b_desc + (uint64_t)(kk * 2)), "r"(idesc), "r"(acc)
                : "memory");
          } else if constexpr (CG == 2) {      // the un-elected issue form the library used before call 20 (single active lane)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_t),
                "l"(a_desc + (uint64_t)(kk * 2)), "l"(b_desc + (uint64_t)(kk * 2)), "r"(idesc), "r"(acc)
                : "memory");
          } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_t),
                "l"(a_desc + (uint64_t)(kk * 2)), "l"(b_desc + (uint64_t)(kk * 2)), "r"(idesc), "r"(acc)
                : "memory");
          }
        }
      if constexpr (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         tc::smem_u32(&mybar[s])),
                     "h"((uint16_t)1)
                     : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&mybar[s]))
                     : "memory");
    }
    for (int i = (groups > 4 ? groups - 4 : 0); i < groups; ++i) tc::mbar_wait_b(&mybar[i & 3], (uint32_t)(i >> 2) & 1);
    const long long t1 = clock64();
    if (warp == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)(groups * 4 * (probe ? 1 : chains));
  }
  tc::tc_fence_before();
  if constexpr (CG == 2) tc::cluster_sync_all(); else __syncthreads();
  tc::tc_fence_after();
  if (probe) {
    uint32_t v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc::tmem_ld_wait();
    out_probe[tid] = __uint_as_float(v[0]);
    out_probe[128 + tid] = __uint_as_float(v[1]);
  }
  tc::tc_fence_before();
  if constexpr (CG == 2) tc::cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if constexpr (CG == 2) tc::tmem_dealloc2<COLS>(tmem); else tc::tmem_dealloc<COLS>(tmem);
  }
}

extern "C" int isf_debug_mma_rate(int m, int n, int cta_group, int a_tmem, int chains, int issuers, int blocks, int iters,
                                  float* out_cycles, float* out_probe, void* stream) {
  const int need = chains * issuers * n + ((a_tmem & 1) ? 32 : 0);
  ISF_CHECK_ARG(out_cycles && (m == 64 || m == 128) && n >= 16 && n <= 256 && n % 16 == 0 && (cta_group == 1 || cta_group == 2) &&
                chains >= 1 && issuers >= 1 && issuers <= 4 && need <= 512 && blocks >= 1 && iters >= 1 && (!(a_tmem & 1) || cta_group == 1) &&
                (!(a_tmem & 6) || (chains == 1 && !(a_tmem & 1) && !out_probe)) &&
                (!out_probe || (cta_group == 1 && blocks == 1)));
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 65536 + 1024 + 256;
  const int cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);      // TMEM columns per CTA: bounds the CTAs per SM
  if (cta_group == 2) {
    ISF_CHECK_ARG(blocks % 2 == 0);
    auto kern = cols == 128 ? mma_rate_kernel<2, 128> : (cols == 256 ? mma_rate_kernel<2, 256> : mma_rate_kernel<2, 512>);
    ISF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ISF_CUDA(cudaLaunchKernelEx(&cfg, kern, m, n, a_tmem, chains, issuers, iters, out_cycles, out_probe));
  } else {
    auto kern = cols == 128 ? mma_rate_kernel<1, 128> : (cols == 256 ? mma_rate_kernel<1, 256> : mma_rate_kernel<1, 512>);
    ISF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<blocks, 128, smem, st>>>(m, n, a_tmem, chains, issuers, iters, out_cycles, out_probe);
  }
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}
