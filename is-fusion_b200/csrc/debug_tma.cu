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
