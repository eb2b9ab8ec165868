// Bring-up probe for TMA tile::gather4 (not on the product path): gathers 128 rows x 64 fp16 columns of a
// row-major matrix into one 128B-swizzled smem tile exactly like the GEMM producer would, and dumps the raw smem.
#include "tc_gemm.cuh"

__global__ void __launch_bounds__(128) gather4_probe_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                            int col0, uint4* __restrict__ out, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 16384);
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(bar, 128 * 128);
    for (int g = 0; g < 32; ++g) {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
          " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(tc::smem_u32(smem + g * 512)),
          "l"((uint64_t)&tm), "r"(col0), "r"(idx[4 * g]), "r"(idx[4 * g + 1]), "r"(idx[4 * g + 2]), "r"(idx[4 * g + 3]),
          "r"(tc::smem_u32(bar))
          : "memory");
    }
  }
  // bounded wait (a wrong byte count must not hang the box)
  int ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it) {
    uint32_t p;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(p) : "r"(tc::smem_u32(bar)) : "memory");
    ok = (int)p;
  }
  if (threadIdx.x == 0) *status = ok;
  __syncthreads();
  for (int e = threadIdx.x; e < 16384 / 16; e += blockDim.x) out[e] = ((const uint4*)smem)[e];
}

extern "C" int isf_debug_tma_gather4(const void* a, int rows, int cols, const int* idx_dev, int col0, int box_rows,
                                     void* out16k, int* status_dev, void* stream) {
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn) { isf_set_error("no cuTensorMapEncodeTiled"); return ISF_ERR_CUDA; }
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { isf_set_error("encode failed %d", (int)r); return ISF_ERR_CUDA; }
  ISF_CUDA(cudaFuncSetAttribute(gather4_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000));
  gather4_probe_kernel<<<1, 128, 20000, (cudaStream_t)stream>>>(tm, idx_dev, col0, (uint4*)out16k, status_dev);
  ISF_LAUNCH_CHECK();
  return ISF_OK;
}
