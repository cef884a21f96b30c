// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is fetched through
// cudaGetDriverEntryPoint so the library does not link against libcuda at build time.
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_common.cuh"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point not available"); return VITMARL_ECUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch_bytes & 15) || box_cols * 2 != 128 || box_rows > 256) {
    set_last_error("tensor map: misaligned base/pitch or unsupported box");
    return VITMARL_EINVAL;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed");
    return VITMARL_ECUDA;
  }
  return VITMARL_OK;
}

}  // namespace vitmarl
