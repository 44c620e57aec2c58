// Host-side TMA tensor-map construction. The driver entry point is resolved at run time through the CUDA runtime so
// the shared library has no link-time dependency on libcuda (it must dlopen on a GPU-less build box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace ada {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// bf16 tensor map, 128-byte swizzle, zero fill out of bounds. dims/box innermost first; strides in BYTES for dims 1..rank-1.
// elem_strides (optional): traversal stride per dimension; a box of `box[i]` tensor elements then delivers
// ceil(box[i] / elem_strides[i]) elements (every elem_strides[i]-th one) -- the stride-2 implicit-GEMM convolution.
// swizzle128 = false: dense (un-swizzled) box, inner box dimension up to 256 elements.
inline CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                  const uint32_t* box, const uint32_t* elem_strides = nullptr, bool swizzle128 = true) {
  CUtensorMap m;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                               const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string msg = "cuTensorMapEncodeTiled failed: code " + std::to_string(static_cast<int>(r)) + " rank " +
                      std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string(dims[i]);
    msg += " box";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string(box[i]);
    throw std::runtime_error(msg);
  }
  return m;
}

inline CUtensorMap make_tmap_2d(const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_elems, uint32_t box_inner,
                                uint32_t box_rows) {
  uint64_t dims[2] = {inner, rows};
  uint64_t str[1] = {pitch_elems * 2};
  uint32_t box[2] = {box_inner, box_rows};
  return make_tmap_bf16(base, 2, dims, str, box);
}

// fp32 row-major [rows, inner] map with 32 x 32 boxes (128-byte rows, 128-byte swizzle): the in-place residual epilogue
inline CUtensorMap make_tmap_f32_2d(const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_elems, uint32_t box_inner,
                                    uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstr[1] = {pitch_elems * 4};
  cuuint32_t bx[2] = {box_inner, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (fp32) failed: code " + std::to_string(static_cast<int>(r)));
  return m;
}

}  // namespace ada
