// Host-side construction of TMA tensor maps for (batch, head, row, feature) tensors whose
// feature dimension is contiguous.  The driver entry point is resolved at run time through
// cudaGetDriverEntryPoint so the library does not link against libcuda.
//
// This is what replaces the reference's strided PackedTensorAccessor32 indexing
// (flash_cosine_sim_attention_cuda.cu:30-35): arbitrary (b, h, n) strides are encoded in
// the tensor map, the kernels never compute element addresses for q/k/v/do themselves.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace fcsa {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Tensor map over a 16-bit tensor addressed as [b][h][n][d] with element strides
// (sb, sh, sn, 1).  The box is box_cols features x box_rows rows: 64 features = one 128-byte swizzle
// span for the operand tiles, 16 features = one 32-byte span (swizzle_bytes = 32) for the K = 16 slivers.
// A dimension of extent 1 gets a dummy stride (the driver rejects zero strides).
// Returns 0 on success, else the CUresult.
inline int make_tensor_map_bhnd(CUtensorMap* tm, const void* base, bool is_bf16, int64_t B,
                                int64_t H, int64_t N, int64_t D, int64_t sb, int64_t sh, int64_t sn,
                                int box_rows, int box_cols = 64, int swizzle_bytes = 128) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  // strides in bytes for dims 1..3
  int64_t sn_b = sn * 2, sh_b = sh * 2, sb_b = sb * 2;
  if (N == 1 || sn_b == 0) sn_b = D * 2;
  if (H == 1 || sh_b == 0) sh_b = sn_b * N;
  if (B == 1 || sb_b == 0) sb_b = sh_b * H;
  cuuint64_t strides[3] = {(cuuint64_t)sn_b, (cuuint64_t)sh_b, (cuuint64_t)sb_b};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                        : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

// Launch with programmatic dependent launch enabled (see pdl_wait / pdl_launch_dependents).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace fcsa
