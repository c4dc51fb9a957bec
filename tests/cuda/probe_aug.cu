// Correctness probe for the K = 16 "sliver" operands (32-byte rows, TMA SWIZZLE_32B, UMMA layout 6):
//   D[128 x 128] = A[128 x 16] * B[128 x 16]^T   with both tiles loaded by TMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_aug probe_aug.cu -lcuda
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../flash_cosine_sim_attention_b200/csrc/sm100_primitives.cuh"
#include "../../flash_cosine_sim_attention_b200/csrc/tensor_map.h"

using namespace fcsa;
typedef __nv_bfloat16 bf16;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 4 KB
  uint8_t* sB = smem + 4096;     // 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8192 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_ld = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 128); tmem_relinquish(); }
  if (tid == 0) { mbar_init(bar_ld, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(bar_ld, 8192);
    tma_load_4d(smem_u32(sA), &tmA, bar_ld, 0, 0, 0, 0);
    tma_load_4d(smem_u32(sB), &tmB, bar_ld, 16, 0, 0, 0);     // the second 16-column block of a 32-wide tensor
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc<bf16>(128, 128, 0, 0);
    umma_ss(tmem, umma_desc_sw32(smem_u32(sA)), umma_desc_sw32(smem_u32(sB)), idesc, 0u);
    umma_commit(bar_mma);
  }
  __syncthreads();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld_x32(lane_base + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[tid * 128 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  // A: [128 rows x 16], B: columns 16..31 of a [128 rows x 32] tensor
  std::vector<bf16> hA(128 * 16), hB(128 * 32);
  std::vector<float> fA(128 * 16), fB(128 * 32);
  unsigned x = 12345;
  auto rnd = [&]() { x = x * 1664525u + 1013904223u; return ((x >> 9) & 0xFFFF) / 65536.0f - 0.5f; };
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
  bf16 *dA, *dB; float* dOut;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, 128 * 128 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  int r1 = make_tensor_map_bhnd(&tmA, dA, true, 1, 1, 128, 16, 128 * 16, 128 * 16, 16, 128, 16, 32);
  int r2 = make_tensor_map_bhnd(&tmB, dB, true, 1, 1, 128, 32, 128 * 32, 128 * 32, 32, 128, 16, 32);
  if (r1 || r2) { printf("tensor map failed %d %d\n", r1, r2); return 2; }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  probe<<<1, 128, 16384>>>(tmA, tmB, dOut);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * 128);
  CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < 16; ++k) ref += (double)fA[m * 16 + k] * fB[n * 32 + 16 + k];
      double e = fabs(ref - out[m * 128 + n]);
      if (e > maxerr) maxerr = e;
      if (e > 1e-3) ++bad;
    }
  printf("[%s] SS K=16 sliver, TMA SWIZZLE_32B + UMMA layout 6: maxerr=%.4g bad=%d\n", bad ? "FAIL" : "PASS", maxerr, bad);
  return bad ? 1 : 0;
}
