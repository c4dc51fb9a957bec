// TMEM -> register bandwidth by tcgen05.ld shape, warps per SM and loads in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_ldtm probe_ldtm.cu
// The element-wise stages of both attention kernels read every S (and dP, dQ) value out of TMEM
// exactly once; this measures the ceiling of that path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../flash_cosine_sim_attention_b200/csrc/sm100_primitives.cuh"

using namespace fcsa;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

#define LD32(SHAPE, taddr, r)                                                                              \
  asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 "                                                    \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                   \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"   \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),       \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),   \
                 "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),             \
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),             \
                 "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])              \
               : "r"(taddr))

// SHAPE: 0 = 32x32b.x32 (32 lanes x 32 columns), 1 = 16x256b.x8 (16 lanes x 64 columns),
//        2 = 16x128b.x16 (16 lanes x 64 columns), 3 = 16x64b.x32 (16 lanes x 64 columns)
// every instruction moves 4 KB; INFLIGHT of them are issued before one wait
template <int SHAPE, int INFLIGHT>
__global__ void __launch_bounds__(512) probe(uint32_t* out, int iters, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[INFLIGHT][32];
#pragma unroll
    for (int c = 0; c < INFLIGHT; ++c) {
      const uint32_t col = ((c * 64 + (warp >> 2) * 128) & 511);
      if (SHAPE == 0) LD32("32x32b.x32", lane_base + (col & 480) + ((c & 1) * 32 & 0), r[c]);
      else if (SHAPE == 1) LD32("16x256b.x8", lane_base + ((uint32_t)((c & 1) * 16) << 16) + (col & 448), r[c]);
      else if (SHAPE == 2) LD32("16x128b.x16", lane_base + ((uint32_t)((c & 1) * 16) << 16) + (col & 448), r[c]);
      else LD32("16x64b.x32", lane_base + ((uint32_t)((c & 1) * 16) << 16) + (col & 448), r[c]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < INFLIGHT; ++c)
#pragma unroll
      for (int i = 0; i < 32; i += 4) acc += (r[c][i] ^ r[c][i + 1]) + (r[c][i + 2] ^ r[c][i + 3]);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int SHAPE, int INFLIGHT>
void run(const char* name, uint32_t* dOut, long long* dCyc, int sms) {
  for (int nthr = 128; nthr <= 512; nthr *= 2) {
    const int iters = 1024;
    probe<SHAPE, INFLIGHT><<<sms, nthr>>>(dOut, iters, dCyc);
    CK(cudaDeviceSynchronize());
    probe<SHAPE, INFLIGHT><<<sms, nthr>>>(dOut, iters, dCyc);
    CK(cudaDeviceSynchronize());
    long long cyc;
    CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
    const double bytes = (double)iters * INFLIGHT * 4096.0 * (nthr / 32);
    printf("[LDTM] %-12s in flight %d, %2d warps/SM: %6.1f B/clk/SM  (%.0f cycles per 4 KB load per warp)\n", name,
           INFLIGHT, nthr / 32, bytes / cyc, (double)cyc / (iters * INFLIGHT));
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  uint32_t* dOut;
  long long* dCyc;
  CK(cudaMalloc(&dOut, prop.multiProcessorCount * 512 * 4));
  CK(cudaMalloc(&dCyc, 8));
  const int sms = prop.multiProcessorCount;
  run<0, 1>("32x32b.x32", dOut, dCyc, sms);
  run<0, 2>("32x32b.x32", dOut, dCyc, sms);
  run<0, 4>("32x32b.x32", dOut, dCyc, sms);
  run<1, 2>("16x256b.x8", dOut, dCyc, sms);
  run<1, 4>("16x256b.x8", dOut, dCyc, sms);
  run<2, 2>("16x128b.x16", dOut, dCyc, sms);
  run<2, 4>("16x128b.x16", dOut, dCyc, sms);
  run<3, 2>("16x64b.x32", dOut, dCyc, sms);
  run<3, 4>("16x64b.x32", dOut, dCyc, sms);
  printf("probe_ldtm done\n");
  return 0;
}
