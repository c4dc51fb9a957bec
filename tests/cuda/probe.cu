// Hardware probe for the sm_100a primitives in csrc/sm100_primitives.cuh.
//
// Part 1 (correctness): small GEMMs through every operand form the attention kernels use,
//   with integer-valued inputs so the fp32 result must match the CPU EXACTLY:
//     T1a  SS   A K-major, B K-major, K=64          (S = Q K^T, D=64)
//     T1b  SS   same, K=128 as two 64-wide chunks   (S = Q K^T, D=128)
//     T2a  TS   A in TMEM (packed 16-bit), B MN-major N=64    (O += P V, D=64)
//     T2b  TS   B MN-major N=128 via LBO                      (O += P V, D=128)
//     T3a  SS   A MN-major written by threads with the 128B swizzle, B MN-major   (dQ = dS K)
//     T3b  SS   the same thread-written buffer read as a K-major A (K=128, two chunks)
// Part 2 (throughput): cycles per tcgen05.mma for each form, MUFU.EX2 and FMA-pipe exp2 rates.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe probe.cu
// Test infrastructure only - not part of the product library.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../flash_cosine_sim_attention_b200/csrc/sm100_primitives.cuh"
#include "../../flash_cosine_sim_attention_b200/csrc/tensor_map.h"

using namespace fcsa;
typedef __nv_bfloat16 bf16;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__host__ __device__ inline float tval(int i, int j, int seed) {
  return (float)(((i * 131 + j * 71 + seed * 29) % 23) - 11) / 8.0f;
}

enum Mode { T1A = 0, T1B, T2A, T2B, T3A, T3B };

// One CTA, 128 threads.  Global operands: gA [128 x 128] bf16 row-major (cols = 128),
// gB [128 x 128] bf16 row-major, both behind 4-D tensor maps with a 64 x 128 box.
// out: [128 x 128] fp32 (only the first N columns written).
__global__ void __launch_bounds__(128, 1)
probe_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           float* out, int mode, int nrep, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;               // 32 KB (two 16 KB chunks)
  uint8_t* sB = smem + 32768;       // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_ld = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- stage operands -------------------------------------------------------------
  if (tid == 0) {
    uint32_t bytes = 0;
    bool loadA = (mode == T1A || mode == T1B);
    if (loadA) {
      tma_load_4d(smem_u32(sA), &tmA, bar_ld, 0, 0, 0, 0);
      bytes += 16384;
      if (mode == T1B) {
        tma_load_4d(smem_u32(sA + 16384), &tmA, bar_ld, 64, 0, 0, 0);
        bytes += 16384;
      }
    }
    tma_load_4d(smem_u32(sB), &tmB, bar_ld, 0, 0, 0, 0);
    bytes += 16384;
    if (mode == T1B || mode == T2B) {
      tma_load_4d(smem_u32(sB + 16384), &tmB, bar_ld, 64, 0, 0, 0);
      bytes += 16384;
    }
    mbar_expect_tx(bar_ld, bytes);
  }

  if (mode == T2A || mode == T2B) {
    // A = P in TMEM: thread = row m, 128 K-values packed 2 per 32-bit column -> 64 columns at col 256
    const int m = tid;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        int k = c * 32 + 2 * i;
        r[i] = pack2<bf16>(tval(m, k, 1), tval(m, k + 1, 1));
      }
      tmem_st_x16(lane_base + 256 + c * 16, r);
    }
    tmem_st_wait();
  }
  if (mode == T3A || mode == T3B) {
    // thread kk writes row kk of the [128 rows][128 contiguous] buffer X[kk][m] = tval(m, kk, 1)
    // as two SWIZZLE_128B chunks of [128 rows x 64 elements].
    const int kk = tid;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int mm = ch * 64 + c * 8 + 2 * i;
          w[i] = pack2<bf16>(tval(mm, kk, 1), tval(mm + 1, kk, 1));
        }
        uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(sA + ch * 16384 + sw128_offset(kk, c)) = v;
      }
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- MMA ------------------------------------------------------------------------
  if (warp == 0 && elect_one()) {
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    long long t0 = clock64();
    for (int rep = 0; rep < nrep; ++rep) {
      if (mode == T1A || mode == T1B) {
        const uint32_t idesc = umma_idesc<bf16>(128, 128, 0, 0);
        const int ksteps = (mode == T1A) ? 4 : 8;
        for (int k = 0; k < ksteps; ++k) {
          uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          uint64_t da = umma_desc_sw128(smem_u32(sA) + off, 16, 1024);
          uint64_t db = umma_desc_sw128(smem_u32(sB) + off, 16, 1024);
          umma_ss(tmem, da, db, idesc, (k > 0) ? 1u : 0u);
        }
      } else if (mode == T2A || mode == T2B) {
        const uint32_t N = (mode == T2A) ? 64 : 128;
        const uint32_t idesc = umma_idesc<bf16>(128, N, 0, 1);
        for (int k = 0; k < 8; ++k) {
          uint64_t db = umma_desc_sw128(smem_u32(sB) + k * 2048, 16384, 1024);
          umma_ts(tmem, tmem + 256 + k * 8, db, idesc, (k > 0) ? 1u : 0u);
        }
      } else if (mode == T3A) {
        // D[128 x 64] = X^T (A MN-major: M contiguous) * B (MN-major)
        const uint32_t idesc = umma_idesc<bf16>(128, 64, 1, 1);
        for (int k = 0; k < 8; ++k) {
          uint64_t da = umma_desc_sw128(smem_u32(sA) + k * 2048, 16384, 1024);
          uint64_t db = umma_desc_sw128(smem_u32(sB) + k * 2048, 16384, 1024);
          umma_ss(tmem, da, db, idesc, (k > 0) ? 1u : 0u);
        }
      } else {  // T3B: D[128 x 64] = X (A K-major, K = 128 in two chunks) * B (MN-major)
        const uint32_t idesc = umma_idesc<bf16>(128, 64, 0, 1);
        for (int k = 0; k < 8; ++k) {
          uint32_t offa = (k >> 2) * 16384 + (k & 3) * 32;
          uint64_t da = umma_desc_sw128(smem_u32(sA) + offa, 16, 1024);
          uint64_t db = umma_desc_sw128(smem_u32(sB) + k * 2048, 16384, 1024);
          umma_ss(tmem, da, db, idesc, (k > 0) ? 1u : 0u);
        }
      }
    }
    umma_commit(bar_mma);
    mbar_wait(bar_mma, 0);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // ---- read back ------------------------------------------------------------------
  {
    const int m = tid;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_x32(lane_base + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) out[m * 128 + c * 32 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// exp2 throughput
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float exp2_poly(float x) {
  // 2^x for x <= 0: round-to-nearest split + degree-4 minimax on [-0.5, 0.5]
  x = fmaxf(x, -125.0f);
  float xf = x + 12582912.0f;
  float n = xf - 12582912.0f;
  float r = x - n;
  float p = 9.6181291e-3f;
  p = fmaf(p, r, 5.5504109e-2f);
  p = fmaf(p, r, 2.4022651e-1f);
  p = fmaf(p, r, 6.9314718e-1f);
  p = fmaf(p, r, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}

__global__ void __launch_bounds__(512) probe_exp(float* out, int iters, int mode, long long* cycles) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + 1) - 0.01f * i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      if (mode == 0) y = ex2_approx(a[i]);
      else if (mode == 1) y = exp2_poly(a[i]);
      else y = (i & 1) ? exp2_poly(a[i]) : ex2_approx(a[i]);
      a[i] = y - 1.0f;  // keeps the argument in (-1, 0] and serialises the chain
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// ------------------------------------------------------------------------------------
// the forward softmax inner loop in isolation: per element FFMA + EX2 + FADD, per pair one F2FP
//   mode 0: scalar math, registers only      mode 1: packed f32x2 math, registers only
//   mode 2: mode 1 + tcgen05.ld of the 128 inputs and tcgen05.st of the 64 packed outputs per tile
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) probe_softmax_loop(float* out, int iters, int mode, long long* cycles) {
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
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 192;
  const float c1 = 11.5f, nc2 = -11.5f;
  float l = 0.f;
  float2 l2 = make_float2(0.f, 0.f);
  uint32_t sreg[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) sreg[c][i] = __float_as_uint(0.001f * ((threadIdx.x * 7 + c * 32 + i) % 97));
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(lane_base + 32 * c, sreg[c]);
      tmem_ld_wait();
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float p0, p1;
        if (mode == 0) {
          p0 = ex2_approx(fmaf(__uint_as_float(sreg[c][2 * i]), c1, nc2));
          p1 = ex2_approx(fmaf(__uint_as_float(sreg[c][2 * i + 1]), c1, nc2));
          l += p0 + p1;
        } else {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sreg[c][2 * i]), __uint_as_float(sreg[c][2 * i + 1])),
                                      make_float2(c1, c1), make_float2(nc2, nc2));
          p0 = ex2_approx(x.x);
          p1 = ex2_approx(x.y);
          l2 = __fadd2_rn(l2, make_float2(p0, p1));
        }
        if (mode == 3) pk[i] = __float_as_uint(p0) ^ __float_as_uint(p1);          // no F2FP
        else if (mode == 4) pk[i] = __float_as_uint(p0 + p1);                        // FADD instead of F2FP
        else if (mode == 5) pk[i] = pack2<__half>(p0, p1);                             // f16 pack
        else pk[i] = pack2<bf16>(p0, p1);
      }
      if (mode == 2) {
        tmem_st_x16(lane_base + 128 + 16 * c, pk);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) sreg[c][i] ^= (pk[i] & 1u);    // keep the packs alive, perturb the inputs
      }
    }
    if (mode == 2) tmem_st_wait();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + l2.x + l2.y;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// same math, but the body is ONE 32-element chunk executed 4x per tile from a rolled loop
// (code footprint ~1/4): tests whether instruction fetch limits the fully unrolled version
template <int UNROLL_CHUNKS>
__global__ void __launch_bounds__(256) probe_softmax_rolled(float* out, int iters, long long* cycles) {
  const float c1 = 11.5f, nc2 = -11.5f;
  float2 l2 = make_float2(0.f, 0.f);
  uint32_t sreg[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) sreg[i] = __float_as_uint(0.001f * ((threadIdx.x * 7 + i) % 97));
  long long t0 = clock64();
#pragma unroll UNROLL_CHUNKS
  for (int it = 0; it < iters * 4; ++it) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 x = __ffma2_rn(make_float2(__uint_as_float(sreg[2 * i]), __uint_as_float(sreg[2 * i + 1])),
                                  make_float2(c1, c1), make_float2(nc2, nc2));
      const float p0 = ex2_approx(x.x), p1 = ex2_approx(x.y);
      l2 = __fadd2_rn(l2, make_float2(p0, p1));
      pk[i] = pack2<bf16>(p0, p1);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      sreg[2 * i] ^= (pk[i] & 1u);
      sreg[2 * i + 1] ^= ((pk[i] >> 16) & 1u);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l2.x + l2.y;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// ------------------------------------------------------------------------------------
// TMEM -> register bandwidth (tcgen05.ld 32x32b.x32), no tensor-pipe activity
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) probe_ldtm(float* out, int iters, int with_exp, long long* cycles) {
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
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_x32(lane_base + ((c * 32 + (warp >> 2) * 128) & 511), r);
      tmem_ld_wait();
      if (with_exp) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += ex2_approx(__uint_as_float(r[i]) * 1e-30f);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i] & 1u);
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// tcgen05.ld latency while the tensor pipe streams MMAs (accumulator traffic in TMEM)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1) probe_ldtm_mma(float* out, int iters, int mma_mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) stop_flag = 0;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < 4) {
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(lane_base + c * 32, r);      // columns [0,128): not written by the MMAs below
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i] & 1u);
      }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) { *cycles = t1 - t0; stop_flag = 1; }
  } else if (warp == 4 && mma_mode) {
    if (elect_one()) {
      const uint32_t sA = smem_u32(smem), sB = sA + 32768;
      // mode 1: SS M128 N128 into columns [256,384); mode 2: TS M128 N64 (A from tmem [384..]) into [256,320)
      const uint32_t idesc = mma_mode == 1 ? umma_idesc<bf16>(128, 128, 0, 0) : umma_idesc<bf16>(128, 64, 0, 1);
      int n = 0;
      while (!stop_flag && n < (1 << 20)) {
        for (int k = 0; k < 8; ++k) {
          if (mma_mode == 1)
            umma_ss(tmem + 256, umma_desc_sw128(sA + (k & 3) * 32, 16, 1024), umma_desc_sw128(sB + (k & 3) * 32, 16, 1024), idesc, 1u);
          else
            umma_ts(tmem + 256, tmem + 384 + k * 8, umma_desc_sw128(sB + k * 2048, 16384, 1024), idesc, 1u);
        }
        n += 8;
        if ((n & 63) == 0) {   // bound the queue depth: wait for completion every 64 MMAs
          __shared__ uint64_t bar;
          if (n == 64) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
          umma_commit(smem_u32(&bar));
          mbar_wait(smem_u32(&bar), ((n >> 6) - 1) & 1);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  const int R = 128, C = 128;
  std::vector<bf16> hA(R * C), hB(R * C);
  bf16 *dA, *dB;
  float* dOut;
  long long* dCyc;
  CK(cudaMalloc(&dA, R * C * 2));
  CK(cudaMalloc(&dB, R * C * 2));
  CK(cudaMalloc(&dOut, R * C * 4));
  CK(cudaMalloc(&dCyc, 8));
  CK(cudaFuncSetAttribute(probe_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000));

  int fails = 0;
  const char* names[6] = {"T1a SS K-major K=64", "T1b SS K-major K=128 (2 chunks)",
                          "T2a TS A=TMEM, B MN-major N=64", "T2b TS A=TMEM, B MN-major N=128 (LBO)",
                          "T3a SS A MN-major (thread-written), B MN-major",
                          "T3b SS A K-major K=128 (thread-written), B MN-major"};
  for (int mode = 0; mode < 6; ++mode) {
    // global operands: gA[i][j] = tval(i, j, 1) (row-major), gB[i][j] = tval(i, j, 2)
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) {
        hA[i * C + j] = __float2bfloat16(tval(i, j, 1));
        hB[i * C + j] = __float2bfloat16(tval(i, j, 2));
      }
    CK(cudaMemcpy(dA, hA.data(), R * C * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), R * C * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0, R * C * 4));
    CUtensorMap tmA, tmB;
    int r1 = make_tensor_map_bhnd(&tmA, dA, true, 1, 1, R, C, R * C, R * C, C, 128);
    int r2 = make_tensor_map_bhnd(&tmB, dB, true, 1, 1, R, C, R * C, R * C, C, 128);
    if (r1 || r2) {
      printf("tensor map encode failed %d %d\n", r1, r2);
      return 2;
    }
    probe_gemm<<<1, 128, 70000>>>(tmA, tmB, dOut, mode, 1, dCyc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hOut(R * C);
    CK(cudaMemcpy(hOut.data(), dOut, R * C * 4, cudaMemcpyDeviceToHost));

    // CPU reference
    int N = (mode == T2A || mode == T3A || mode == T3B) ? 64 : 128;
    int bad = 0;
    double maxerr = 0;
    int first_bad_m = -1, first_bad_n = -1;
    float first_got = 0, first_exp = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        if (mode == T1A || mode == T1B) {
          int K = (mode == T1A) ? 64 : 128;
          for (int k = 0; k < K; ++k) acc += (double)tval(m, k, 1) * tval(n, k, 2);  // A[m][k] B[n][k]
        } else if (mode == T2A || mode == T2B) {
          for (int k = 0; k < 128; ++k) acc += (double)tval(m, k, 1) * tval(k, n, 2);  // P[m][k] V[k][n]
        } else if (mode == T3A) {
          // A[m][k] = X[k][m] = tval(m, k, 1)
          for (int k = 0; k < 128; ++k) acc += (double)tval(m, k, 1) * tval(k, n, 2);
        } else {
          // A[m][k] = X[m][k] = tval(k, m, 1)
          for (int k = 0; k < 128; ++k) acc += (double)tval(k, m, 1) * tval(k, n, 2);
        }
        double e = fabs(acc - hOut[m * 128 + n]);
        if (e > maxerr) maxerr = e;
        if (e > 1e-3) {
          if (!bad) {
            first_bad_m = m;
            first_bad_n = n;
            first_got = hOut[m * 128 + n];
            first_exp = (float)acc;
          }
          ++bad;
        }
      }
    printf("[%s] %s  maxerr=%.4g bad=%d", bad ? "FAIL" : "PASS", names[mode], maxerr, bad);
    if (bad) printf(" first bad (m=%d,n=%d) got %.4f exp %.4f", first_bad_m, first_bad_n, first_got, first_exp);
    printf("\n");
    if (bad) {
      ++fails;
      // dump a corner to help diagnose layout errors
      for (int m = 0; m < 4; ++m) {
        printf("   row %d got:", m);
        for (int n = 0; n < 8; ++n) printf(" %8.3f", hOut[m * 128 + n]);
        printf("\n");
      }
    }
  }

  // ---- throughput: cycles per MMA (single CTA on one SM) ----------------------------
  {
    CUtensorMap tmA, tmB;
    make_tensor_map_bhnd(&tmA, dA, true, 1, 1, R, C, R * C, R * C, C, 128);
    make_tensor_map_bhnd(&tmB, dB, true, 1, 1, R, C, R * C, R * C, C, 128);
    const int nrep = 64;
    int mmas_per_rep[6] = {4, 8, 8, 8, 8, 8};
    const char* shp[6] = {"SS  M128 N128 K16 (A,B K-major)", "SS  M128 N128 K16 (K=128)",
                          "TS  M128 N64  K16 (B MN-major)", "TS  M128 N128 K16 (B MN-major)",
                          "SS  M128 N64  K16 (A MN, B MN)", "SS  M128 N64  K16 (A K, B MN)"};
    for (int mode = 0; mode < 6; ++mode) {
      probe_gemm<<<1, 128, 70000>>>(tmA, tmB, dOut, mode, nrep, dCyc);
      CK(cudaDeviceSynchronize());
      probe_gemm<<<1, 128, 70000>>>(tmA, tmB, dOut, mode, nrep, dCyc);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
      printf("[RATE] %s : %.1f cycles / MMA (%lld cycles for %d MMAs)\n", shp[mode],
             (double)cyc / (nrep * mmas_per_rep[mode]), cyc, nrep * mmas_per_rep[mode]);
    }
  }
  // ---- exp2 throughput --------------------------------------------------------------
  {
    float* dE;
    CK(cudaMalloc(&dE, 148 * 512 * 4));
    const char* en[3] = {"MUFU ex2.approx", "FMA-pipe polynomial exp2", "1:1 MUFU + polynomial"};
    for (int nthr = 128; nthr <= 512; nthr *= 2)
    for (int mode = 0; mode < 3; ++mode) {
      const int iters = 2048;
      probe_exp<<<prop.multiProcessorCount, nthr>>>(dE, iters, mode, dCyc);
      CK(cudaDeviceSynchronize());
      probe_exp<<<prop.multiProcessorCount, nthr>>>(dE, iters, mode, dCyc);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
      double per_clk = (double)iters * 8 * nthr / (double)cyc;
      printf("[RATE] %s : %.2f exp2 / clk / SM (%d warps/SM, dependent chains x8)\n", en[mode], per_clk, nthr / 32);
    }
  }
  {
    float* dE;
    CK(cudaMalloc(&dE, 148 * 256 * 4));
    const char* mm[6] = {"scalar FFMA/EX2/FADD + F2FP, registers only", "packed f32x2 math, registers only",
                         "packed math + tcgen05.ld x128 / st x64 per tile", "packed math, XOR instead of F2FP",
                         "packed math, FADD instead of F2FP", "packed math, f16 F2FP"};
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 256;
      probe_softmax_loop<<<prop.multiProcessorCount, 256>>>(dE, iters, mode, dCyc);
      CK(cudaDeviceSynchronize());
      probe_softmax_loop<<<prop.multiProcessorCount, 256>>>(dE, iters, mode, dCyc);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
      printf("[RATE] softmax loop, 8 warps/SM, %s: %.0f cycles per 128-element tile row-set (MUFU bound 2048 per 2 warps/SMSP), %.2f exp/clk/SM\n",
             mm[mode], (double)cyc / iters, 256.0 * 128 * iters / cyc);
    }
  }
  {
    float* dE;
    CK(cudaMalloc(&dE, 148 * 256 * 4));
    const int iters = 256;
    for (int v = 0; v < 3; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        if (v == 0) probe_softmax_rolled<1><<<prop.multiProcessorCount, 256>>>(dE, iters, dCyc);
        if (v == 1) probe_softmax_rolled<4><<<prop.multiProcessorCount, 256>>>(dE, iters, dCyc);
        if (v == 2) probe_softmax_rolled<16><<<prop.multiProcessorCount, 256>>>(dE, iters, dCyc);
        CK(cudaDeviceSynchronize());
      }
      long long cyc;
      CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
      printf("[RATE] softmax loop rolled (unroll %d chunks of 32), 8 warps/SM: %.0f cycles per 128 elements, %.2f exp/clk/SM\n",
             v == 0 ? 1 : (v == 1 ? 4 : 16), (double)cyc / iters, 256.0 * 128 * iters / cyc);
    }
  }
  // ---- TMEM read bandwidth ------------------------------------------------------------
  {
    float* dE;
    CK(cudaMalloc(&dE, 148 * 512 * 4));
    for (int with_exp = 0; with_exp < 2; ++with_exp)
      for (int nthr = 128; nthr <= 512; nthr *= 2) {
        const int iters = 512;
        probe_ldtm<<<prop.multiProcessorCount, nthr>>>(dE, iters, with_exp, dCyc);
        CK(cudaDeviceSynchronize());
        probe_ldtm<<<prop.multiProcessorCount, nthr>>>(dE, iters, with_exp, dCyc);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
        double bytes = (double)iters * 4 * 32 * 4 * nthr;
        printf("[RATE] tcgen05.ld 32x32b.x32 %s: %.1f B / clk / SM (%d warps/SM; %.0f cycles per x32 per warp)\n",
               with_exp ? "+32 ex2 each" : "(ld only)   ", bytes / cyc, nthr / 32, (double)cyc / (iters * 4));
      }
  }
  {
    float* dE;
    CK(cudaMalloc(&dE, 512 * 4));
    CK(cudaFuncSetAttribute(probe_ldtm_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000));
    const char* mm[3] = {"tensor pipe idle", "while SS M128N128 MMAs stream", "while TS M128N64 MMAs stream"};
    for (int mode = 0; mode < 3; ++mode) {
      const int iters = 256;
      probe_ldtm_mma<<<1, 192, 70000>>>(dE, iters, mode, dCyc);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
      printf("[RATE] tcgen05.ld x32 + wait, 4 warps, %s: %.0f cycles per load\n", mm[mode], (double)cyc / (iters * 4));
    }
  }
  printf("probe done, %d correctness failures\n", fails);
  return fails ? 1 : 0;
}
