// In-kernel timeline of one CTA of the forward kernel at the benchmark shape (4,8,4096,64) causal.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFCSA_TRACE -o trace_fwd trace_fwd.cu
// Test infrastructure only.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../flash_cosine_sim_attention_b200/csrc/fwd_kernel.cuh"
#include "../../flash_cosine_sim_attention_b200/csrc/tensor_map.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__global__ void fill(__nv_bfloat16* p, size_t n, unsigned seed, float amp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 2654435761u + seed;
  x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  p[i] = __float2bfloat16(((x & 0xFFFF) / 65536.0f - 0.5f) * amp);
}

int main(int argc, char** argv) {
  // usage: [B H Nq Nk causal]  (default: the benchmark shape).  Without -DFCSA_TRACE this is a plain
  // timing harness (A/B experiments): ... -o time_fwd trace_fwd.cu
  const int B = argc > 1 ? atoi(argv[1]) : 4, H = argc > 2 ? atoi(argv[2]) : 8, D = 64;
  const int Nq = argc > 3 ? atoi(argv[3]) : 4096, Nk = argc > 4 ? atoi(argv[4]) : 4096;
  const int causal = argc > 5 ? atoi(argv[5]) : 1;
  const int N = Nq > Nk ? Nq : Nk;
  const size_t n = (size_t)B * H * N * D;
  __nv_bfloat16 *q, *k, *v, *o;
  float* inv_l;
  for (auto pp : {&q, &k, &v, &o}) CK(cudaMalloc(pp, n * 2));
  CK(cudaMalloc(&inv_l, (size_t)B * H * N * 4));
  fill<<<(n + 255) / 256, 256>>>(q, n, 1, 0.25f);
  fill<<<(n + 255) / 256, 256>>>(k, n, 2, 0.25f);
  fill<<<(n + 255) / 256, 256>>>(v, n, 3, 2.f);
  CUtensorMap tq, tk, tv;
  long long sb = (long long)H * N * D, sh = (long long)N * D, sn = D;
  if (fcsa::make_tensor_map_bhnd(&tq, q, true, B, H, Nq, D, sb, sh, sn, 128) ||
      fcsa::make_tensor_map_bhnd(&tk, k, true, B, H, Nk, D, sb, sh, sn, 128) ||
      fcsa::make_tensor_map_bhnd(&tv, v, true, B, H, Nk, D, sb, sh, sn, 128)) { printf("tmap fail\n"); return 1; }
  fcsa::FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.causal = causal; a.has_mask = 0; a.kv_heads = H; a.n_qblk = (Nq + 255) / 256;
  a.c1 = 8.f * 1.44269504f; a.c2 = a.c1; a.mask = nullptr; a.mask_sb = 0;
  a.o = o; a.o_sb = sb; a.o_sh = sh; a.o_sn = sn; a.inv_l = inv_l;
  using Cfg = fcsa::FwdCfg<D>;
  auto kern = fcsa::fcsa_fwd_kernel<__nv_bfloat16, D>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
#ifdef FCSA_TRACE
  const int reps = 3;
#else
  const int reps = 12;
#endif
  float best = 1e9f;
  for (int rep = 0; rep < reps; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<a.n_qblk * B * H, Cfg::kThreads, Cfg::kSmem>>>(tq, tk, tv, a);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("B %d H %d Nq %d Nk %d causal %d: fwd kernel best of %d: %.1f us\n", B, H, Nq, Nk, causal, reps, best * 1e3);
#ifdef FCSA_TRACE
  static long long tr[8][48][8];
  CK(cudaMemcpyFromSymbol(tr, g_fcsa_trace, sizeof(tr)));
  long long t0 = tr[1][0][0];
  const char* names[3] = {"MMA  [S_FREE0 seen, S_FREE1 seen, P_FULL0 seen, V ok, P_FULL1 seen, V ok]",
                          "SM0  [top, S_FULL seen, S in regs, P_FREE ok, exps done, P arrived]", "SM1  [same]"};
  for (int role = 0; role < 3; ++role) {
    printf("--- %s\n", names[role]);
    for (int i = 0; i < 14; ++i) {
      printf("  it %2d:", i);
      for (int s = 0; s < 6; ++s) printf(" %8lld", tr[role][i][s] ? tr[role][i][s] - t0 : -1);
      printf("\n");
    }
  }
  printf("SM0 period:");
  for (int i = 1; i < 31; ++i) printf(" %lld", tr[1][i][1] - tr[1][i - 1][1]);
  printf("\n");
#endif
  return 0;
}
