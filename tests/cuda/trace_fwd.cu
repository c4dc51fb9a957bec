// In-kernel timeline of one CTA of the forward kernel at the benchmark shape (4,8,4096,64) causal.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFCSA_TRACE -o trace_fwd trace_fwd.cu
// Test infrastructure only.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

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
  const int B = argc > 1 ? atoi(argv[1]) : 4, H = argc > 2 ? atoi(argv[2]) : 8;
#ifndef FCSA_HARNESS_D
#define FCSA_HARNESS_D 64
#endif
  constexpr int D = FCSA_HARNESS_D;
  const int Nq = argc > 3 ? atoi(argv[3]) : 4096, Nk = argc > 4 ? atoi(argv[4]) : 4096;
  const int causal = argc > 5 ? atoi(argv[5]) : 1;
  const int grid_arg = argc > 6 ? atoi(argv[6]) : 148;    // CTAs (persistent); 0 = one CTA per work item
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
    const int items = a.n_qblk * B * H;
    kern<<<(grid_arg > 0 && grid_arg < items) ? grid_arg : items, Cfg::kThreads, Cfg::kSmem>>>(tq, tk, tv, a);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("B %d H %d Nq %d Nk %d causal %d grid %d: fwd kernel best of %d: %.1f us\n", B, H, Nq, Nk, causal, grid_arg, reps, best * 1e3);
  if (argc > 7) {
    // sustained: N launches back to back (no host synchronisation in between), average time per launch
    const int nrun = atoi(argv[7]);
    const int items = a.n_qblk * B * H;
    const int grid = (grid_arg > 0 && grid_arg < items) ? grid_arg : items;
    for (int pass = 0; pass < 3; ++pass) {
      CK(cudaEventRecord(e0));
      for (int i = 0; i < nrun; ++i) kern<<<grid, Cfg::kThreads, Cfg::kSmem>>>(tq, tk, tv, a);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("  sustained pass %d: %d launches back to back: %.1f us per launch\n", pass, nrun, ms * 1e3 / nrun);
    }
  }
#ifdef FCSA_CTA_TIMELINE
  {
    // per SM: the CTAs it ran, in start order; columns relative to the SM's first CTA start
    const int nblk = a.n_qblk * B * H;
    static long long ct[4096][10];
    CK(cudaMemcpyFromSymbol(ct, g_fcsa_cta_t, sizeof(ct)));
    printf("slots: 0 entry, 1 setup done, 2/3 tile0/1 second iteration top, 4/5 tile0/1 loop end, 6 tile1 epilogue done, 7 exit\n");
    for (int sm = 0; sm < 3; ++sm) {
      std::vector<int> ids;
      for (int i = 0; i < nblk && i < 4096; ++i) if (ct[i][8] == sm) ids.push_back(i);
      std::sort(ids.begin(), ids.end(), [&](int x, int y) { return ct[x][0] < ct[y][0]; });
      if (ids.empty()) continue;
      const long long t0 = ct[ids[0]][0];
      long long prev_exit = t0;
      for (int id : ids) {
        printf("sm %d cta %4d (qblk rank %3d): gap %6lld |", sm, id, id / (B * H), ct[id][0] - prev_exit);
        for (int s2 = 0; s2 < 8; ++s2) printf(" %7lld", ct[id][s2] - ct[id][0]);
        printf("\n");
        prev_exit = ct[id][7];
      }
    }
    // totals over all SMs
    double gap = 0, setup = 0, body = 0, epi = 0, exitt = 0; int n = 0, ngap = 0;
    for (int sm = 0; sm < 148; ++sm) {
      std::vector<int> ids;
      for (int i = 0; i < nblk && i < 4096; ++i) if (ct[i][8] == sm) ids.push_back(i);
      std::sort(ids.begin(), ids.end(), [&](int x, int y) { return ct[x][0] < ct[y][0]; });
      for (size_t k2 = 0; k2 < ids.size(); ++k2) {
        const int id = ids[k2];
        if (k2 > 0) { gap += ct[id][0] - ct[ids[k2 - 1]][7]; ++ngap; }
        setup += ct[id][1] - ct[id][0];
        epi += ct[id][6] - ct[id][5];
        exitt += ct[id][7] - ct[id][6];
        body += ct[id][5] - ct[id][1];
        ++n;
      }
    }
    printf("mean per CTA (cycles): launch gap %.0f, setup %.0f, body(setup..tile1 loop end) %.0f, tile1 epilogue %.0f, exit %.0f  (%d CTAs)\n",
           gap / (ngap ? ngap : 1), setup / n, body / n, epi / n, exitt / n, n);
  }
#endif
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
