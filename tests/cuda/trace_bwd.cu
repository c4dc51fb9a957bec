// In-kernel timeline of one CTA of the backward kernel at the benchmark shape (4,8,4096,64) causal.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFCSA_TRACE -o trace_bwd trace_bwd.cu
// Without -DFCSA_TRACE the same file is a plain timing harness of the main kernel (used for A/B
// experiments with -DFCSA_EXP_* switches): ... -o time_bwd trace_bwd.cu
// Test infrastructure only.
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "../../flash_cosine_sim_attention_b200/csrc/bwd_kernel.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__global__ void fill(__nv_bfloat16* p, size_t n, unsigned seed, float amp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 2654435761u + seed;
  x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  p[i] = __float2bfloat16(((x & 0xFFFF) / 65536.0f - 0.5f) * amp);
}
__global__ void fillf(float* p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int main(int argc, char** argv) {
  // usage: [B H Nq Nk causal]   (default: the benchmark shape)
  const int B = argc > 1 ? atoi(argv[1]) : 4, H = argc > 2 ? atoi(argv[2]) : 8, D = 64;
  const int Nq = argc > 3 ? atoi(argv[3]) : 4096, Nk = argc > 4 ? atoi(argv[4]) : 4096;
  const int causal = argc > 5 ? atoi(argv[5]) : 1;
  const int N = Nq > Nk ? Nq : Nk;                       // allocation size of every tensor
  const size_t n = (size_t)B * H * N * D;
  __nv_bfloat16 *q, *k, *v, *o, *d_o, *dq, *dk, *dv;
  float* inv_l;
  for (auto pp : {&q, &k, &v, &o, &d_o, &dq, &dk, &dv}) CK(cudaMalloc(pp, n * 2));
  CK(cudaMalloc(&inv_l, (size_t)B * H * N * 4));
  fill<<<(n + 255) / 256, 256>>>(q, n, 1, 0.25f);     // |q.k| small: p ~ exp(-8)...; timing only
  fill<<<(n + 255) / 256, 256>>>(k, n, 2, 0.25f);
  fill<<<(n + 255) / 256, 256>>>(v, n, 3, 2.f);
  fill<<<(n + 255) / 256, 256>>>(o, n, 4, 1.f);
  fill<<<(n + 255) / 256, 256>>>(d_o, n, 5, 2.f);
  fillf<<<((size_t)B * H * N + 255) / 256, 256>>>(inv_l, (size_t)B * H * N, 1.0f);
  size_t wsb = fcsa::bwd_workspace_bytes(B, H, H, Nq, Nk, D);
  void* ws;
  CK(cudaMalloc(&ws, wsb));
  size_t zsb = fcsa::bwd_zeroed_workspace_bytes(B, H, H, Nq, Nk, D);
  void* zs;
  CK(cudaMalloc(&zs, zsb));
  CK(cudaMemset(zs, 0, zsb));
  fcsa::BwdHostArgs h;
  h.dtype_bf16 = true; h.B = B; h.H = H; h.kv_heads = H; h.Nq = Nq; h.Nk = Nk; h.D = D; h.causal = causal;
  h.scale = 8.f; h.shift = 8.f; h.mask = nullptr; h.mask_sb = 0;
  auto T = [&](void* p) { fcsa_tensor t; t.ptr = p; t.sb = (long long)H * N * D; t.sh = (long long)N * D; t.sn = D; return t; };
  h.q = T(q); h.k = T(k); h.v = T(v); h.o = T(o); h.d_o = T(d_o); h.dq = T(dq); h.dk = T(dk); h.dv = T(dv);
  h.inv_l = inv_l; h.workspace = ws; h.zeroed = zs;
  int launches = 0; const char* err = nullptr; cudaError_t ce = cudaSuccess;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  h.ev_start = e0; h.ev_stop = e1;
#ifdef FCSA_TRACE
  const int reps = 3;
#else
  const int reps = 12;
#endif
  float best = 1e9f;
  for (int rep = 0; rep < reps; ++rep) {
    int r = fcsa::run_backward(h, 0, &launches, &err, &ce);
    if (r) { printf("run_backward failed %d %s\n", r, err ? err : ""); return 1; }
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
#ifdef FCSA_TRACE
    printf("rep %d: main kernel %.1f us\n", rep, ms * 1e3);
#endif
  }
  printf("B %d H %d Nq %d Nk %d causal %d: main kernel best of %d: %.1f us\n", B, H, Nq, Nk, causal, reps, best * 1e3);
#ifdef FCSA_CTA_TIMELINE
  {
    // one row per work item (key tile x batch x head); per SM: the items it ran, in start order
    const int nitems = ((Nk + 127) / 128) * B * H;
    static long long ct[4096][10];
    CK(cudaMemcpyFromSymbol(ct, g_fcsa_cta_t, sizeof(ct)));
    printf("per item: start(after previous item's last dS) | start->first S seen, ->first dS handed over, ->last dS handed over, ->dK/dV stored; tiles\n");
    double s_first = 0, s_ds = 0, s_body = 0, s_epi = 0, s_gap = 0; int n = 0, ngap = 0; double tiles = 0;
    for (int sm = 0; sm < 148; ++sm) {
      std::vector<int> ids;
      for (int i = 0; i < nitems && i < 4096; ++i) if (ct[i][8] == sm && ct[i][0]) ids.push_back(i);
      std::sort(ids.begin(), ids.end(), [&](int x, int y) { return ct[x][0] < ct[y][0]; });
      for (size_t k2 = 0; k2 < ids.size(); ++k2) {
        const int id = ids[k2];
        const int jt = id / (B * H);
        int ni = (Nq + 127) / 128;
        if (causal) { int x = jt * 128 - (Nk - Nq) - 127; int ilo = x <= 0 ? 0 : (x + 127) / 128; ni -= ilo; }
        const long long gap = k2 ? ct[id][0] - ct[ids[k2 - 1]][4] : 0;
        if (k2) { s_gap += gap; ++ngap; }
        s_first += ct[id][2] - ct[id][0]; s_ds += ct[id][3] - ct[id][2]; s_body += ct[id][4] - ct[id][3]; s_epi += ct[id][6] - ct[id][4];
        tiles += ni - 1; ++n;
        if (sm < 2)
          printf("sm %d item %4d (key tile %2d, %2d tiles): after prev last dS %6lld | %6lld %6lld %7lld %6lld\n", sm, id, jt, ni, gap,
                 ct[id][2] - ct[id][0], ct[id][3] - ct[id][2], ct[id][4] - ct[id][3], ct[id][6] - ct[id][4]);
      }
    }
    printf("means (cycles): prev last dS -> item start %.0f | start -> first S %.0f, -> first dS %.0f, first -> last dS %.0f (%.0f per tile), last dS -> dK/dV stored %.0f  (%d items)\n",
           s_gap / (ngap ? ngap : 1), s_first / n, s_ds / n, s_body / n, s_body / tiles, s_epi / n, n);
  }
#endif
#ifdef FCSA_TRACE
  static long long tr[8][48][8];
  CK(cudaMemcpyFromSymbol(tr, g_fcsa_trace, sizeof(tr)));
  long long t0 = tr[1][0][0];
  const char* names[6] = {"MMA  [S(i+1) issued, dV issued, dK issued, dP(i+1) issued, dQ issued]",
                          "EXP  [top, S_FULL seen, S in regs, exps done, P arrived]",
                          "DS   [DP_FULL seen, dP in regs, DS_FREE ok, DS arrived, -]",
                          "RED  [DQ_FULL seen, reduce issued]",
                          "TMA  [Q_EMPTY seen, DO_EMPTY seen]", "OBS  [S_FULL, DP_FULL, PV_DONE, DQ_FULL complete]"};
  int nslots[6] = {5, 5, 5, 2, 2, 4};
  for (int role = 0; role < 6; ++role) {
    printf("--- %s\n", names[role]);
    for (int i = 0; i < 12; ++i) {
      printf("  it %2d:", i);
      for (int s = 0; s < nslots[role]; ++s) printf(" %8lld", tr[role][i][s] ? tr[role][i][s] - t0 : -1);
      printf("\n");
    }
  }
  // iteration period from the MMA role
  printf("MMA iteration period (cycles):");
  for (int i = 1; i < 32; ++i) printf(" %lld", tr[1][i][4] - tr[1][i - 1][4]);
  printf("\n");
#endif
  return 0;
}
