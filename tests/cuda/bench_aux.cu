// Times the HBM-bound helper kernels (l2norm pair, backward preprocess) at the benchmark
// shape with an L2 flush between repetitions.  Test infrastructure only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bench_aux bench_aux.cu
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../flash_cosine_sim_attention_b200/csrc/bwd_kernel.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
typedef __nv_bfloat16 bf16;

int main() {
  const int B = 4, H = 8, N = 4096, D = 64, G = 1;
  const size_t n = (size_t)B * H * N * D;
  bf16 *q, *k, *qn, *kn, *o, *d_o, *dq;
  float *rq, *rk, *inv_l, *ws;
  for (auto pp : {&q, &k, &qn, &kn, &o, &d_o, &dq}) CK(cudaMalloc(pp, n * 2));
  CK(cudaMemset(q, 0x3c, n * 2)); CK(cudaMemset(k, 0x3c, n * 2)); CK(cudaMemset(o, 0x3c, n * 2)); CK(cudaMemset(d_o, 0x3c, n * 2));
  CK(cudaMalloc(&rq, (size_t)B * H * N * G * 4)); CK(cudaMalloc(&rk, (size_t)B * H * N * G * 4));
  CK(cudaMalloc(&inv_l, (size_t)B * H * N * 4));
  CK(cudaMemset(inv_l, 0x3f, (size_t)B * H * N * 4));
  const fcsa::BwdWorkspace w = fcsa::bwd_workspace_layout(B, H, H, N, N, D);
  CK(cudaMalloc(&ws, w.total));
  void* flush; CK(cudaMalloc(&flush, 256u << 20));
  const long long sb = (long long)H * N * D, sh = (long long)N * D, sn = D;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time_it = [&](const char* name, auto launch, double bytes, bool do_flush = true) {
    float best = 1e9, sum = 0;
    for (int rep = 0; rep < 12; ++rep) {
      if (do_flush) CK(cudaMemsetAsync(flush, rep, 256u << 20));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    printf("%-34s %-9s mean %.1f us  best %.1f us  -> %.0f GB/s (algorithmic %.1f MB)\n", name, do_flush ? "L2 cold" : "L2 warm",
           sum / 10 * 1e3, best * 1e3, bytes / (sum / 10 * 1e-3) / 1e9, bytes / 1e6);
  };
  // l2norm pair
  fcsa::L2PairArgs pa; memset(&pa, 0, sizeof(pa));
  for (int t = 0; t < 2; ++t) {
    fcsa::L2Args& a = pa.t[t];
    a.B = B; a.H = H; a.N = N; a.D = D; a.G = G;
    a.x_sb = a.y_sb = sb; a.x_sh = a.y_sh = sh; a.x_sn = a.y_sn = sn;
    a.x = t ? k : q; a.y = t ? kn : qn; a.rnorm = t ? rk : rq;
  }
  {
    dim3 grid((N + 127) / 128, 2 * B * H);
    time_it("l2norm_fwd_pair", [&] { fcsa::l2norm_fwd_pair_kernel<bf16, 4, 2><<<grid, 256>>>(pa); }, 4.0 * n * 2);
    time_it("l2norm_fwd_pair", [&] { fcsa::l2norm_fwd_pair_kernel<bf16, 4, 2><<<grid, 256>>>(pa); }, 4.0 * n * 2, false);
  }
  // prep (per-row constants + slivers; no accumulator zeroing any more)
  {
    fcsa::PrepArgs p; memset(&p, 0, sizeof(p));
    p.aug = (char*)ws + w.aug_off; p.ones = (char*)ws + w.ones_off; p.inv_c1 = 1.f; p.B = B; p.H = H; p.Nq = N; p.Nk = N; p.D = D;
    p.nqt = w.nqt; p.QT = w.QT; p.c2 = 11.5f; p.causal = 1; p.shift_extra = nullptr;
    p.o = o; p.o_sb = sb; p.o_sh = sh; p.o_sn = sn; p.d_o = d_o; p.do_sb = sb; p.do_sh = sh; p.do_sn = sn;
    p.inv_l = inv_l; p.stats = (float*)((char*)ws + w.stats_off);
    const int rows_per_block = 2 * (256 / (D / 8));
    p.bpb = (w.nqt * w.QT + rows_per_block - 1) / rows_per_block;
    dim3 grid(p.bpb * B * H);
    time_it("bwd_prep (delta, slivers)", [&] { fcsa::bwd_prep_kernel<bf16, 8><<<grid, 256>>>(p); }, 2.0 * n * 2 + (double)B * H * N * 64);
    time_it("bwd_prep (delta, slivers)", [&] { fcsa::bwd_prep_kernel<bf16, 8><<<grid, 256>>>(p); }, 2.0 * n * 2 + (double)B * H * N * 64, false);
  }
  // dq conversion (reads the fp32 accumulator + q_hat, writes dq, clears the accumulator)
  {
    float* zs; CK(cudaMalloc(&zs, w.ztotal)); CK(cudaMemset(zs, 0, w.ztotal));
    fcsa::BwdArgs f; memset(&f, 0, sizeof(f));
    f.B = B; f.H = H; f.Nq = N; f.Nk = N; f.nqt = w.nqt; f.scale = 8.f; f.dq_acc = zs;
    f.dq = dq; f.dq_sb = sb; f.dq_sh = sh; f.dq_sn = sn; f.q_hat = qn; f.q_sb = sb; f.q_sh = sh; f.q_sn = sn;
    f.q_rnorm = rq; f.G = G;
    dim3 grid(w.nqt * B * H);
    const double bytes = n * 4.0 * 2 + 2.0 * n * 2;
    time_it("bwd_dq_finish64 (+l2 bwd, clears)", [&] { fcsa::bwd_dq_finish64_kernel<bf16><<<grid, 512>>>(f); }, bytes);
    time_it("bwd_dq_finish64 (+l2 bwd, clears)", [&] { fcsa::bwd_dq_finish64_kernel<bf16><<<grid, 512>>>(f); }, bytes, false);
    f.q_rnorm = nullptr;
    time_it("bwd_dq_finish64 (plain, clears)", [&] { fcsa::bwd_dq_finish64_kernel<bf16><<<grid, 512>>>(f); }, n * 4.0 * 2 + n * 2.0, false);
  }
  // reference points: plain copies of the same sizes
  time_it("cudaMemcpy D2D 2 x 16.8 MB", [&] { CK(cudaMemcpyAsync(qn, q, n * 2, cudaMemcpyDeviceToDevice)); CK(cudaMemcpyAsync(kn, k, n * 2, cudaMemcpyDeviceToDevice)); }, 4.0 * n * 2);
  time_it("cudaMemcpy D2D 2 x 16.8 MB", [&] { CK(cudaMemcpyAsync(qn, q, n * 2, cudaMemcpyDeviceToDevice)); CK(cudaMemcpyAsync(kn, k, n * 2, cudaMemcpyDeviceToDevice)); }, 4.0 * n * 2, false);
  return 0;
}
