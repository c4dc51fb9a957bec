// Grouped l2-normalisation of q / k rows and its backward.
//
// Replaces the PyTorch ops the reference runs outside its kernels
// (flash_cosine_sim_attention.py:38-65: F.normalize per `groups` chunk of the head dim, cast
// back to the input dtype; backward left to autograd).  One pass each way, HBM-bound:
// 16-byte vector loads, fp32 reduction with warp shuffles, one thread owns 8 features.
#pragma once

#include "sm100_primitives.cuh"

namespace fcsa {

struct L2Args {
  int B, H, N, D, G;
  long long x_sb, x_sh, x_sn;     // input  (x for forward, dy for backward)
  long long y_sb, y_sh, y_sn;     // normalised tensor (output of forward, input of backward)
  long long o_sb, o_sh, o_sn;     // dx (backward only)
  const void* x;
  void* y;
  void* dx;
  float* rnorm;                   // (B, H, N, G) fp32
};

// sum over the `tpg` consecutive lanes that share a group (tpg is a power of two <= 16)
__device__ __forceinline__ float group_reduce(float v, int tpg) {
  for (int m = 1; m < tpg; m <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, m);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const L2Args a) {
  const int tpr = a.D >> 3;                       // threads per row (8 features each)
  const int rows_per_block = 256 / tpr;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / tpr;
  const int tr = threadIdx.x % tpr;
  const long long total_rows = (long long)a.B * a.H * a.N;
  const bool ok = row < total_rows;
  const long long rr = ok ? row : 0;
  const int n = (int)(rr % a.N);
  const int h = (int)((rr / a.N) % a.H);
  const int b = (int)(rr / ((long long)a.N * a.H));
  const T* xp = reinterpret_cast<const T*>(a.x) + b * a.x_sb + h * a.x_sh + n * a.x_sn + tr * 8;
  uint4 raw = make_uint4(0, 0, 0, 0);
  if (ok) raw = *reinterpret_cast<const uint4*>(xp);
  float f[8];
  {
    float2 t0 = unpack2<T>(raw.x), t1 = unpack2<T>(raw.y), t2 = unpack2<T>(raw.z), t3 = unpack2<T>(raw.w);
    f[0] = t0.x; f[1] = t0.y; f[2] = t1.x; f[3] = t1.y; f[4] = t2.x; f[5] = t2.y; f[6] = t3.x; f[7] = t3.y;
  }
  const int gs = a.D / a.G;                        // features per group
  float rn[8];
  if (gs >= 8) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
    const int tpg = gs >> 3;
    ss = group_reduce(ss, tpg);
    const float r = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = r;
    if (ok && a.rnorm && (tr % tpg) == 0) a.rnorm[row * a.G + tr / tpg] = r;
  } else {
    // several groups inside one thread's 8 features (gs = 1, 2 or 4)
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = 0.f;
    for (int g0 = 0; g0 < 8; g0 += gs) {
      float ss = 0.f;
      for (int i = 0; i < gs; ++i) ss += f[g0 + i] * f[g0 + i];
      const float r = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      for (int i = 0; i < gs; ++i) rn[g0 + i] = r;
      if (ok && a.rnorm) a.rnorm[row * a.G + (tr * 8 + g0) / gs] = r;
    }
  }
  if (ok) {
    uint4 w;
    w.x = pack2<T>(f[0] * rn[0], f[1] * rn[1]);
    w.y = pack2<T>(f[2] * rn[2], f[3] * rn[3]);
    w.z = pack2<T>(f[4] * rn[4], f[5] * rn[5]);
    w.w = pack2<T>(f[6] * rn[6], f[7] * rn[7]);
    T* yp = reinterpret_cast<T*>(a.y) + b * a.y_sb + h * a.y_sh + n * a.y_sn + tr * 8;
    *reinterpret_cast<uint4*>(yp) = w;
  }
}

// q and k in one launch (blockIdx.y selects the tensor), two rows per thread so that two
// independent 16-byte loads are in flight before the first reduction.
struct L2PairArgs {
  L2Args t[2];
};

template <typename T>
__global__ void __launch_bounds__(256) l2norm_fwd_pair_kernel(const L2PairArgs pa) {
  // grid = (row blocks, batch*heads, tensor): no per-thread integer division on the address path
  pdl_launch_dependents();
  pdl_wait();
  const L2Args& a = pa.t[blockIdx.z];
  const int bh = blockIdx.y;
  if (bh >= a.B * a.H) return;                    // k may have fewer heads than q (block-uniform exit)
  const int b = bh / a.H, h = bh - b * a.H;
  const int tpr = a.D >> 3;
  const int rows_per_block = 256 / tpr;
  const int tr = threadIdx.x % tpr;
  const int gs = a.D / a.G;
  const T* xbase = reinterpret_cast<const T*>(a.x) + b * a.x_sb + h * a.x_sh + tr * 8;
  T* ybase = reinterpret_cast<T*>(a.y) + b * a.y_sb + h * a.y_sh + tr * 8;
  int nn[2];
  bool ok[2];
  uint4 raw[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    nn[u] = (blockIdx.x * 2 + u) * rows_per_block + threadIdx.x / tpr;
    ok[u] = nn[u] < a.N;
    raw[u] = make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
    if (ok[u]) raw[u] = *reinterpret_cast<const uint4*>(xbase + (long long)nn[u] * a.x_sn);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long row = (long long)bh * a.N + nn[u];
    float f[8];
    {
      float2 t0 = unpack2<T>(raw[u].x), t1 = unpack2<T>(raw[u].y), t2 = unpack2<T>(raw[u].z), t3 = unpack2<T>(raw[u].w);
      f[0] = t0.x; f[1] = t0.y; f[2] = t1.x; f[3] = t1.y; f[4] = t2.x; f[5] = t2.y; f[6] = t3.x; f[7] = t3.y;
    }
    float rn[8];
    if (gs >= 8) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
      const int tpg = gs >> 3;
      ss = group_reduce(ss, tpg);
      const float r = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < 8; ++i) rn[i] = r;
      if (ok[u] && a.rnorm && (tr % tpg) == 0) a.rnorm[row * a.G + tr / tpg] = r;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) rn[i] = 0.f;
      for (int g0 = 0; g0 < 8; g0 += gs) {
        float ss = 0.f;
        for (int i = 0; i < gs; ++i) ss += f[g0 + i] * f[g0 + i];
        const float r = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        for (int i = 0; i < gs; ++i) rn[g0 + i] = r;
        if (ok[u] && a.rnorm) a.rnorm[row * a.G + (tr * 8 + g0) / gs] = r;
      }
    }
    if (ok[u]) {
      uint4 w;
      w.x = pack2<T>(f[0] * rn[0], f[1] * rn[1]);
      w.y = pack2<T>(f[2] * rn[2], f[3] * rn[3]);
      w.z = pack2<T>(f[4] * rn[4], f[5] * rn[5]);
      w.w = pack2<T>(f[6] * rn[6], f[7] * rn[7]);
      *reinterpret_cast<uint4*>(ybase + (long long)nn[u] * a.y_sn) = w;
    }
  }
}

// dx = (dy - y * <y, dy>_group) * rnorm_group
template <typename T>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const L2Args a) {
  const int tpr = a.D >> 3;
  const int rows_per_block = 256 / tpr;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / tpr;
  const int tr = threadIdx.x % tpr;
  const long long total_rows = (long long)a.B * a.H * a.N;
  const bool ok = row < total_rows;
  const long long rr = ok ? row : 0;
  const int n = (int)(rr % a.N);
  const int h = (int)((rr / a.N) % a.H);
  const int b = (int)(rr / ((long long)a.N * a.H));
  const T* dyp = reinterpret_cast<const T*>(a.x) + b * a.x_sb + h * a.x_sh + n * a.x_sn + tr * 8;
  const T* yp = reinterpret_cast<const T*>(a.y) + b * a.y_sb + h * a.y_sh + n * a.y_sn + tr * 8;
  uint4 rdy = make_uint4(0, 0, 0, 0), ry = make_uint4(0, 0, 0, 0);
  if (ok) {
    rdy = *reinterpret_cast<const uint4*>(dyp);
    ry = *reinterpret_cast<const uint4*>(yp);
  }
  float dy[8], y[8];
  {
    float2 t0 = unpack2<T>(rdy.x), t1 = unpack2<T>(rdy.y), t2 = unpack2<T>(rdy.z), t3 = unpack2<T>(rdy.w);
    dy[0] = t0.x; dy[1] = t0.y; dy[2] = t1.x; dy[3] = t1.y; dy[4] = t2.x; dy[5] = t2.y; dy[6] = t3.x; dy[7] = t3.y;
    float2 u0 = unpack2<T>(ry.x), u1 = unpack2<T>(ry.y), u2 = unpack2<T>(ry.z), u3 = unpack2<T>(ry.w);
    y[0] = u0.x; y[1] = u0.y; y[2] = u1.x; y[3] = u1.y; y[4] = u2.x; y[5] = u2.y; y[6] = u3.x; y[7] = u3.y;
  }
  const int gs = a.D / a.G;
  float out[8];
  if (gs >= 8) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dot += y[i] * dy[i];
    const int tpg = gs >> 3;
    dot = group_reduce(dot, tpg);
    const float r = ok ? a.rnorm[row * a.G + tr / tpg] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = (dy[i] - y[i] * dot) * r;
  } else {
    for (int g0 = 0; g0 < 8; g0 += gs) {
      float dot = 0.f;
      for (int i = 0; i < gs; ++i) dot += y[g0 + i] * dy[g0 + i];
      const float r = ok ? a.rnorm[row * a.G + (tr * 8 + g0) / gs] : 0.f;
      for (int i = 0; i < gs; ++i) out[g0 + i] = (dy[g0 + i] - y[g0 + i] * dot) * r;
    }
  }
  if (ok) {
    uint4 w;
    w.x = pack2<T>(out[0], out[1]);
    w.y = pack2<T>(out[2], out[3]);
    w.z = pack2<T>(out[4], out[5]);
    w.w = pack2<T>(out[6], out[7]);
    T* op = reinterpret_cast<T*>(a.dx) + b * a.o_sb + h * a.o_sh + n * a.o_sn + tr * 8;
    *reinterpret_cast<uint4*>(op) = w;
  }
}

}  // namespace fcsa
