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

// 1 / max(sqrt(ss), 1e-12)  (py:38-55: F.normalize's eps) = min(rsqrt(ss), 1e12): one MUFU.RSQ instead of an
// IEEE sqrt + divide (~60 instructions per row and thread in a pass that must stay HBM-bound); the
// approximation error (2^-22) vanishes under the 16-bit rounding of the normalised output
__device__ __forceinline__ float rnorm_of(float ss) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ss));
  return fminf(r, 1e12f);
}

// sum over the `tpg` consecutive lanes that share a group (tpg is a power of two <= 16)
__device__ __forceinline__ float group_reduce(float v, int tpg) {
  for (int m = 1; m < tpg; m <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, m);
  return v;
}

// Groups narrower than the 8 features one thread owns (group size gs = 1, 2 or 4): for every
// feature, the sum of v over the aligned gs-wide block it belongs to.  Written with compile-time
// register indices only - a `for (i < gs)` loop over the arrays would push them to local memory.
__device__ __forceinline__ void subgroup_sums8(const float (&v)[8], int gs, float (&out)[8]) {
  float p2[8], p4[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p2[i] = v[i] + v[i ^ 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) p4[i] = p2[i] + p2[i ^ 2];
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = gs == 1 ? v[i] : (gs == 2 ? p2[i] : p4[i]);
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
    const float r = rnorm_of(ss);
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = r;
    if (ok && a.rnorm && (tr % tpg) == 0) a.rnorm[row * a.G + tr / tpg] = r;
  } else {
    // several groups inside one thread's 8 features (gs = 1, 2 or 4)
    float sq[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sq[i] = f[i] * f[i];
    subgroup_sums8(sq, gs, ss);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      rn[i] = rnorm_of(ss[i]);
      if (ok && a.rnorm && (i & (gs - 1)) == 0) a.rnorm[row * a.G + (tr * 8 + i) / gs] = rn[i];
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

// q and k in one launch.  grid = (row blocks, batch*heads of q + batch*heads of k): blockIdx.y selects the
// tensor and its (batch, head), so no thread divides 64-bit indices; one thread owns 16 features of U rows
// and issues its 2 U 16-byte loads before the first reduction.  TPR = threads per row (D / 16) when known at
// compile time (4 or 8), 0 = read it from the arguments (head dims 32 and 256: D must be a multiple of 16).  A short-lived-CTA grid (many waves of small CTAs,
// 8 resident per SM) measured faster here than a persistent loop: the per-item bookkeeping of the loop
// (item decode with 64-bit divisions, register copies of the prefetched rows) cost more instructions than the
// rows themselves (profiles/r02_aux_kernels.txt).
struct L2PairArgs {
  L2Args t[2];
};

template <typename T, int TPR, int U>
__global__ void __launch_bounds__(256) l2norm_fwd_pair_kernel(const L2PairArgs pa) {
  // TPR threads per row, SIXTEEN features per thread (two 16-byte loads): the pass is issue-bound before it is
  // memory-bound (ncu: 84 % issue-slot utilisation with 8 features per thread), so the per-row bookkeeping
  // (addresses, shuffle steps, predicates) is spread over twice the bytes.
  pdl_launch_dependents();
  pdl_wait();
  const int bh0 = pa.t[0].B * pa.t[0].H;
  const int which = blockIdx.y >= bh0;
  const L2Args& a = pa.t[which];
  const int bh = blockIdx.y - (which ? bh0 : 0);
  const int tpr = TPR ? TPR : (a.D >> 4);
  const int rpp = 256 / tpr;                          // rows per pass
  const int tr = threadIdx.x % tpr;
  const int r_in = threadIdx.x / tpr;
  const int row0 = blockIdx.x * (U * rpp);
  if (row0 >= a.N) return;                            // the grid is sized for the longer of the two tensors
  const int b = bh / a.H, h = bh - b * a.H;
  const T* xbase = reinterpret_cast<const T*>(a.x) + b * a.x_sb + h * a.x_sh + tr * 16;
  T* ybase = reinterpret_cast<T*>(a.y) + b * a.y_sb + h * a.y_sh + tr * 16;
  uint4 raw[U][2];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int n = row0 + u * rpp + r_in;
    raw[u][0] = raw[u][1] = make_uint4(0, 0, 0, 0);
    if (n < a.N) {
      raw[u][0] = ldg_stream128(xbase + (long long)n * a.x_sn);
      raw[u][1] = ldg_stream128(xbase + (long long)n * a.x_sn + 8);
    }
  }
  const int gs = a.D / a.G;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int n = row0 + u * rpp + r_in;
    const bool ok = n < a.N;
    const long long row = (long long)bh * a.N + n;
    float f[16];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const float2 t0 = unpack2<T>(raw[u][s2].x), t1 = unpack2<T>(raw[u][s2].y), t2 = unpack2<T>(raw[u][s2].z),
                   t3 = unpack2<T>(raw[u][s2].w);
      f[8 * s2] = t0.x; f[8 * s2 + 1] = t0.y; f[8 * s2 + 2] = t1.x; f[8 * s2 + 3] = t1.y;
      f[8 * s2 + 4] = t2.x; f[8 * s2 + 5] = t2.y; f[8 * s2 + 6] = t3.x; f[8 * s2 + 7] = t3.y;
    }
    float rn[16];
    if (gs >= 16) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) ss += f[i] * f[i];
      const int tpg = gs >> 4;                         // threads per group
      ss = group_reduce(ss, tpg);
      const float r = rnorm_of(ss);
#pragma unroll
      for (int i = 0; i < 16; ++i) rn[i] = r;
      if (ok && a.rnorm && (tr & (tpg - 1)) == 0) a.rnorm[row * a.G + tr / tpg] = r;
    } else if (gs == 8) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += f[8 * s2 + i] * f[8 * s2 + i];
        const float r = rnorm_of(ss);
#pragma unroll
        for (int i = 0; i < 8; ++i) rn[8 * s2 + i] = r;
        if (ok && a.rnorm) a.rnorm[row * a.G + 2 * tr + s2] = r;
      }
    } else {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        float sq[8], ss[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) sq[i] = f[8 * s2 + i] * f[8 * s2 + i];
        subgroup_sums8(sq, gs, ss);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rn[8 * s2 + i] = rnorm_of(ss[i]);
          if (ok && a.rnorm && (i & (gs - 1)) == 0) a.rnorm[row * a.G + (tr * 16 + 8 * s2 + i) / gs] = rn[8 * s2 + i];
        }
      }
    }
    if (ok) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        uint4 wv;
        wv.x = pack2<T>(f[8 * s2 + 0] * rn[8 * s2 + 0], f[8 * s2 + 1] * rn[8 * s2 + 1]);
        wv.y = pack2<T>(f[8 * s2 + 2] * rn[8 * s2 + 2], f[8 * s2 + 3] * rn[8 * s2 + 3]);
        wv.z = pack2<T>(f[8 * s2 + 4] * rn[8 * s2 + 4], f[8 * s2 + 5] * rn[8 * s2 + 5]);
        wv.w = pack2<T>(f[8 * s2 + 6] * rn[8 * s2 + 6], f[8 * s2 + 7] * rn[8 * s2 + 7]);
        *reinterpret_cast<uint4*>(ybase + (long long)n * a.y_sn + 8 * s2) = wv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// float32 front end (reference: Float is dispatched in forward and backward, cu:1702-1703 / 1832-1834).  float32
// callers run on the 16-bit tensor-core kernels; these two passes are the only extra work around them:
//   f32_cast_kernel     : y(16 bit) = round( x / max(||x||_group, 1e-12) * mul )   G > 0: l2norm over G groups
//                                     round( x * mul )                              G = 0: plain scaled cast
//                         `mul` is an optional DEVICE scalar (the power-of-two range scale of v / dO, chosen on
//                         the device - no host synchronisation); rnorm (fp32, per group) is kept for the backward
//   f32_cast_bwd_kernel : dx(f32) = (dy - y <y, dy>_group) * rnorm_group * mul      G > 0   (dy: f32 gradient
//                                   dy * mul                                         G = 0    w.r.t. y; y: 16 bit)
// One thread owns 8 features (two float4 loads); TPR threads per row.
// ------------------------------------------------------------------------------------------------
struct F32CastArgs {
  int B, H, N, D, G;
  long long x_sb, x_sh, x_sn;     // float32 tensor (x forward, dy / dx backward): element strides
  long long y_sb, y_sh, y_sn;     // 16-bit tensor y
  long long o_sb, o_sh, o_sn;     // backward only: dx (float32)
  const float* x;
  void* y;
  float* dx;
  float* rnorm;                   // (B, H, N, G) fp32 or nullptr (G = 0)
  const float* mul;               // device scalar or nullptr (= 1)
  int mul_reciprocal;             // 1: use 1 / *mul
};

__device__ __forceinline__ float f32cast_mul(const F32CastArgs& a) {
  if (a.mul == nullptr) return 1.f;
  const float m = __ldg(a.mul);
  return a.mul_reciprocal ? 1.f / m : m;
}

template <typename T>
__global__ void __launch_bounds__(256) f32_cast_kernel(const F32CastArgs a) {
  const int tpr = a.D >> 3;
  const int rpb = 256 / tpr;
  const long long row = (long long)blockIdx.x * rpb + threadIdx.x / tpr;
  const int tr = threadIdx.x % tpr;
  const long long total = (long long)a.B * a.H * a.N;
  const bool ok = row < total;
  const long long rr = ok ? row : 0;
  const int n = (int)(rr % a.N);
  const int h = (int)((rr / a.N) % a.H);
  const int b = (int)(rr / ((long long)a.N * a.H));
  const float* xp = a.x + b * a.x_sb + h * a.x_sh + (long long)n * a.x_sn + tr * 8;
  float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
  if (ok) {
    lo = __ldg(reinterpret_cast<const float4*>(xp));
    hi = __ldg(reinterpret_cast<const float4*>(xp + 4));
  }
  float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  const float mul = f32cast_mul(a);
  float rn[8];
  if (a.G == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = mul;
  } else {
    const int gs = a.D / a.G;
    if (gs >= 8) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
      const int tpg = gs >> 3;
      ss = group_reduce(ss, tpg);
      const float r = rnorm_of(ss);
#pragma unroll
      for (int i = 0; i < 8; ++i) rn[i] = r * mul;
      if (ok && a.rnorm && (tr & (tpg - 1)) == 0) a.rnorm[row * a.G + tr / tpg] = r;
    } else {
      float sq[8], ss[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) sq[i] = f[i] * f[i];
      subgroup_sums8(sq, gs, ss);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float r = rnorm_of(ss[i]);
        rn[i] = r * mul;
        if (ok && a.rnorm && (i & (gs - 1)) == 0) a.rnorm[row * a.G + (tr * 8 + i) / gs] = r;
      }
    }
  }
  if (ok) {
    uint4 w;
    w.x = pack2<T>(f[0] * rn[0], f[1] * rn[1]);
    w.y = pack2<T>(f[2] * rn[2], f[3] * rn[3]);
    w.z = pack2<T>(f[4] * rn[4], f[5] * rn[5]);
    w.w = pack2<T>(f[6] * rn[6], f[7] * rn[7]);
    T* yp = reinterpret_cast<T*>(a.y) + b * a.y_sb + h * a.y_sh + (long long)n * a.y_sn + tr * 8;
    *reinterpret_cast<uint4*>(yp) = w;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) f32_cast_bwd_kernel(const F32CastArgs a) {
  const int tpr = a.D >> 3;
  const int rpb = 256 / tpr;
  const long long row = (long long)blockIdx.x * rpb + threadIdx.x / tpr;
  const int tr = threadIdx.x % tpr;
  const long long total = (long long)a.B * a.H * a.N;
  const bool ok = row < total;
  const long long rr = ok ? row : 0;
  const int n = (int)(rr % a.N);
  const int h = (int)((rr / a.N) % a.H);
  const int b = (int)(rr / ((long long)a.N * a.H));
  const float* dyp = a.x + b * a.x_sb + h * a.x_sh + (long long)n * a.x_sn + tr * 8;
  float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
  uint4 ry = make_uint4(0, 0, 0, 0);
  if (ok) {
    lo = __ldg(reinterpret_cast<const float4*>(dyp));
    hi = __ldg(reinterpret_cast<const float4*>(dyp + 4));
    if (a.G > 0)
      ry = ldg_stream128(reinterpret_cast<const T*>(a.y) + b * a.y_sb + h * a.y_sh + (long long)n * a.y_sn + tr * 8);
  }
  float dy[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  const float mul = f32cast_mul(a);
  float out[8];
  if (a.G == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = dy[i] * mul;
  } else {
    float y[8];
    {
      const float2 u0 = unpack2<T>(ry.x), u1 = unpack2<T>(ry.y), u2 = unpack2<T>(ry.z), u3 = unpack2<T>(ry.w);
      y[0] = u0.x; y[1] = u0.y; y[2] = u1.x; y[3] = u1.y; y[4] = u2.x; y[5] = u2.y; y[6] = u3.x; y[7] = u3.y;
    }
    const int gs = a.D / a.G;
    if (gs >= 8) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dot += y[i] * dy[i];
      const int tpg = gs >> 3;
      dot = group_reduce(dot, tpg);
      const float r = (ok ? a.rnorm[row * a.G + tr / tpg] : 0.f) * mul;
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = (dy[i] - y[i] * dot) * r;
    } else {
      float pr[8], dot[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pr[i] = y[i] * dy[i];
      subgroup_sums8(pr, gs, dot);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float r = (ok ? a.rnorm[row * a.G + (tr * 8 + i) / gs] : 0.f) * mul;
        out[i] = (dy[i] - y[i] * dot[i]) * r;
      }
    }
  }
  if (ok) {
    float* op = a.dx + b * a.o_sb + h * a.o_sh + (long long)n * a.o_sn + tr * 8;
    *reinterpret_cast<float4*>(op) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(out[4], out[5], out[6], out[7]);
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
    float pr[8], dot[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pr[i] = y[i] * dy[i];
    subgroup_sums8(pr, gs, dot);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float r = ok ? a.rnorm[row * a.G + (tr * 8 + i) / gs] : 0.f;
      out[i] = (dy[i] - y[i] * dot[i]) * r;
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
