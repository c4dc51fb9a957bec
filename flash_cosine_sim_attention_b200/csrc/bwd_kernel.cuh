// Backward: dq, dk, dv of the fixed-shift cosine-sim attention.
//
// Replaces the reference's backward_preprocess (flash_cosine_sim_attention_cuda.cu:1256-1335)
// and backward_kernel (cu:1339-1626).  The math is the reference's (restated in SURVEY.md par. 0):
//     P_ij  = exp(scale*s_ij - shift) * inv_l_i          cu:1513-1526
//     delta = rowsum(dO * O)                               cu:1293-1334
//     dV    = P^T dO          dP = dO V^T                  cu:1534-1553
//     dS    = P * (dP - delta)                             cu:1564-1570
//     dK    = scale * dS^T q  dQ = scale * dS k            cu:1580-1610
// The machine mapping is new.  Three kernels:
//   1. bwd_prep_kernel    : per query row {c3 = log2(inv_l) - shift*log2e, delta} (fp32, never
//                           rounded to 16 bit as the reference does at cu:1260/1820), for D = 64 also as
//                           16-bit "slivers" (see 2.)
//   2. fcsa_bwd_kernel    : work item = (key tile of 128, batch, head), key/value tile stationary, loop over the
//                           query tiles of QT rows that can see it.  D = 64: persistent CTAs (one per SM) walk the
//                           items - continuous query-tile stream, next K prefetched into V's shared-memory buffer,
//                           dQ drain of an item's last tile under the next item's first, dV/dK epilogue of one item
//                           under the first MMAs of the next; D = 128: one item per CTA.  Everything is computed TRANSPOSED
//                           (rows = keys): S^T = K Q^T and dP^T = V dO^T so that P^T and dS^T land in
//                           TMEM exactly in the layout tcgen05 wants for an A operand (dV += P^T dO,
//                           dK += dS^T Q read A from TMEM); dS is also staged in shared memory for the
//                           dQ product.  dQ partial tiles leave through shared memory and a TMA bulk
//                           reduce-add into an fp32 accumulator - no per-element global atomics
//                           (reference: cu:1602-1610).
//   3. bwd_dq_finish_kernel: fp32 accumulator * scale -> 16-bit dq (+ l2norm backward of q); writes the accumulator
//                           back as zeros (the "zeroed" workspace is zero on entry and on exit of every call).
//
// The main kernel is bound by shared-memory bandwidth (MMA operand fetch, TMA fills and the element-wise
// warps' own traffic share 128 B/clk), so its structure minimises shared-memory accesses:
//   * 16 compute warps, four per scheduler: warpgroup g owns a quarter of the query columns of every tile
//     and runs the whole element-wise chain on them (S^T -> P^T -> dS^T -> drain of its dQ columns);
//     P^T stays in registers between the exp and the dS step.
//   * D = 64: the per-query constants are added INSIDE the accumulators by one extra K = 16 MMA step
//     (ones[128x16] x sliver[QT x 16]^T, SWIZZLE_32B operands) - no shared-memory loads in the
//     element-wise stage; K and V are copied into TMEM once per CTA, so S^T and dP^T are TS MMAs.
//   * TMEM aliasing without long dependency chains: P^T over the dP^T columns its thread has read,
//     dS^T into the dQ accumulator columns the same thread has just drained.
//
// Two shapes of the same kernel (TMEM has 512 columns; dV and dK need D each):
//   D = 64 : QT = 128.  S^T [0,128) dP^T/P^T [128,256) dV [256,320) dK [320,384) dQ/dS^T [384,448) K,V [448,512)
//            dQ = dS K      (M = queries, A = dS from smem M-major, B = K)
//   D = 128: QT = 64.   S^T [0,64)  dP^T/P^T [64,128)  dV [128,256) dK [256,384) dQ^T/dS^T [384,448)
//            dQ^T = K^T dS^T (M = features, A = K from smem M-major, B = dS^T); constants from shared memory
// BIAS instantiation: additive attention bias (16-bit loads along the query axis) and d_bias = dS
// (fp32 reductions), reference cu:1474-1476 / 1574-1576.
#pragma once

#include <stdlib.h>
#include <algorithm>
#include <type_traits>

// register split of the backward CTA (setmaxnreg): 4 compute warpgroups + 1 service warpgroup, 4 c + s = 5 x 96
#ifndef FCSA_BWD_SVC_REGS
#define FCSA_BWD_SVC_REGS 64
#define FCSA_BWD_CMP_REGS 104
#endif
#ifndef FCSA_BWD_POLY_EVERY
#define FCSA_BWD_POLY_EVERY 3   // backward exp stage: 1 of every N exp pairs runs on the FMA pipe (0 = none)
#endif

#include "../../include/fcsa_b200.h"
#include "l2norm_kernels.cuh"
#include "sm100_primitives.cuh"
#include "tensor_map.h"

namespace fcsa {

template <int D>
struct BwdCfg {
  static constexpr int QT = (D == 64) ? 128 : 64;        // query rows per tile
  static constexpr int kKV = 128 * D * 2;                 // bytes of the K (or V) tile
  static constexpr int kQ = QT * D * 2;                   // bytes of one Q (or dO) stage
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kKV;
  static constexpr int NST = 3;                           // Q / dO / stats ring depth
  static constexpr int kOffQ = kOffV + kKV;
  static constexpr int kOffDO = kOffQ + NST * kQ;
  static constexpr int kDS = 128 * QT * 2;                // dS^T staging: 128 keys x QT queries, 16-bit
  static constexpr int kOffDS = kOffDO + NST * kQ;
  static constexpr int kOffDQ = kOffDS + kDS;             // fp32 staging 128 x 64 = 32 KB
  static constexpr int kOffStats = kOffDQ + 32768;        // NST stages x 1 KB (2*QT floats used)
  // D = 64: the per-query constants reach the accumulators through one extra K = 16 MMA step
  // ("augmented" contraction) instead of through per-column shared-memory loads:
  //   ones   : 128 keys x 16, columns 0..2 = 1                      (A operand, 4 KB, once per CTA)
  //   aug S  : QT queries x 16, columns 0..2 = c3/c1 split in three 16-bit parts   (B operand)
  //   aug dP : QT queries x 16, columns 0..2 = -delta split in three 16-bit parts  (B operand)
  static constexpr bool kAug = (D == 64);
  static constexpr int kSliver = QT * 32;                 // bytes of one QT x 16 sliver
  static constexpr int kOffOnes = kOffStats + NST * 1024;
  static constexpr int kOffAug = kOffOnes + (kAug ? 4096 : 0);        // NST stages x {aug S, aug dP}
  static constexpr int kOffBar = kOffAug + (kAug ? NST * 2 * kSliver : 0);
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static_assert(kSmem <= 232448, "backward kernel shared memory");
  static constexpr int kThreads = 640;                  // 16 compute warps + 4 service warps
  // TMEM columns
  static constexpr uint32_t TM_S = 0, TM_DP = QT, TM_DV = 2 * QT, TM_DK = 2 * QT + D, TM_DQ = 2 * QT + 2 * D,
                            TM_X = 2 * QT + 2 * D + 64;   // D = 64: K and V as TMEM A operands (2 x 32 packed columns)
};

// ------------------------------------------------------------------------------------------
// Two caller-provided device buffers:
//
// SCRATCH workspace (contents irrelevant on entry, garbage on exit):
//   stats : fp32 [B*H][nqt][2][QT]   c3 then -delta for each query tile (padded rows = 0)
//   dkv_acc (kv_heads == 1 < heads only): fp32 dk [B][Nk][D] then dv [B][Nk][D]
//   aug   : 16-bit [B*H][nqt*QT][32]  cols 0..2 = c3/c1 in three parts, cols 16..18 = -delta in three
//           parts, rest 0 (the extra K = 16 step of S^T and dP^T, D = 64); ones: 16-bit [128][16]
//
// ZEROED workspace (all zero on entry - the caller zero-fills it ONCE, fcsa_zeroed_init - and all
// zero again on exit: the dq conversion pass clears every accumulator tile it converts, so no per-call
// memset / zeroing pass is needed; reference: a 33.5 MB cudaMemset of dq per backward, cu:1818):
//   dq_acc: fp32 [B*H][nqt][4 warps][16 chunks][32 lanes][4]   32 KB per query tile, in the order the
//           reduce warps produce it (D = 64: lane = query row, chunk = 4 features;
//           D = 128: lane = feature, chunk = 4 query rows)
// ------------------------------------------------------------------------------------------
struct BwdWorkspace {
  size_t stats_off, dkv_off, aug_off, ones_off, total;     // scratch
  size_t dq_off, ztotal;                                    // zeroed
  int nqt, QT;
};

inline BwdWorkspace bwd_workspace_layout(int B, int H, int kv_heads, int Nq, int Nk, int D) {
  BwdWorkspace w;
  w.QT = (D == 64) ? 128 : 64;
  w.nqt = (Nq + w.QT - 1) / w.QT;
  size_t stats = (size_t)B * H * w.nqt * 2 * w.QT * 4;
  size_t dq = (size_t)B * H * w.nqt * 32768;
  size_t dkv = (kv_heads == 1 && H > 1) ? (size_t)2 * B * Nk * D * 4 : 0;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  // augmented-contraction operands (16-bit): [B*H][nqt*QT][32] and the 128 x 16 ones tile
  size_t aug = (size_t)B * H * w.nqt * w.QT * 32 * 2;
  w.stats_off = 0;
  w.dkv_off = up(stats);
  w.aug_off = w.dkv_off + up(dkv);
  w.ones_off = w.aug_off + up(aug);
  w.total = w.ones_off + 4096;
  w.dq_off = 0;
  w.ztotal = up(dq);
  return w;
}

inline size_t bwd_workspace_bytes(int B, int H, int kv_heads, int Nq, int Nk, int D) {
  return bwd_workspace_layout(B, H, kv_heads, Nq, Nk, D).total;
}
inline size_t bwd_zeroed_workspace_bytes(int B, int H, int kv_heads, int Nq, int Nk, int D) {
  return bwd_workspace_layout(B, H, kv_heads, Nq, Nk, D).ztotal;
}

// ------------------------------------------------------------------------------------------
// 1. preprocess
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  int B, H, Nq, Nk, D, nqt, QT, causal;
  int bpb;                          // blocks per (batch, head): the grid is 1-D (no 65535 limit on batch*heads)
  float c2;                         // shift * log2e
  const float* shift_extra;         // optional device scalar: shift += max(*shift_extra, 0) (bias range guard)
  int o_f32;                        // 1: o is float32 (strides in float32 elements)
  const void* o;  long long o_sb, o_sh, o_sn;
  const void* d_o; long long do_sb, do_sh, do_sn;
  const float* inv_l;               // (B, H, Nq)
  float* stats;
  void* aug;                        // 16-bit [B*H][nqt*QT][32] (see workspace layout) or nullptr
  void* ones;                       // 16-bit [128][16]
  float inv_c1;                     // 1 / (scale * log2e)
};

// x = p0 + p1 + p2 with each part representable in T (16 bit): 24 bits of x survive
template <typename T>
__device__ __forceinline__ void split3(float x, uint32_t& w01, uint32_t& w2) {
  const float2 r0 = unpack2<T>(pack2<T>(x, 0.f));
  const float e1 = x - r0.x;
  const float2 r1 = unpack2<T>(pack2<T>(e1, 0.f));
  const float e2 = e1 - r1.x;
  w01 = pack2<T>(r0.x, r1.x);
  w2 = pack2<T>(e2, 0.f);
}

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TPR = threads per row = D / 8 (8 or 16); QT is 128 (D = 64) or 64 (D = 128).
template <typename T, int TPR>
__global__ void __launch_bounds__(256) bwd_prep_kernel(const PrepArgs a) {
  // 1-D grid = batch*heads x (row blocks of the padded sequence); TPR threads per row, two rows per
  // thread (four 16-byte loads in flight), shuffle reduce.  The pass must stay memory-bound (33.5 MB in, 8 MB
  // out at the metric shape): no runtime-divisor divisions, MUFU log2, the 64-byte sliver row of a query
  // written by four lanes instead of one.
  pdl_launch_dependents();
  pdl_wait();
  constexpr int RPB = 256 / TPR;             // rows per block and pass
  constexpr int QT = (TPR == 8) ? 128 : 64;
  constexpr int QSH = (TPR == 8) ? 7 : 6;
  const int bh = blockIdx.x / a.bpb;
  const int blk = blockIdx.x - bh * a.bpb;
  const int b = bh / a.H, h = bh - b * a.H;
  const int tr = threadIdx.x % TPR;
  const int padded = a.nqt * QT;
  const float c2 = a.c2 + (a.shift_extra != nullptr ? fmaxf(__ldg(a.shift_extra), 0.f) * 1.4426950408889634f : 0.f);
  const long long o_off = b * a.o_sb + h * a.o_sh + tr * 8;
  const T* obase = reinterpret_cast<const T*>(a.o) + o_off;
  const float* obase32 = reinterpret_cast<const float*>(a.o) + o_off;
  const T* dbase = reinterpret_cast<const T*>(a.d_o) + b * a.do_sb + h * a.do_sh + tr * 8;
  int row[2];
  bool in[2], valid[2];
  uint4 ro[2], rd[2];
  float4 rof[2][2];                           // o as float32 (o_f32 problems): 8 features = two float4
  float il[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    row[u] = (blk * 2 + u) * RPB + threadIdx.x / TPR;     // row inside the padded (nqt*QT) range
    in[u] = row[u] < padded;
    valid[u] = in[u] && row[u] < a.Nq;
    ro[u] = rd[u] = make_uint4(0, 0, 0, 0);
    rof[u][0] = rof[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    il[u] = 1.f;
    if (valid[u]) {
      if (a.o_f32) {
        rof[u][0] = __ldg(reinterpret_cast<const float4*>(obase32 + (long long)row[u] * a.o_sn));
        rof[u][1] = __ldg(reinterpret_cast<const float4*>(obase32 + (long long)row[u] * a.o_sn + 4));
      } else {
        ro[u] = ldg_stream128(obase + (long long)row[u] * a.o_sn);
      }
      rd[u] = ldg_stream128(dbase + (long long)row[u] * a.do_sn);
      if (tr == 0) il[u] = __ldg(a.inv_l + (long long)bh * a.Nq + row[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    float2 a0 = unpack2<T>(ro[u].x), a1 = unpack2<T>(ro[u].y), a2 = unpack2<T>(ro[u].z), a3 = unpack2<T>(ro[u].w);
    if (a.o_f32) {
      a0 = make_float2(rof[u][0].x, rof[u][0].y); a1 = make_float2(rof[u][0].z, rof[u][0].w);
      a2 = make_float2(rof[u][1].x, rof[u][1].y); a3 = make_float2(rof[u][1].z, rof[u][1].w);
    }
    const float2 b0 = unpack2<T>(rd[u].x), b1 = unpack2<T>(rd[u].y), b2 = unpack2<T>(rd[u].z), b3 = unpack2<T>(rd[u].w);
    float dot = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y +
                a3.x * b3.x + a3.y * b3.y;
#pragma unroll
    for (int m = 1; m < TPR; m <<= 1) dot += __shfl_xor_sync(0xFFFFFFFFu, dot, m);
    // lane tr == 0 of the row holds inv_l: c3 = log2(inv_l) - shift*log2e, broadcast to the row's lanes
    float c3 = valid[u] ? lg2_approx(il[u]) - c2 : 0.f;
    c3 = __shfl_sync(0xFFFFFFFFu, c3, (threadIdx.x & 31) & ~(TPR - 1));
    const float ndl = valid[u] ? -dot : 0.f;      // stored negated: the dS stage computes P * (dP + (-delta))
    if (!in[u]) continue;
    if (a.aug) {
      // D = 64: the 64-byte sliver row {c3/c1 in three 16-bit parts, 0.., -delta in three parts, 0..}; lanes 0..3
      // of the row write one 16-byte quarter each
      if (tr < 4) {
        uint4 q4 = make_uint4(0, 0, 0, 0);
        if (valid[u] && (tr & 1) == 0) split3<T>(tr == 0 ? c3 * a.inv_c1 : ndl, q4.x, q4.y);
        reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.aug) + ((long long)bh * padded + row[u]) * 64)[tr] = q4;
      }
    } else if (tr == 0) {
      // D = 128: fp32 {c3, -delta} per query tile, read by the main kernel through shared memory
      const int qt = row[u] >> QSH, r = row[u] & (QT - 1);
      float* st = a.stats + ((long long)bh * a.nqt + qt) * 2 * QT;
      st[r] = c3;
      st[QT + r] = ndl;
    }
  }
  if (a.aug && blockIdx.x == 0 && threadIdx.x < 128) {
    uint4* dst = reinterpret_cast<uint4*>(a.ones) + threadIdx.x * 2;
    const uint32_t one2 = pack2<T>(1.f, 1.f), one1 = pack2<T>(1.f, 0.f);
    dst[0] = make_uint4(one2, one1, 0, 0);
    dst[1] = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------
// 2. main kernel
// ------------------------------------------------------------------------------------------
struct BwdArgs {
  int B, H, Nq, Nk, kv_heads, causal, has_mask, nqt;
  float c1;                         // scale * log2e
  float scale;
  const uint8_t* mask; long long mask_sb;
  const float* stats;
  float* dq_acc;
  float* dk_acc; float* dv_acc;     // fp32 (B, Nk, D) accumulators when kv_heads == 1 < H
  void* dk; long long dk_sb, dk_sh, dk_sn;
  void* dv; long long dv_sb, dv_sh, dv_sn;
  // when set, dk is the gradient w.r.t. the RAW keys: the l2norm backward
  // (dk_hat - k_hat <k_hat, dk_hat>_group) * rnorm_group is applied in the epilogue
  const float* k_rnorm;             // (B, H, Nk, G) or nullptr
  int G;                            // groups (group size D/G >= 8 when k_rnorm is set)
  // additive bias on the logits (BIAS instantiation only): element type T, [b][h][query][key], element
  // strides (bias_sb = 0: no batch dimension), keys contiguous.  dbias: fp32 accumulator of dS in the same
  // index space ([.][.][Nq][Nk] planes, strides dbias_sb (0 = summed over the batch) / dbias_sh) or nullptr.
  const void* bias; long long bias_sb, bias_sh, bias_sn;
  float* dbias; long long dbias_sb, dbias_sh;
  int out_f32;                      // 1: dq, dk, dv are float32 tensors (strides in float32 elements)
  // dq conversion pass (D = 64, bwd_dq_finish64_kernel): fp32 accumulator tile -> 16-bit dq (x scale,
  // optional l2norm backward w.r.t. the raw q), tile cleared
  void* dq; long long dq_sb, dq_sh, dq_sn;
  const void* q_hat; long long q_sb, q_sh, q_sn;   // normalised q (only read when q_rnorm is set)
  const float* q_rnorm;             // (B, H, Nq, G) or nullptr
};

// 16-bit global load through the read-only path (bias elements along the query axis are strided)
__device__ __forceinline__ uint32_t ldg_u16(const void* p) {
  uint16_t v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// 16-byte load served by L2 (never a stale L1 line): accumulator tiles other CTAs have reduced into
__device__ __forceinline__ float4 ldg_cg128f(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// dq conversion, D = 64 (reference: the fp32 -> scalar_t cast of dq after its atomics, cu:1904).
// Accumulator tile = [4 row quarters][16 feature chunks][32 rows][4 features].  One warp converts 8 rows:
// four adjacent lanes share a row, 16 features each - every 16-byte load of a warp covers whole 128-byte
// lines, and group sums of the l2norm backward
//     dq_raw = (dq_hat - q_hat <q_hat, dq_hat>_group) * rnorm_group            (py:38-65 + autograd)
// are shuffles among those lanes.  The tile is read from L2 (the bulk reduce-adds of the main kernel have
// just produced it there), converted, and written back as zeros: the accumulator is zero between launches.
template <typename T>
__device__ __forceinline__ void finish_dq_tile64(const BwdArgs& a, int bh, int b, int h, int qt, int row, int part) {
  float* tile = a.dq_acc + ((long long)bh * a.nqt + qt) * 8192;
  const int wq = row >> 5, rl = row & 31;
  float* src = tile + ((wq * 16 + 4 * part) * 32 + rl) * 4;
  float4 acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = ldg_cg128f(src + c * 128);
  const int row_g = qt * 128 + row;
  const bool ok = row_g < a.Nq;
  const bool l2 = a.q_rnorm != nullptr;            // uniform across the grid
  uint4 qraw[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
  if (l2 && ok) {
    const T* qp = reinterpret_cast<const T*>(a.q_hat) + b * a.q_sb + h * a.q_sh + (long long)row_g * a.q_sn + part * 16;
    qraw[0] = ldg_stream128(qp);
    qraw[1] = ldg_stream128(qp + 8);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(src + c * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
  float g[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    g[4 * c] = acc[c].x * a.scale; g[4 * c + 1] = acc[c].y * a.scale;
    g[4 * c + 2] = acc[c].z * a.scale; g[4 * c + 3] = acc[c].w * a.scale;
  }
  if (l2) {
    float y[16];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const float2 u0 = unpack2<T>(qraw[s2].x), u1 = unpack2<T>(qraw[s2].y), u2 = unpack2<T>(qraw[s2].z),
                   u3 = unpack2<T>(qraw[s2].w);
      y[8 * s2] = u0.x; y[8 * s2 + 1] = u0.y; y[8 * s2 + 2] = u1.x; y[8 * s2 + 3] = u1.y;
      y[8 * s2 + 4] = u2.x; y[8 * s2 + 5] = u2.y; y[8 * s2 + 6] = u3.x; y[8 * s2 + 7] = u3.y;
    }
    const int gs = 64 / a.G;                       // features per group (power of two)
    const long long rbase = (((long long)b * a.H + h) * a.Nq + (ok ? row_g : 0)) * a.G;
    if (gs >= 8) {
      float d0 = 0.f, d1 = 0.f;                    // <q_hat, dq_hat> over this thread's two 8-feature segments
#pragma unroll
      for (int i = 0; i < 8; ++i) { d0 += y[i] * g[i]; d1 += y[8 + i] * g[8 + i]; }
      if (gs >= 16) { d0 += d1; d1 = d0; }
      if (gs >= 32) { d0 += __shfl_xor_sync(0xFFFFFFFFu, d0, 1); d1 = d0; }
      if (gs >= 64) { d0 += __shfl_xor_sync(0xFFFFFFFFu, d0, 2); d1 = d0; }
      const int seg0 = 2 * part;                   // 8-feature segment index of this thread's first half
      const float r0 = ok ? a.q_rnorm[rbase + (seg0 * 8) / gs] : 0.f;
      const float r1 = ok ? a.q_rnorm[rbase + (seg0 * 8 + 8) / gs] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g[i] = (g[i] - y[i] * d0) * r0;
        g[8 + i] = (g[8 + i] - y[8 + i] * d1) * r1;
      }
    } else {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        float pr[8], dot[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pr[i] = y[8 * s2 + i] * g[8 * s2 + i];
        subgroup_sums8(pr, gs, dot);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float rn = ok ? a.q_rnorm[rbase + (part * 16 + 8 * s2 + i) / gs] : 0.f;
          g[8 * s2 + i] = (g[8 * s2 + i] - y[8 * s2 + i] * dot[i]) * rn;
        }
      }
    }
  }
  if (ok && a.out_f32) {
    float* dst = reinterpret_cast<float*>(a.dq) + b * a.dq_sb + h * a.dq_sh + (long long)row_g * a.dq_sn + part * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(g[4 * c], g[4 * c + 1], g[4 * c + 2], g[4 * c + 3]);
  } else if (ok) {
    T* dst = reinterpret_cast<T*>(a.dq) + b * a.dq_sb + h * a.dq_sh + (long long)row_g * a.dq_sn + part * 16;
    uint4 o0, o1;
    o0.x = pack2<T>(g[0], g[1]);   o0.y = pack2<T>(g[2], g[3]);   o0.z = pack2<T>(g[4], g[5]);   o0.w = pack2<T>(g[6], g[7]);
    o1.x = pack2<T>(g[8], g[9]);   o1.y = pack2<T>(g[10], g[11]); o1.z = pack2<T>(g[12], g[13]); o1.w = pack2<T>(g[14], g[15]);
    *reinterpret_cast<uint4*>(dst) = o0;
    *reinterpret_cast<uint4*>(dst + 8) = o1;
  }
}

template <typename T, int D, bool BIAS = false>
__global__ void __launch_bounds__(640, 1)
fcsa_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_aug, const __grid_constant__ CUtensorMap tm_ones,
                const BwdArgs a) {
  static_assert(D == 64 || D == 128, "head dim 64 or 128");
  using Cfg = BwdCfg<D>;
  constexpr int QT = Cfg::QT;              // query rows per tile
  constexpr int KCH = D / 64;              // 64-feature chunks of every operand tile
  constexpr int QCHUNK = QT * 128;         // bytes of one 64-feature chunk of a Q / dO stage
  constexpr int NST = Cfg::NST;
  constexpr uint32_t TM_S = Cfg::TM_S, TM_DP = Cfg::TM_DP, TM_DV = Cfg::TM_DV, TM_DK = Cfg::TM_DK,
                     TM_DQ = Cfg::TM_DQ, TM_X = Cfg::TM_X;
  constexpr bool KV_IN_TMEM = (D == 64);
  // D = 128 runs one item per CTA (grid = number of items): every item loop below is then a single pass that the
  // compiler can see through (running tile index = tile index, no loop-carried item state in registers)
  constexpr bool PERSIST = KV_IN_TMEM;
  constexpr bool AUG = Cfg::kAug;          // per-query constants enter through an extra K = 16 MMA step

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  // two 128 x D operand buffers: item n keeps K in buffer (n & 1) and V in the other one.  D = 64: V only
  // passes through shared memory on its way to TMEM, so the next item's K is prefetched into it; the next
  // item's V follows into K's buffer when the last dQ MMA and the dK epilogue are done with it.
  const uint32_t sKV0 = smem_u32(smem + Cfg::kOffK);
  // (D = 128: one item per CTA, K and V stay where they are - and their descriptors stay compile-time offsets)
  auto buf_k = [&](int n) { return sKV0 + ((KV_IN_TMEM && (n & 1)) ? Cfg::kKV : 0); };
  auto buf_v = [&](int n) { return sKV0 + ((KV_IN_TMEM && (n & 1)) ? 0 : Cfg::kKV); };
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ);
  const uint32_t sDO = smem_u32(smem + Cfg::kOffDO);
  const uint32_t sDS = smem_u32(smem + Cfg::kOffDS);
  const uint32_t sDQ = smem_u32(smem + Cfg::kOffDQ);
  const uint32_t sStats = smem_u32(smem + Cfg::kOffStats);
  const uint32_t sOnes = smem_u32(smem + Cfg::kOffOnes);
  const uint32_t sAug = smem_u32(smem + Cfg::kOffAug);     // stage st: S sliver, then dP sliver
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  enum {
    K_FULL = 0, V_FULL = 1, Q_FULL = 2, Q_EMPTY = Q_FULL + NST, DO_FULL = Q_EMPTY + NST, DO_EMPTY = DO_FULL + NST,
    S_FULL = DO_EMPTY + NST, S_FREE, P_FULL, PV_DONE, DP_FULL, K_TMEM, V_TMEM, DS_FULL,
    DQ_FULL, DKV_FULL, DV_FREE, DK_FREE, KS_FREE, NBARS
  };
  static_assert(NBARS * 8 + 4 <= 256, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  pdl_launch_dependents();

  // ---- work items -------------------------------------------------------------------------
  // One item = one 128-key tile of one (batch, head) with all the query tiles that can see it.  Items are
  // numbered heaviest first (causal: the first key tiles are seen by the most queries) and dealt to the CTAs
  // in snake order.  A grid of one CTA per SM makes the CTA persistent (D = 64): TMEM, barriers and the
  // Q / dO ring live on across items, the stream of query tiles is continuous (the dQ tile of an item's last
  // query tile is drained under the next item's first one), the next item's K is prefetched, and the dK / dV
  // epilogue of one item overlaps the first MMAs of the next.  A grid of n_items CTAs = one item per CTA.
  const int bh_count = a.B * a.H;
  const int n_items = ((a.Nk + 127) >> 7) * bh_count;
  const int off = a.Nk - a.Nq;
  struct Item {
    int b, h, hk, bh, key0, i_lo, NI;
  };
  auto item_index = [&](int r) -> int {
    const int G = gridDim.x, c = blockIdx.x;
    const int idx = r * G + ((r & 1) ? (G - 1 - c) : c);
    return idx < n_items ? idx : -1;
  };
  auto load_item = [&](int idx) -> Item {
    Item it;
    const int jt = idx / bh_count;                  // key tile; ascending = heaviest first (causal)
    it.bh = idx - jt * bh_count;
    it.b = it.bh / a.H;
    it.h = it.bh - it.b * a.H;
    it.hk = (a.kv_heads == 1) ? 0 : it.h;
    it.key0 = jt * 128;
    it.i_lo = 0;
    if (a.causal) {
      // first query tile with a row that can see key0: QT*i + QT-1 + off >= key0
      const int x = it.key0 - off - (QT - 1);
      it.i_lo = x <= 0 ? 0 : (x + QT - 1) / QT;
    }
    it.NI = a.nqt - it.i_lo;                        // number of query tiles to visit (>= 1)
    return it;
  };

  // ---- setup ------------------------------------------------------------------------------
  if (warp == 16 && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(BAR(K_FULL), 1);
    mbar_init(BAR(V_FULL), 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(BAR(Q_FULL + s), 1);
      mbar_init(BAR(Q_EMPTY + s), 2);
      mbar_init(BAR(DO_FULL + s), 1);
      mbar_init(BAR(DO_EMPTY + s), 2);
    }
    mbar_init(BAR(S_FULL), 1);
    mbar_init(BAR(S_FREE), 512);
    mbar_init(BAR(P_FULL), 512);
    mbar_init(BAR(PV_DONE), 1);
    mbar_init(BAR(DP_FULL), 1);
    mbar_init(BAR(DS_FULL), 512);
    mbar_init(BAR(DQ_FULL), 1);
    mbar_init(BAR(K_TMEM), 128);
    mbar_init(BAR(V_TMEM), 128);
    mbar_init(BAR(DKV_FULL), 2);
    mbar_init(BAR(DV_FREE), 128);       // dV / dK accumulators read out by the epilogue warpgroups
    mbar_init(BAR(DK_FREE), 128);
    mbar_init(BAR(KS_FREE), 129);       // K's shared-memory buffer: last dQ MMA (commit) + the dK epilogue warps
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();          // slivers / stats come from the preprocess kernel; accumulators are zero on entry

  if (wg == 4) {
    reg_dealloc<FCSA_BWD_SVC_REGS>();
    if (warp == 19) {
      // =============================== K / V producer ==============================
      // (a thread of its own: its waits - V copied to TMEM, K's buffer released - must not hold up the ring)
      if (elect_one()) {
        for (int n = 0; PERSIST || n == 0; ++n) {
          const int idx = PERSIST ? item_index(n) : static_cast<int>(blockIdx.x);
          if (idx < 0) break;
          const Item it = load_item(idx);
          if (n > 0) {
            mbar_wait(BAR(V_TMEM), (n - 1) & 1);      // buf_k(n) held V(n-1): copied to TMEM
            mbar_wait(BAR(K_TMEM), (n - 1) & 1);      // K(n-1) has landed and K_FULL's only waiters are past it:
          }                                           // the barrier may be re-armed
          mbar_expect_tx(BAR(K_FULL), Cfg::kKV + ((AUG && n == 0) ? 4096 : 0));
#pragma unroll
          for (int ch = 0; ch < KCH; ++ch) tma_load_4d(buf_k(n) + ch * 16384, &tm_k, BAR(K_FULL), ch * 64, it.key0, it.hk, it.b);
          if constexpr (AUG) {
            if (n == 0) tma_load_4d(sOnes, &tm_ones, BAR(K_FULL), 0, 0, 0, 0);
          }
          if (n > 0) mbar_wait(BAR(KS_FREE), (n - 1) & 1);       // buf_v(n) held K(n-1): dQ MMAs and dK epilogue done
          mbar_expect_tx(BAR(V_FULL), Cfg::kKV);
#pragma unroll
          for (int ch = 0; ch < KCH; ++ch) tma_load_4d(buf_v(n) + ch * 16384, &tm_v, BAR(V_FULL), ch * 64, it.key0, it.hk, it.b);
        }
      }
    } else if (warp == 16) {
      // =============================== Q / dO ring producer =======================
      if (elect_one()) {
        int tq = 0;                                 // query tiles so far, over all items (ring position)
        for (int n = 0; PERSIST || n == 0; ++n) {
          const int idx = PERSIST ? item_index(n) : static_cast<int>(blockIdx.x);
          if (idx < 0) break;
          const Item it = load_item(idx);
          for (int i = 0; i < it.NI; ++i, ++tq) {
            const int st = tq % NST, qt = it.i_lo + i;
            const uint32_t par = ((tq / NST) & 1) ^ 1;
            mbar_wait(BAR(Q_EMPTY + st), par);
            if (n == 0) FCSA_TR(4, i, 0);
            mbar_expect_tx(BAR(Q_FULL + st), Cfg::kQ + (AUG ? Cfg::kSliver : 8 * QT));
#pragma unroll
            for (int ch = 0; ch < KCH; ++ch)
              tma_load_4d(sQ + st * Cfg::kQ + ch * QCHUNK, &tm_q, BAR(Q_FULL + st), ch * 64, qt * QT, it.h, it.b);
            if constexpr (AUG)
              tma_load_4d(sAug + st * 2 * Cfg::kSliver, &tm_aug, BAR(Q_FULL + st), 0, qt * QT, it.bh, 0);
            else
              bulk_load_1d(sStats + st * 1024,
                           a.stats + ((long long)it.bh * a.nqt + qt) * 2 * QT, 8 * QT, BAR(Q_FULL + st));
            mbar_wait(BAR(DO_EMPTY + st), par);
            if (n == 0) FCSA_TR(4, i, 1);
            mbar_expect_tx(BAR(DO_FULL + st), Cfg::kQ + (AUG ? Cfg::kSliver : 0));
#pragma unroll
            for (int ch = 0; ch < KCH; ++ch)
              tma_load_4d(sDO + st * Cfg::kQ + ch * QCHUNK, &tm_do, BAR(DO_FULL + st), ch * 64, qt * QT, it.h, it.b);
            if constexpr (AUG)
              tma_load_4d(sAug + st * 2 * Cfg::kSliver + Cfg::kSliver, &tm_aug, BAR(DO_FULL + st), 16, qt * QT, it.bh, 0);
          }
        }
      }
    } else {
      // =============================== MMA issuers ================================
      // Two issuing threads, one per dependency chain, so that neither waits behind the other's
      // barriers (the tensor pipe interleaves the two instruction streams):
      //   warp 17 (chain A): S^T(i+1) as soon as the compute warps hold S^T(i) in registers;
      //                      dV(i) += P^T(i) dO(i) when P^T(i) is in place; dP^T(i+1) right behind it
      //   warp 18 (chain B): dK(i) += dS^T(i) Q(i), then dQ(i)
      // Q / dO ring slots are read by both chains: their empty barriers count 2.
      if (elect_one()) {
        constexpr uint32_t idesc_s = umma_idesc<T>(128, QT, 0, 0);    // S^T, dP^T  (A, B K-major)
        constexpr uint32_t idesc_ts = umma_idesc<T>(128, D, 0, 1);    // dV, dK (A from TMEM, B MN-major)
        constexpr uint32_t idesc_dq = umma_idesc<T>(128, 64, 1, 1);   // dQ / dQ^T (A, B MN-major)
        // S^T = K Q^T (and dP^T = V dO^T): contraction over the D features, 16 per instruction
        auto issue_ST = [&](uint32_t d_col, uint32_t a_smem, uint32_t b_smem) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss(tmem + d_col,
                    umma_desc_sw128(a_smem + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(b_smem + (k >> 2) * QCHUNK + (k & 3) * 32, 16, 1024), idesc_s,
                    k > 0 ? 1u : 0u);
        };
        // packed dS^T of queries [16kk, 16kk+16) lives in the dQ accumulator columns: warpgroup g keeps
        // its QT/8 packed columns at offset 16g, i.e. inside the 16 accumulator columns it drains itself
        auto ds_col = [](int kk) -> uint32_t {
          constexpr int CW = QT / 4;
          const int w = (16 * kk) / CW;
          return static_cast<uint32_t>(w * 16 + (16 * kk - w * CW) / 2);
        };
        // P^T(i) (A operand of dV) is written, packed, over the dP^T columns its producer has consumed:
        // warpgroup g keeps its QT/8 packed columns at the start of its own QT/4 dP^T columns
        auto p_col = [](int kk) -> uint32_t {
          constexpr int CW = QT / 4;
          const int w = (16 * kk) / CW;
          return static_cast<uint32_t>(w * CW + (16 * kk - w * CW) / 2);
        };
        // D = 64: K and V also sit in TMEM (the X columns, written once per item by the compute warps),
        // so S^T and dP^T read only their B operand from shared memory
        auto issue_ST_ts = [&](uint32_t d_col, uint32_t a_col, uint32_t b_smem) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ts(tmem + d_col, tmem + a_col + k * 8,
                    umma_desc_sw128(b_smem + (k >> 2) * QCHUNK + (k & 3) * 32, 16, 1024), idesc_s,
                    k > 0 ? 1u : 0u);
        };
        int tq = 0;                                 // query tiles so far, over all items
        if (warp == 17) {
          // chain A: everything that does not depend on dS.
          for (int n = 0; PERSIST || n == 0; ++n) {
            const int idx = PERSIST ? item_index(n) : static_cast<int>(blockIdx.x);
            if (idx < 0) break;
            const Item it = load_item(idx);
            const int NI = it.NI;
            const uint32_t sK = buf_k(n), sV = buf_v(n);
            // st = ring stage of the Q / dO tile (and of its sliver)
            auto issue_S = [&](int st) {
              const uint32_t q_smem = sQ + st * Cfg::kQ;
              if constexpr (KV_IN_TMEM) issue_ST_ts(TM_S, TM_X, q_smem);
              else issue_ST(TM_S, sK, q_smem);
              if constexpr (AUG)      // S^T += ones * (c3/c1)^T : the exponent offset of every query column
                umma_ss(tmem + TM_S, umma_desc_sw32(sOnes), umma_desc_sw32(sAug + st * 2 * Cfg::kSliver), idesc_s, 1u);
            };
            auto issue_dP = [&](int st) {
              const uint32_t do_smem = sDO + st * Cfg::kQ;
              if constexpr (KV_IN_TMEM) issue_ST_ts(TM_DP, TM_X + 32, do_smem);
              else issue_ST(TM_DP, sV, do_smem);
              if constexpr (AUG)      // dP^T += ones * (-delta)^T
                umma_ss(tmem + TM_DP, umma_desc_sw32(sOnes),
                        umma_desc_sw32(sAug + st * 2 * Cfg::kSliver + Cfg::kSliver), idesc_s, 1u);
            };
            const int st0 = tq % NST;
            mbar_wait(BAR(KV_IN_TMEM ? K_TMEM : K_FULL), n & 1);
            if (tq > 0) mbar_wait(BAR(S_FREE), (tq - 1) & 1);     // the previous item's last S^T is in registers
            mbar_wait(BAR(Q_FULL + st0), (tq / NST) & 1);
            tc_fence_after();
            issue_S(st0);
            umma_commit(BAR(S_FULL));
            umma_commit(BAR(Q_EMPTY + st0));         // Q: this chain is done with it once S^T completes
            mbar_wait(BAR(KV_IN_TMEM ? V_TMEM : V_FULL), n & 1);
            mbar_wait(BAR(DO_FULL + st0), (tq / NST) & 1);
            tc_fence_after();
            issue_dP(st0);                           // (behind the previous item's last dV: same thread, in-order pipe)
            umma_commit(BAR(DP_FULL));
            umma_commit(BAR(DO_EMPTY + st0));
            for (int i = 0; i < NI; ++i) {
              const int t = tq + i;
              const int st = t % NST, sn = (t + 1) % NST;
              if (i + 1 < NI) {
                mbar_wait(BAR(S_FREE), t & 1);
                mbar_wait(BAR(Q_FULL + sn), ((t + 1) / NST) & 1);
                tc_fence_after();
                issue_S(sn);
                umma_commit(BAR(S_FULL));
                umma_commit(BAR(Q_EMPTY + sn));
                if (n == 0) FCSA_TR(0, i, 0);
              }
              mbar_wait(BAR(P_FULL), t & 1);
              if (i == 0 && n > 0) mbar_wait(BAR(DV_FREE), (n - 1) & 1);   // the previous item's dV has been read out
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < QT / 16; ++kk)
                umma_ts(tmem + TM_DV, tmem + TM_DP + p_col(kk),
                        umma_desc_sw128(sDO + st * Cfg::kQ + kk * 2048, QCHUNK, 1024), idesc_ts,
                        (i > 0 || kk > 0) ? 1u : 0u);
              umma_commit(BAR(PV_DONE));
              umma_commit(BAR(DO_EMPTY + st));       // dO(i): dV is done with it
              if (n == 0) FCSA_TR(0, i, 1);
              if (i + 1 < NI) {
                mbar_wait(BAR(DO_FULL + sn), ((t + 1) / NST) & 1);
                tc_fence_after();
                issue_dP(sn);
                umma_commit(BAR(DP_FULL));
                umma_commit(BAR(DO_EMPTY + sn));
                if (n == 0) FCSA_TR(0, i, 3);
              }
            }
            umma_commit(BAR(DKV_FULL));
            tq += NI;
          }
        } else {
          // chain B: the consumers of dS(i).  dK reads it from TMEM (the dQ accumulator columns),
          // then dQ(i) overwrites those columns - same issuing thread, in-order pipe.
          for (int n = 0; PERSIST || n == 0; ++n) {
            const int idx = PERSIST ? item_index(n) : static_cast<int>(blockIdx.x);
            if (idx < 0) break;
            const Item it = load_item(idx);
            const int NI = it.NI;
            const uint32_t sK = buf_k(n);
            mbar_wait(BAR(KV_IN_TMEM ? K_TMEM : K_FULL), n & 1);   // dQ reads K from shared memory (K_TMEM: it has landed)
            for (int i = 0; i < NI; ++i) {
              const int t = tq + i;
              const int st = t % NST;
              mbar_wait(BAR(DS_FULL), t & 1);
              if (i == 0 && n > 0) mbar_wait(BAR(DK_FREE), (n - 1) & 1);   // the previous item's dK has been read out
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < QT / 16; ++kk)
                umma_ts(tmem + TM_DK, tmem + TM_DQ + ds_col(kk),
                        umma_desc_sw128(sQ + st * Cfg::kQ + kk * 2048, QCHUNK, 1024), idesc_ts,
                        (i > 0 || kk > 0) ? 1u : 0u);
              umma_commit(BAR(Q_EMPTY + st));
              if (n == 0) FCSA_TR(0, i, 2);
#pragma unroll
#ifdef FCSA_EXP_NO_DQMMA
              if (false)
#endif
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t d_ds = umma_desc_sw128(sDS + kk * 2048, 16384, 1024);
                const uint64_t d_k = umma_desc_sw128(sK + kk * 2048, 16384, 1024);
                if (D == 64) umma_ss(tmem + TM_DQ, d_ds, d_k, idesc_dq, kk > 0 ? 1u : 0u);
                else umma_ss(tmem + TM_DQ, d_k, d_ds, idesc_dq, kk > 0 ? 1u : 0u);
              }
              umma_commit(BAR(DQ_FULL));
              if (n == 0) FCSA_TR(0, i, 4);
            }
            umma_commit(BAR(KS_FREE));                // the MMAs are done with K's shared-memory buffer
            umma_commit(BAR(DKV_FULL));
            tq += NI;
          }
        }
      }
    }
  } else {
    // =============================== compute warpgroups =============================
    // Four warpgroups run the SAME fused element-wise stage, each on one quarter of the query
    // columns of every tile (thread = key row, warpgroup g owns columns [g*QT/4, (g+1)*QT/4)):
    //   S^T -> P^T = exp2(S^T*c1 + c3)             -> X columns (packed 16 bit, A operand of dV)
    //   dP^T, P^T (still in registers) -> dS^T = P^T * (dP^T - delta)
    //                                              -> packed over the dP^T columns this warpgroup
    //                                                 has consumed (A operand of dK) and to smem
    //   dQ(i-1) partial tile: 16 accumulator columns -> smem -> TMA bulk reduce-add (2 KB per warp)
    // Splitting by columns (not by stage) puts four warps on every scheduler, all working on the
    // same tile: the TMEM / MUFU / shared-memory latencies of one warp hide under the others, and
    // P^T never makes a round trip through TMEM.
    reg_alloc<FCSA_BWD_CMP_REGS>();   // 4 x compute + service = 5 x 96: the pool is what the CTA was launched with (640 x 96)
    const int wq = warp & 3;
    const int r = wq * 32 + lane;            // key row inside the tile
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    constexpr int CW = QT / 4;               // query columns per warpgroup (32 or 16)
    const int cq0 = wg * CW;
    const uint32_t tS = lane_base + TM_S + cq0;
    const uint32_t tDP = lane_base + TM_DP + cq0;
    const uint32_t tP = lane_base + TM_DP + cq0;             // packed P^T goes over the dP^T columns this thread has read
    const uint32_t tDQ = lane_base + TM_DQ + 16 * wg;       // 16 dQ accumulator columns drained by this warpgroup
    const uint32_t tDS = tDQ;                                // ... which also hold its packed dS^T (CW/2 columns)
    uint32_t my_stage = sDQ + wq * 8192 + wg * 2048;           // 2 KB of dQ staging per warp
#ifndef FCSA_EXP_NO_PIN
    // Pin the per-thread addresses of the tile loop in registers: inside the item loop the compiler otherwise
    // re-derives them from tmem / the shared-memory base / the warp index in EVERY tile (rematerialisation),
    // +10 % instructions in the loop.  An empty asm with a "+r" operand makes the value opaque.
    uint32_t tS_p = tS, tDP_p = tDP, tDQ_p = tDQ;
    asm volatile("" : "+r"(tS_p), "+r"(tDP_p), "+r"(tDQ_p), "+r"(my_stage));
#define tS tS_p
#define tDP tDP_p
#define tP tDP_p
#define tDQ tDQ_p
#define tDS tDQ_p
#endif
    const bool tr_lane = (wg == 0 && wq == 0 && lane == 0);
    auto ld_cw = [&](uint32_t addr, uint32_t (&dst)[CW]) {
      if constexpr (CW == 32) tmem_ld_x32(addr, dst);
      else tmem_ld_x16(addr, dst);
    };
    auto st_cw = [&](uint32_t addr, const uint32_t (&src)[CW / 2]) {
      if constexpr (CW == 32) tmem_st_x16(addr, src);
      else tmem_st_x8(addr, src);
    };
    // dQ of tile j: this warp's 32 rows x 16 accumulator columns -> registers.  The same (lanes,
    // columns) receive this thread's packed dS^T of the next tile, so no other thread is involved.
    uint32_t dqv[16];
    // t = running index of the tile over all items (parity of its DQ_FULL phase)
    auto load_dq = [&](int t) {
      mbar_wait(BAR(DQ_FULL), t & 1);
      if (tr_lane) FCSA_TR(3, t, 0);
      tc_fence_after();
      tmem_ld_x16(tDQ, dqv);
      tmem_ld_wait();
      // the previous bulk reduce must have finished reading the staging buffer
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts128(my_stage + c * 512 + lane * 16, dqv[4 * c], dqv[4 * c + 1], dqv[4 * c + 2], dqv[4 * c + 3]);
    };
    // after a fence.proxy.async: the staged dQ tile -> global accumulator (TMA reduce-add, 2 KB per warp);
    // dst = this warp's 2 KB of the tile's accumulator
    // tile = index of the query tile's accumulator in the workspace (bh * nqt + qt)
    auto reduce_dq = [&](int tile) {
      __syncwarp();
#ifndef FCSA_EXP_SKIP_REDUCE
      if (lane == 0) {
        bulk_reduce_add_f32(a.dq_acc + ((long long)tile * 4 + wq) * 2048 + wg * 512, my_stage, 2048);
        bulk_commit_group();
      }
#endif
    };
    int prev_tile = -1;                      // accumulator tile of the previous query tile: drained one tile late
    int tq = 0;                              // query tiles so far, over all items
    const float c1 = a.c1;
    for (int n = 0; PERSIST || n == 0; ++n) {
    // (broadcast from lane 0: tells the compiler that everything derived from the item is warp-uniform, so the
    //  barrier waits of the tile loop stay on the uniform datapath, without reconvergence points)
    const int idx = PERSIST ? __shfl_sync(0xFFFFFFFFu, item_index(n), 0) : static_cast<int>(blockIdx.x);
    if (idx < 0) break;
    const Item it = load_item(idx);
    // (only what the tile loop needs stays live across it; the epilogue re-derives batch / head from idx)
    const int key0 = it.key0, i_lo = it.i_lo, NI = it.NI;
    const int tile0 = it.bh * a.nqt;         // accumulator tile index of query tile 0 of this (batch, head)
    [[maybe_unused]] const int b = it.b, h = it.h;   // tile loop: bias instantiation only
    const int key_g = key0 + r;
    FCSA_ITEM_T(tr_lane, idx, 0);
    const uint32_t sK = buf_k(n);
    if constexpr (KV_IN_TMEM) {
      // K (warpgroup 2) and V (warpgroup 3) rows -> TMEM as A operands: 64 features = 32 packed columns.
      // (Warpgroups 0 and 1 are the ones that come out of the previous item's epilogue.)  The previous item's
      // last S^T / dP^T MMAs - the readers of these columns - completed before these warps saw its last tile.
      if (wg >= 2) {
        mbar_wait(BAR(wg == 2 ? K_FULL : V_FULL), n & 1);
        const uint32_t src = wg == 2 ? sK : buf_v(n);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t kv[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 v4 = lds128f(src + sw128_offset(r, 4 * hf + c));
            kv[4 * c] = __float_as_uint(v4.x); kv[4 * c + 1] = __float_as_uint(v4.y);
            kv[4 * c + 2] = __float_as_uint(v4.z); kv[4 * c + 3] = __float_as_uint(v4.w);
          }
          tmem_st_x16(lane_base + TM_X + 32 * (wg - 2) + 16 * hf, kv);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(BAR(wg == 2 ? K_TMEM : V_TMEM));
      }
    }
    {
      bool key_ok = key_g < a.Nk;
      if (a.has_mask && key_ok) key_ok = a.mask[(long long)it.b * a.mask_sb + key_g] != 0;
      const bool tile_key_ragged = (key0 + 127 >= a.Nk) || a.has_mask;

      for (int i = 0; i < NI; ++i) {
        // running tile index: ring stage and barrier parities (the broadcast keeps it, and every wait below,
        // on the uniform datapath - a loop-carried sum over items is not recognised as warp-uniform)
        const int t = PERSIST ? __shfl_sync(0xFFFFFFFFu, tq + i, 0) : i;
        const int st = t % NST, qt = i_lo + i;
        const int row0 = qt * QT;
        const uint32_t c3a = sStats + st * 1024 + cq0 * 4;           // c3 of this warpgroup's queries
        const uint32_t dla = sStats + st * 1024 + (QT + cq0) * 4;    // -delta of the same
        // visible iff lo <= cc <= hi  (cc = query index inside the tile)
        const bool need_mask = tile_key_ragged || (row0 + QT - 1 >= a.Nq) ||
                               (a.causal && (key0 + 127 > row0 + off));
        int lo = 0, hi = QT - 1;
        if (need_mask) {
          hi = min(QT - 1, a.Nq - 1 - row0);
          if (a.causal) lo = max(0, key_g - off - row0);
          if (!key_ok) lo = 1000;
        }
        if (tr_lane) FCSA_TR(1, i, 0);
        mbar_wait(BAR(Q_FULL + st), (t / NST) & 1);    // c3 / delta of this tile are in smem
        mbar_wait(BAR(S_FULL), t & 1);
        if (tr_lane) FCSA_TR(1, i, 1);
        if (i == 0) FCSA_ITEM_T(tr_lane, idx, 2);
        tc_fence_after();
        // ---- exp stage.  The shared-memory operands are fetched in one batch so the exp chain
        // (FFMA -> MUFU -> pack) of CW independent elements can be pipelined freely.  The masked
        // variant is a separate instantiation: a per-element `if (need_mask)` compiles to a taken
        // branch per element pair and starves the warp of instructions.
        uint32_t pk[CW / 2];                 // P^T packed; lives until the dS stage below
        uint32_t s[CW];
        // bias[query][this key] of the CW queries, as packed pairs; the loads fly under the S^T wait
        uint32_t bw[BIAS ? CW / 2 : 1];
        if constexpr (BIAS) {
          const uint8_t* bcol = reinterpret_cast<const uint8_t*>(a.bias) +
                                2 * ((long long)b * a.bias_sb + (long long)h * a.bias_sh + min(key_g, a.Nk - 1));
#pragma unroll
          for (int e = 0; e < CW; e += 2) {
            const int q0 = min(row0 + cq0 + e, a.Nq - 1), q1 = min(row0 + cq0 + e + 1, a.Nq - 1);
            bw[e / 2] = ldg_u16(bcol + 2 * (long long)q0 * a.bias_sn) |
                        (ldg_u16(bcol + 2 * (long long)q1 * a.bias_sn) << 16);
          }
        }
        ld_cw(tS, s);
        float c3v[AUG ? 2 : CW];
        if constexpr (!AUG) {
#pragma unroll
          for (int e = 0; e < CW; e += 4) {
            const float4 k0 = lds128f(c3a + e * 4);
            c3v[e] = k0.x; c3v[e + 1] = k0.y; c3v[e + 2] = k0.z; c3v[e + 3] = k0.w;
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(BAR(S_FREE));                         // S^T(i+1) may be produced now
        if (tr_lane) FCSA_TR(1, i, 2);
        auto exp_tile = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
          for (int e = 0; e < CW; e += 2) {
            const float2 sv = make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1]));
            float2 x;
            if constexpr (AUG) x = __fmul2_rn(sv, make_float2(c1, c1));     // the accumulator already holds s + c3/c1
            else x = __ffma2_rn(sv, make_float2(c1, c1), make_float2(c3v[AUG ? 0 : e], c3v[AUG ? 1 : e + 1]));
            if constexpr (BIAS)
              x = __ffma2_rn(unpack2<T>(bw[BIAS ? e / 2 : 0]), make_float2(1.4426950408889634f, 1.4426950408889634f), x);
            // some of the pairs on the FMA pipe (cubic minimax exp2), the rest on the MUFU
            const bool poly = FCSA_BWD_POLY_EVERY > 0 && ((e / 2) % (FCSA_BWD_POLY_EVERY > 0 ? FCSA_BWD_POLY_EVERY : 1)) == FCSA_BWD_POLY_EVERY - 1;
#ifdef FCSA_EXP_NO_EXP
            const float2 pe = x;
#else
            const float2 pe = poly ? ex2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
#endif
            float p0 = pe.x;
            float p1 = pe.y;
            if (MASKED) {
              const int cc = cq0 + e;
              p0 = (cc >= lo && cc <= hi) ? p0 : 0.f;
              p1 = (cc + 1 >= lo && cc + 1 <= hi) ? p1 : 0.f;
            }
            pk[e / 2] = pack2<T>(p0, p1);
          }
        };
        if (need_mask) exp_tile(std::true_type{});
        else exp_tile(std::false_type{});
        if (tr_lane) FCSA_TR(1, i, 3);
        // ---- dP^T(i) into registers first: P^T(i) is then written over the columns it occupied
        // (DP_FULL(i) also tells that dV(i-1) has finished reading P^T(i-1): same in-order pipe)
        float dlv[AUG ? 2 : CW];
        if constexpr (!AUG) {
#pragma unroll
          for (int e = 0; e < CW; e += 4) {
            const float4 dl = lds128f(dla + e * 4);
            dlv[e] = dl.x; dlv[e + 1] = dl.y; dlv[e + 2] = dl.z; dlv[e + 3] = dl.w;
          }
        }
        mbar_wait(BAR(DP_FULL), t & 1);
        if (tr_lane) FCSA_TR(2, i, 0);
        tc_fence_after();
        uint32_t (&d)[CW] = s;                // the S^T registers are dead: reuse them for dP^T
        ld_cw(tDP, d);
        tmem_ld_wait();
        st_cw(tP, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(BAR(P_FULL));
        if (tr_lane) FCSA_TR(1, i, 4);
        if (tr_lane) FCSA_TR(2, i, 1);
        uint32_t ds[CW / 2];
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          const float2 pa = unpack2<T>(pk[e / 2]);
          float2 t = make_float2(__uint_as_float(d[e]), __uint_as_float(d[e + 1]));   // AUG: already dP - delta
          if constexpr (!AUG) t = __fadd2_rn(t, make_float2(dlv[AUG ? 0 : e], dlv[AUG ? 1 : e + 1]));  // delta stored negated
          const float2 v = __fmul2_rn(pa, t);
          ds[e / 2] = pack2<T>(v.x, v.y);
          if constexpr (BIAS) {
            // d bias = dS (reference cu:1574-1576): fp32 reduction, lanes = consecutive keys of one query row
            if (a.dbias != nullptr && key_g < a.Nk) {
              float* dst = a.dbias + (long long)b * a.dbias_sb + (long long)h * a.dbias_sh + key_g;
              const int q0 = row0 + cq0 + e;
              if (q0 < a.Nq) red_add_f32(dst + (long long)q0 * a.Nk, v.x);
              if (q0 + 1 < a.Nq) red_add_f32(dst + (long long)(q0 + 1) * a.Nk, v.y);
            }
          }
        }
        // ---- dQ(i-1) out of its accumulator columns (its MMA sits right behind dK(i-1): long done;
        // it has also released the shared-memory dS^T), then dS^T(i) into them
        if (t > 0) load_dq(t - 1);            // (t is warp-uniform by construction: no reconvergence code)
        if (tr_lane) FCSA_TR(2, i, 2);
        st_cw(tDS, ds);
        // the same CW queries -> shared memory: row = key, query-contiguous 64-wide chunks, 128B swizzle
#ifndef FCSA_EXP_NO_STS
#pragma unroll
        for (int q4 = 0; q4 < CW / 8; ++q4)
          sts128(sDS + (cq0 >> 6) * 16384 + sw128_offset(r, ((cq0 & 63) >> 3) + q4), ds[4 * q4],
                 ds[4 * q4 + 1], ds[4 * q4 + 2], ds[4 * q4 + 3]);
#endif
        tmem_st_wait();
        fence_proxy_async_smem();             // covers the dS^T tile and the staged dQ(i-1)
        tc_fence_before();
        mbar_arrive(BAR(DS_FULL));
        if (tr_lane) FCSA_TR(2, i, 3);
        if (i == 0) FCSA_ITEM_T(tr_lane, idx, 3);
        if (i == NI - 1) FCSA_ITEM_T(tr_lane, idx, 4);
        if (t > 0) reduce_dq(prev_tile);
        prev_tile = tile0 + qt;
      }
      tq += NI;
    }

    // ---- epilogue: warpgroup 0 stores dV, warpgroup 1 stores dK * scale -------------------
    if (wg < 2) {
    const int w = wg;
    int idx_e = idx;
    asm volatile("" : "+r"(idx_e));            // opaque copy: batch / head are recomputed here, not kept in registers
    const Item ie = load_item(idx_e);
    const int b = ie.b, h = ie.h, hk = ie.hk;
    if (NI > 0) {
      mbar_wait(BAR(DKV_FULL), n & 1);
      tc_fence_after();
    }
    const bool store_ok = key_g < a.Nk;
    // The stores below are staged through the chunks of the dS^T tile that THIS warp writes in the main loop
    // (32 key rows x its warpgroup's 64 bytes of query columns): the last dQ MMA - the tile's only reader -
    // completed before DKV_FULL, and the next writer is this warp itself.
    // (D = 128, one item per CTA, 64-query tiles: a warp's own chunks are only 32 bytes wide - warpgroups 0 / 1
    //  take chunk columns 0-3 / 4-7 of their key rows; the other warps wrote them last before the final dS hand-over)
    const uint32_t st_tile = sDS + (D == 64 ? (cq0 >> 6) * 16384 : 0);
    const int st_c0 = (D == 64) ? ((cq0 & 63) >> 3) : 4 * wg;
    const uint32_t tACC = lane_base + (w == 0 ? TM_DV : TM_DK);
    const float mul = (w == 0) ? 1.0f : a.scale;
    const bool shared_kv = (a.kv_heads == 1 && a.H > 1);
    const bool fuse_l2 = (w == 1) && (a.k_rnorm != nullptr) && !shared_kv && NI > 0;
    float gd[D / 8];                          // <k_hat, dk_hat> of the group each 8-feature segment is in
    float rk[D / 8];
    if (fuse_l2) {
      float sd[D / 8];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t acc[32];
        tmem_ld_x32(tACC + c * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int seg = 4 * c + s4;
          const float4 kw = lds128f(sK + (seg >> 3) * 16384 + sw128_offset(r, seg & 7));
          const float2 k0 = unpack2<T>(__float_as_uint(kw.x)), k1 = unpack2<T>(__float_as_uint(kw.y)),
                       k2 = unpack2<T>(__float_as_uint(kw.z)), k3 = unpack2<T>(__float_as_uint(kw.w));
          sd[seg] = __uint_as_float(acc[8 * s4 + 0]) * k0.x + __uint_as_float(acc[8 * s4 + 1]) * k0.y +
                    __uint_as_float(acc[8 * s4 + 2]) * k1.x + __uint_as_float(acc[8 * s4 + 3]) * k1.y +
                    __uint_as_float(acc[8 * s4 + 4]) * k2.x + __uint_as_float(acc[8 * s4 + 5]) * k2.y +
                    __uint_as_float(acc[8 * s4 + 6]) * k3.x + __uint_as_float(acc[8 * s4 + 7]) * k3.y;
        }
      }
      const int segs_per_group = (D / a.G) >> 3;            // power of two
      int lg = 0;
      while ((1 << lg) < segs_per_group) ++lg;
      const long long rbase = (((long long)b * a.H + h) * a.Nk + (store_ok ? key_g : 0)) * a.G;
#pragma unroll
      for (int s0 = 0; s0 < D / 8; ++s0) {
        float t = 0.f;
#pragma unroll
        for (int s1 = 0; s1 < D / 8; ++s1) t += ((s1 >> lg) == (s0 >> lg)) ? sd[s1] : 0.f;
        gd[s0] = t;
        rk[s0] = a.k_rnorm[rbase + (s0 >> lg)];
      }
    }
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t acc[32];
      if (NI > 0) {
        tmem_ld_x32(tACC + c * 32, acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x) acc[x] = 0;
      }
      if (c == D / 32 - 1) {                   // the accumulator is in registers: the next item may overwrite it
        tc_fence_before();
        mbar_arrive(BAR(w == 0 ? DV_FREE : DK_FREE));
      }
      if (fuse_l2) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int seg = 4 * c + s4;
          const float4 kw = lds128f(sK + (seg >> 3) * 16384 + sw128_offset(r, seg & 7));
          const uint32_t kk[4] = {__float_as_uint(kw.x), __float_as_uint(kw.y), __float_as_uint(kw.z),
                                  __float_as_uint(kw.w)};
          const float m2 = rk[seg];              // mul (= scale) is applied below
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 kf = unpack2<T>(kk[e2]);
            const float a0 = (__uint_as_float(acc[8 * s4 + 2 * e2]) - kf.x * gd[seg]) * m2;
            const float a1 = (__uint_as_float(acc[8 * s4 + 2 * e2 + 1]) - kf.y * gd[seg]) * m2;
            acc[8 * s4 + 2 * e2] = __float_as_uint(a0);
            acc[8 * s4 + 2 * e2 + 1] = __float_as_uint(a1);
          }
        }
      }
      if (!shared_kv) {
        // this warp's 32 key rows x 32 features, coalesced through shared memory (see warp_store_rows64)
        const int valid = min(32, max(0, a.Nk - (key0 + wq * 32)));
        const long long row_off = (long long)(key0 + wq * 32);
        uint32_t w16[16];
        if (a.out_f32) {
          float* base = ((w == 0) ? reinterpret_cast<float*>(a.dv) + b * a.dv_sb + hk * a.dv_sh + row_off * a.dv_sn
                                  : reinterpret_cast<float*>(a.dk) + b * a.dk_sb + hk * a.dk_sh + row_off * a.dk_sn) + c * 32;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int x = 0; x < 16; ++x) w16[x] = __float_as_uint(__uint_as_float(acc[16 * hf + x]) * mul);
            warp_store_rows64_sw128(st_tile, wq * 32, st_c0, lane, w16, reinterpret_cast<uint8_t*>(base + 16 * hf),
                                    ((w == 0) ? a.dv_sn : a.dk_sn) * 4, valid);
          }
        } else {
          T* base = ((w == 0) ? reinterpret_cast<T*>(a.dv) + b * a.dv_sb + hk * a.dv_sh + row_off * a.dv_sn
                              : reinterpret_cast<T*>(a.dk) + b * a.dk_sb + hk * a.dk_sh + row_off * a.dk_sn) + c * 32;
#pragma unroll
          for (int x = 0; x < 16; ++x)
            w16[x] = pack2<T>(__uint_as_float(acc[2 * x]) * mul, __uint_as_float(acc[2 * x + 1]) * mul);
          warp_store_rows64_sw128(st_tile, wq * 32, st_c0, lane, w16, reinterpret_cast<uint8_t*>(base),
                                  ((w == 0) ? a.dv_sn : a.dk_sn) * 2, valid);
        }
      } else if (store_ok) {
        // keys/values shared by all heads: sum over heads in fp32 (reference: cu:1613-1619)
        float* accp = ((w == 0) ? a.dv_acc : a.dk_acc) + ((long long)b * a.Nk + key_g) * D + c * 32;
#pragma unroll
        for (int x = 0; x < 32; ++x) atomicAdd(accp + x, __uint_as_float(acc[x]) * mul);
      }
    }
    if (w == 1) mbar_arrive(BAR(KS_FREE));     // this thread's reads of K's shared-memory buffer are done
    FCSA_ITEM_T(tr_lane, idx, 6);
    }
    }   // items
    // the last query tile of the last item: its dQ tile is still in the accumulator columns
    if (prev_tile >= 0) {
      load_dq(tq - 1);
      fence_proxy_async_smem();
      reduce_dq(prev_tile);
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }

#ifndef FCSA_EXP_NO_PIN
#undef tS
#undef tDP
#undef tP
#undef tDQ
#undef tDS
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// 3. finish kernels (D = 128 dq; shared-kv dk/dv)
// ------------------------------------------------------------------------------------------
struct DqFinishArgs {
  int B, H, Nq, D, nqt;
  float scale;
  int out_f32;                      // 1: dq is float32
  float* dq_acc;
  void* dq; long long sb, sh, sn;
  // when q_rnorm is set, dq is the gradient w.r.t. the RAW queries (l2norm backward applied here)
  const void* q_hat; long long q_sb, q_sh, q_sn;
  const float* q_rnorm;             // (B, H, Nq, G) or nullptr
  int G;
};

// (dq_hat - q_hat <q_hat, dq_hat>_group) * rnorm_group on the 8 features one thread owns; the
// `tpr` threads of a row are consecutive lanes, groups are aligned sub-blocks of them.
template <typename T>
__device__ __forceinline__ void finish_l2norm_bwd(float (&g)[8], const DqFinishArgs& a, bool ok, uint4 raw, int b,
                                                  int h, int row, int c8) {
  if (a.q_rnorm == nullptr) return;            // uniform across the grid
  float y[8];
  {
    const float2 u0 = unpack2<T>(raw.x), u1 = unpack2<T>(raw.y), u2 = unpack2<T>(raw.z), u3 = unpack2<T>(raw.w);
    y[0] = u0.x; y[1] = u0.y; y[2] = u1.x; y[3] = u1.y; y[4] = u2.x; y[5] = u2.y; y[6] = u3.x; y[7] = u3.y;
  }
  const int gs = a.D / a.G;
  const long long rbase = (((long long)b * a.H + h) * a.Nq + row) * a.G;
  if (gs >= 8) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dot += y[i] * g[i];
    const int tpg = gs >> 3;
    for (int m = 1; m < tpg; m <<= 1) dot += __shfl_xor_sync(0xFFFFFFFFu, dot, m);
    const float rn = ok ? a.q_rnorm[rbase + c8 / tpg] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = (g[i] - y[i] * dot) * rn;
  } else {
    float pr[8], dot[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pr[i] = y[i] * g[i];
    subgroup_sums8(pr, gs, dot);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float rn = ok ? a.q_rnorm[rbase + (c8 * 8 + i) / gs] : 0.f;
      g[i] = (g[i] - y[i] * dot[i]) * rn;
    }
  }
}

// D = 64: one block of 512 threads = one query tile of 128 rows (16 warps x 8 rows).
template <typename T>
__global__ void __launch_bounds__(512) bwd_dq_finish64_kernel(const BwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long unit = blockIdx.x;                       // (bh, qt)
  const int qt = (int)(unit % a.nqt);
  const int bh = (int)(unit / a.nqt);
  const int b = bh / a.H, h = bh - b * a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  finish_dq_tile64<T>(a, bh, b, h, qt, warp * 8 + (lane >> 2), lane & 3);
}

// D = 128: accumulator tile (64 query rows) = [4 warps][16 row-chunks][32 features][4 rows], i.e.
// transposed.  One block = one tile: coalesced float4 loads -> shared memory -> row-major stores; the
// tile is written back as zeros (the accumulator is zero between launches, see the workspace layout).

template <typename T>
__global__ void __launch_bounds__(256) bwd_dq_finish128_kernel(const DqFinishArgs a) {
  __shared__ float tile[64][129];
  pdl_launch_dependents();
  pdl_wait();
  const long long unit = blockIdx.x;                       // (bh, qt)
  const int qt = (int)(unit % a.nqt);
  const int bh = (int)(unit / a.nqt);
  const int b = bh / a.H, h = bh % a.H;
  float* src = a.dq_acc + unit * 8192;
  // 2048 float4 per tile = 8 per thread: all eight loads first (L2, the bulk reduce-adds have just produced the
  // tile there), then the zeros - a store between two loads of the same array would serialise them
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ldg_cg128f(src + (threadIdx.x + 256 * i) * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = threadIdx.x + 256 * i;
    *reinterpret_cast<float4*>(src + f * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    const int wq = f >> 9, c = (f >> 5) & 15, ln = f & 31;
    const int d = wq * 32 + ln, r0 = c * 4;
    tile[r0 + 0][d] = v[i].x;
    tile[r0 + 1][d] = v[i].y;
    tile[r0 + 2][d] = v[i].z;
    tile[r0 + 3][d] = v[i].w;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {      // 16 x 8 features per row
    const int rr = e >> 4, c8 = e & 15;
    const int row = qt * 64 + rr;
    const bool ok = row < a.Nq;
    const float* s = &tile[rr][c8 * 8];
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = s[i] * a.scale;
    uint4 qraw = make_uint4(0, 0, 0, 0);
    if (ok && a.q_rnorm != nullptr)
      qraw = ldg_stream128(reinterpret_cast<const T*>(a.q_hat) + b * a.q_sb + h * a.q_sh + (long long)row * a.q_sn +
                           c8 * 8);
    finish_l2norm_bwd<T>(g, a, ok, qraw, b, h, ok ? row : 0, c8);
    if (ok && a.out_f32) {
      float* dst = reinterpret_cast<float*>(a.dq) + b * a.sb + h * a.sh + (long long)row * a.sn + c8 * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(g[0], g[1], g[2], g[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(g[4], g[5], g[6], g[7]);
    } else if (ok) {
      uint4 o4;
      o4.x = pack2<T>(g[0], g[1]);
      o4.y = pack2<T>(g[2], g[3]);
      o4.z = pack2<T>(g[4], g[5]);
      o4.w = pack2<T>(g[6], g[7]);
      T* dst = reinterpret_cast<T*>(a.dq) + b * a.sb + h * a.sh + (long long)row * a.sn + c8 * 8;
      *reinterpret_cast<uint4*>(dst) = o4;
    }
  }
}

struct KvFinishArgs {
  int B, Nk, D, out_f32;
  const float* acc;                 // (B, Nk, D) fp32
  void* out; long long sb, sn;
};
template <typename T>
__global__ void __launch_bounds__(256) bwd_kv_finish_kernel(const KvFinishArgs a) {
  const int tpr = a.D >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)a.B * a.Nk * tpr;
  if (idx >= total) return;
  const int c8 = (int)(idx % tpr);
  const long long rowid = idx / tpr;
  const int n = (int)(rowid % a.Nk);
  const int b = (int)(rowid / a.Nk);
  const float* src = a.acc + rowid * a.D + c8 * 8;
  const float4 lo = *reinterpret_cast<const float4*>(src);
  const float4 hi = *reinterpret_cast<const float4*>(src + 4);
  if (a.out_f32) {
    float* dst = reinterpret_cast<float*>(a.out) + b * a.sb + (long long)n * a.sn + c8 * 8;
    *reinterpret_cast<float4*>(dst) = lo;
    *reinterpret_cast<float4*>(dst + 4) = hi;
    return;
  }
  uint4 o4;
  o4.x = pack2<T>(lo.x, lo.y);
  o4.y = pack2<T>(lo.z, lo.w);
  o4.z = pack2<T>(hi.x, hi.y);
  o4.w = pack2<T>(hi.z, hi.w);
  T* dst = reinterpret_cast<T*>(a.out) + b * a.sb + (long long)n * a.sn + c8 * 8;
  *reinterpret_cast<uint4*>(dst) = o4;
}

// ------------------------------------------------------------------------------------------
// host runner
// ------------------------------------------------------------------------------------------
struct BwdHostArgs {
  bool dtype_bf16;
  int B, H, kv_heads, Nq, Nk, D, causal;
  float scale, shift;
  const uint8_t* mask; long long mask_sb;
  fcsa_tensor q, k, v, o, d_o, dq, dk, dv;
  const float* inv_l;
  bool out_f32 = false;             // o (read), dq, dk, dv (written) are float32
  void* workspace;                  // scratch
  void* zeroed;                     // zero on entry, zero again on exit (dq accumulator, tile counters)
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;   // optional: recorded around the main kernel
  cudaEvent_t ev_prep[2] = {nullptr, nullptr}, ev_finish[2] = {nullptr, nullptr};   // ... the preprocess / dq finish
  // fused l2norm backward (q, k above are then the NORMALISED tensors and dq, dk the gradients
  // w.r.t. the raw ones); both null = plain backward
  const float* q_rnorm = nullptr;
  const float* k_rnorm = nullptr;
  int groups = 1;
  // additive bias (nullptr = none) and its fp32 gradient accumulator (nullptr = not needed)
  const void* bias = nullptr; long long bias_sb = 0, bias_sh = 0, bias_sn = 0;
  const float* bias_amax = nullptr;           // optional device scalar added (if positive) to the shift
  float* dbias = nullptr; long long dbias_sb = 0, dbias_sh = 0;
};

template <typename T, int D, bool BIAS = false>
int run_backward_t(const BwdHostArgs& h, cudaStream_t stream, int* launches, const char** err,
                   cudaError_t* ce) {
  using Cfg = BwdCfg<D>;
  const BwdWorkspace w = bwd_workspace_layout(h.B, h.H, h.kv_heads, h.Nq, h.Nk, D);
  uint8_t* ws = reinterpret_cast<uint8_t*>(h.workspace);
  uint8_t* zs = reinterpret_cast<uint8_t*>(h.zeroed);
  float* stats = reinterpret_cast<float*>(ws + w.stats_off);
  float* dq_acc = reinterpret_cast<float*>(zs + w.dq_off);
  float* dkv_acc = reinterpret_cast<float*>(ws + w.dkv_off);
  const bool shared_kv = (h.kv_heads == 1 && h.H > 1);
  const float log2e = 1.4426950408889634f;
  cudaError_t e;
  BwdArgs main_args;                  // filled in step 2; the D = 64 finish kernel reads its dq / q_hat fields

  // dk/dv accumulators of shared keys/values are per call (the dq accumulator is self-cleaning)
  if (shared_kv) {
    e = cudaMemsetAsync(dkv_acc, 0, (size_t)2 * h.B * h.Nk * D * 4, stream);
    if (e != cudaSuccess) { *err = "cudaMemsetAsync(workspace)"; *ce = e; return FCSA_ERR_CUDA; }
  }
  const int gs = D / (h.groups > 0 ? h.groups : 1);
  const bool fuse_k = h.k_rnorm != nullptr && !shared_kv && gs >= 8;

  // 1. preprocess
  {
    PrepArgs pa;
    pa.B = h.B; pa.H = h.H; pa.Nq = h.Nq; pa.Nk = h.Nk; pa.D = D; pa.nqt = w.nqt; pa.QT = Cfg::QT; pa.causal = h.causal;
    pa.c2 = h.shift * log2e;
    pa.shift_extra = h.bias_amax;
    pa.o_f32 = h.out_f32 ? 1 : 0;
    pa.o = h.o.ptr; pa.o_sb = h.o.sb; pa.o_sh = h.o.sh; pa.o_sn = h.o.sn;
    pa.d_o = h.d_o.ptr; pa.do_sb = h.d_o.sb; pa.do_sh = h.d_o.sh; pa.do_sn = h.d_o.sn;
    pa.inv_l = h.inv_l; pa.stats = stats;
    pa.aug = Cfg::kAug ? ws + w.aug_off : nullptr;
    pa.ones = ws + w.ones_off;
    pa.inv_c1 = 1.0f / (h.scale * log2e);
    const int rows_per_block = 2 * (256 / (D / 8));
    const int padded = w.nqt * Cfg::QT;
    pa.bpb = (padded + rows_per_block - 1) / rows_per_block;
    const long long grid = (long long)pa.bpb * h.B * h.H;
    if (grid > 0x7FFFFFFFLL) { *err = "problem too large for one launch"; return FCSA_ERR_INVALID; }
    if (h.ev_prep[0]) cudaEventRecord(h.ev_prep[0], stream);
    e = launch_pdl(bwd_prep_kernel<T, D / 8>, dim3((unsigned)grid), dim3(256), 0, stream, pa);
    if (h.ev_prep[1]) cudaEventRecord(h.ev_prep[1], stream);
    if (e != cudaSuccess) { *err = "backward preprocess launch"; *ce = e; return FCSA_ERR_CUDA; }
    ++*launches;
  }
  // 2. main
  {
    CUtensorMap tq, tk, tv, tdo;
    if (make_tensor_map_bhnd(&tq, h.q.ptr, h.dtype_bf16, h.B, h.H, h.Nq, D, h.q.sb, h.q.sh, h.q.sn, Cfg::QT) ||
        make_tensor_map_bhnd(&tk, h.k.ptr, h.dtype_bf16, h.B, h.kv_heads, h.Nk, D, h.k.sb, h.k.sh, h.k.sn, 128) ||
        make_tensor_map_bhnd(&tv, h.v.ptr, h.dtype_bf16, h.B, h.kv_heads, h.Nk, D, h.v.sb, h.v.sh, h.v.sn, 128) ||
        make_tensor_map_bhnd(&tdo, h.d_o.ptr, h.dtype_bf16, h.B, h.H, h.Nq, D, h.d_o.sb, h.d_o.sh, h.d_o.sn, Cfg::QT)) {
      *err = "cuTensorMapEncodeTiled failed (backward): q, k, v, d_o must be views a TMA tensor map can express "
             "(positive strides, 16-byte aligned)";
      return FCSA_ERR_INVALID;
    }
    // slivers of the augmented contraction: 16 columns x QT rows out of the [B*H][nqt*QT][32] tensor,
    // and the constant 128 x 16 ones tile (32-byte rows: SWIZZLE_32B).  D = 128 does not use them.
    CUtensorMap taug = tq, tones = tq;
    if (Cfg::kAug) {
      const long long padded = (long long)w.nqt * Cfg::QT;
      if (make_tensor_map_bhnd(&taug, ws + w.aug_off, h.dtype_bf16, 1, (long long)h.B * h.H, padded, 32,
                               (long long)h.B * h.H * padded * 32, padded * 32, 32, Cfg::QT, 16, 32) ||
          make_tensor_map_bhnd(&tones, ws + w.ones_off, h.dtype_bf16, 1, 1, 128, 16, 128 * 16, 128 * 16, 16, 128,
                               16, 32)) {
        *err = "cuTensorMapEncodeTiled failed (backward, slivers)";
        return FCSA_ERR_INVALID;
      }
    }
    BwdArgs& a = main_args;
    a.B = h.B; a.H = h.H; a.Nq = h.Nq; a.Nk = h.Nk; a.kv_heads = h.kv_heads; a.causal = h.causal;
    a.has_mask = h.mask ? 1 : 0; a.nqt = w.nqt;
    a.c1 = h.scale * log2e; a.scale = h.scale;
    a.mask = h.mask; a.mask_sb = h.mask_sb;
    a.stats = stats; a.dq_acc = dq_acc;
    a.dk_acc = dkv_acc; a.dv_acc = dkv_acc + (size_t)h.B * h.Nk * D;
    a.dk = h.dk.ptr; a.dk_sb = h.dk.sb; a.dk_sh = h.dk.sh; a.dk_sn = h.dk.sn;
    a.dv = h.dv.ptr; a.dv_sb = h.dv.sb; a.dv_sh = h.dv.sh; a.dv_sn = h.dv.sn;
    a.k_rnorm = fuse_k ? h.k_rnorm : nullptr;
    a.G = h.groups;
    a.bias = h.bias; a.bias_sb = h.bias_sb; a.bias_sh = h.bias_sh; a.bias_sn = h.bias_sn;
    a.dbias = h.dbias; a.dbias_sb = h.dbias_sb; a.dbias_sh = h.dbias_sh;
    a.dq = h.dq.ptr; a.dq_sb = h.dq.sb; a.dq_sh = h.dq.sh; a.dq_sn = h.dq.sn;
    a.q_hat = h.q.ptr; a.q_sb = h.q.sb; a.q_sh = h.q.sh; a.q_sn = h.q.sn;
    a.q_rnorm = h.q_rnorm;
    a.out_f32 = h.out_f32 ? 1 : 0;
    auto kern = fcsa_bwd_kernel<T, D, BIAS>;
    e = ensure_dynamic_smem<fcsa_bwd_kernel<T, D, BIAS>>(Cfg::kSmem);
    if (e != cudaSuccess) { *err = "cudaFuncSetAttribute(bwd)"; *ce = e; return FCSA_ERR_CUDA; }
    // D = 64: persistent CTAs, one per SM, each walking its share of the (key tile, batch, head) items;
    // D = 128 (K and V stay in shared memory for the whole item): one CTA per item.
    // FCSA_BWD_GRID (tuning / A-B knob): number of CTAs; 0 = one CTA per item
    const long long items = (long long)((h.Nk + 127) / 128) * h.B * h.H;
    static const long long grid_env = [] { const char* ev = getenv("FCSA_BWD_GRID"); return ev ? atoll(ev) : -1LL; }();
    long long grid = items;
    if (D == 64 && grid_env != 0) grid = std::min<long long>(items, grid_env > 0 ? grid_env : device_sm_count());
    if (h.ev_start) cudaEventRecord(h.ev_start, stream);
    e = launch_pdl(kern, dim3((unsigned)grid), dim3(Cfg::kThreads), Cfg::kSmem, stream, tq, tk, tv, tdo, taug, tones, a);
    if (h.ev_stop) cudaEventRecord(h.ev_stop, stream);
    if (e != cudaSuccess) { *err = "backward kernel launch"; *ce = e; return FCSA_ERR_CUDA; }
    ++*launches;
  }
  // 3. finish
  {
    DqFinishArgs fa;
    fa.B = h.B; fa.H = h.H; fa.Nq = h.Nq; fa.D = D; fa.nqt = w.nqt; fa.scale = h.scale;
    fa.dq_acc = dq_acc; fa.dq = h.dq.ptr; fa.sb = h.dq.sb; fa.sh = h.dq.sh; fa.sn = h.dq.sn;
    fa.q_hat = h.q.ptr; fa.q_sb = h.q.sb; fa.q_sh = h.q.sh; fa.q_sn = h.q.sn;
    fa.q_rnorm = h.q_rnorm; fa.G = h.groups; fa.out_f32 = h.out_f32 ? 1 : 0;
    const long long tiles = (long long)h.B * h.H * w.nqt;
    if (tiles > 0x7FFFFFFFLL) { *err = "problem too large for one launch"; return FCSA_ERR_INVALID; }
    if (h.ev_finish[0]) cudaEventRecord(h.ev_finish[0], stream);
    if (D == 64) e = launch_pdl(bwd_dq_finish64_kernel<T>, dim3((unsigned)tiles), dim3(512), 0, stream, main_args);
    else e = launch_pdl(bwd_dq_finish128_kernel<T>, dim3((unsigned)tiles), dim3(256), 0, stream, fa);
    if (h.ev_finish[1]) cudaEventRecord(h.ev_finish[1], stream);
    if (e != cudaSuccess) { *err = "dq finish launch"; *ce = e; return FCSA_ERR_CUDA; }
    ++*launches;
    if (shared_kv) {
      for (int which = 0; which < 2; ++which) {
        KvFinishArgs ka;
        ka.B = h.B; ka.Nk = h.Nk; ka.D = D; ka.out_f32 = h.out_f32 ? 1 : 0;
        ka.acc = dkv_acc + (which ? (size_t)h.B * h.Nk * D : 0);
        const fcsa_tensor& t = which ? h.dv : h.dk;
        ka.out = t.ptr; ka.sb = t.sb; ka.sn = t.sn;
        const long long tot = (long long)h.B * h.Nk * (D / 8);
        bwd_kv_finish_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(ka);
        e = cudaGetLastError();
        if (e != cudaSuccess) { *err = "dk/dv finish launch"; *ce = e; return FCSA_ERR_CUDA; }
        ++*launches;
      }
    }
    if (h.k_rnorm != nullptr && !fuse_k) {
      // l2norm backward of dk as its own pass (shared keys/values, or groups finer than 8 features)
      L2Args la;
      la.B = h.B; la.H = h.kv_heads; la.N = h.Nk; la.D = D; la.G = h.groups;
      la.x_sb = h.dk.sb; la.x_sh = h.dk.sh; la.x_sn = h.dk.sn;
      la.y_sb = h.k.sb; la.y_sh = h.k.sh; la.y_sn = h.k.sn;
      la.o_sb = h.dk.sb; la.o_sh = h.dk.sh; la.o_sn = h.dk.sn;
      la.x = h.dk.ptr; la.y = h.k.ptr; la.dx = h.dk.ptr; la.rnorm = const_cast<float*>(h.k_rnorm);
      const int rows_per_block = 256 / (D / 8);
      const long long rows = (long long)la.B * la.H * la.N;
      l2norm_bwd_kernel<T><<<(unsigned)((rows + rows_per_block - 1) / rows_per_block), 256, 0, stream>>>(la);
      e = cudaGetLastError();
      if (e != cudaSuccess) { *err = "l2norm backward (dk) launch"; *ce = e; return FCSA_ERR_CUDA; }
      ++*launches;
    }
  }
  return FCSA_OK;
}

inline int run_backward(const BwdHostArgs& h, cudaStream_t stream, int* launches, const char** err,
                        cudaError_t* ce) {
  if (h.bias != nullptr) {
    if (h.D == 64) {
      if (h.dtype_bf16) return run_backward_t<__nv_bfloat16, 64, true>(h, stream, launches, err, ce);
      return run_backward_t<__half, 64, true>(h, stream, launches, err, ce);
    }
    if (h.dtype_bf16) return run_backward_t<__nv_bfloat16, 128, true>(h, stream, launches, err, ce);
    return run_backward_t<__half, 128, true>(h, stream, launches, err, ce);
  }
  if (h.D == 64) {
    if (h.dtype_bf16) return run_backward_t<__nv_bfloat16, 64>(h, stream, launches, err, ce);
    return run_backward_t<__half, 64>(h, stream, launches, err, ce);
  }
  if (h.dtype_bf16) return run_backward_t<__nv_bfloat16, 128>(h, stream, launches, err, ce);
  return run_backward_t<__half, 128>(h, stream, launches, err, ce);
}

}  // namespace fcsa
