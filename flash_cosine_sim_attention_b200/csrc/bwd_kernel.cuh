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
//                           rounded to 16 bit as the reference does at cu:1260/1820)
//   2. fcsa_bwd_kernel    : one CTA per (key tile of 128, batch, head), key/value tile stationary
//                           in shared memory, loop over 128-row query tiles.  Everything is computed
//                           TRANSPOSED (rows = keys): S^T = K Q^T and dP^T = V dO^T so that P^T and
//                           dS^T land in TMEM exactly in the layout tcgen05 wants for an A operand
//                           (dV += P^T dO, dK += dS^T Q read A from TMEM); dS is also staged in
//                           shared memory (M-major) for dQ = dS K.  dQ partial tiles leave through
//                           shared memory and a TMA bulk reduce-add into an fp32 accumulator -
//                           no per-element global atomics (reference: cu:1602-1610).
//   3. bwd_dq_finish_kernel: fp32 accumulator * scale -> 16-bit dq.
//
// TMEM columns (D = 64): S^T [0,128) dP^T [128,256) dV [256,320) dK [320,384) dQ [384,448).
// P^T / dS^T (packed 16-bit) overwrite the half of S^T / dP^T that the same warpgroup has just
// read, so no extra columns and no cross-warpgroup hazards.
#pragma once

#include "../../include/fcsa_b200.h"
#include "sm100_primitives.cuh"
#include "tensor_map.h"

namespace fcsa {

// ------------------------------------------------------------------------------------------
// workspace layout (all fp32):
//   stats : [B*H][nqt][2][128]   c3 then delta for each 128-row query tile (padded rows = 0)
//   dq_acc: [B*H][nqt][4 warps][D/4 chunks][32 rows][4]   (only what the dq finish kernel reads)
//   dkv_acc (kv_heads == 1 only): dk [B][Nk][D] then dv [B][Nk][D]
// ------------------------------------------------------------------------------------------
struct BwdWorkspace {
  size_t stats_off, dq_off, dkv_off, total;
  int nqt;
};

inline BwdWorkspace bwd_workspace_layout(int B, int H, int kv_heads, int Nq, int Nk, int D) {
  BwdWorkspace w;
  w.nqt = (Nq + 127) / 128;
  size_t stats = (size_t)B * H * w.nqt * 256 * 4;
  size_t dq = (size_t)B * H * w.nqt * 128 * D * 4;
  size_t dkv = (kv_heads == 1 && H > 1) ? (size_t)2 * B * Nk * D * 4 : 0;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  w.stats_off = 0;
  w.dq_off = up(stats);
  w.dkv_off = w.dq_off + up(dq);
  w.total = w.dkv_off + up(dkv);
  return w;
}

inline size_t bwd_workspace_bytes(int B, int H, int kv_heads, int Nq, int Nk, int D) {
  return bwd_workspace_layout(B, H, kv_heads, Nq, Nk, D).total;
}

// ------------------------------------------------------------------------------------------
// 1. preprocess
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  int B, H, Nq, D, nqt;
  float c2;                         // shift * log2e
  const void* o;  long long o_sb, o_sh, o_sn;
  const void* d_o; long long do_sb, do_sh, do_sn;
  const float* inv_l;               // (B, H, Nq)
  float* stats;
};

template <typename T>
__global__ void __launch_bounds__(256) bwd_prep_kernel(const PrepArgs a) {
  // D/8 threads per row, 16-byte loads, shuffle reduce
  const int tpr = a.D >> 3;
  const int rows_per_block = 256 / tpr;
  const long long prow = (long long)blockIdx.x * rows_per_block + threadIdx.x / tpr;  // padded row id
  const int tr = threadIdx.x % tpr;
  const long long padded = (long long)a.nqt * 128;
  const long long total = (long long)a.B * a.H * padded;
  const bool in = prow < total;
  const long long pr = in ? prow : 0;
  const int bh = (int)(pr / padded);
  const int row = (int)(pr % padded);
  const int b = bh / a.H, h = bh % a.H;
  const bool valid = in && row < a.Nq;
  float dot = 0.f;
  if (valid) {
    const T* op = reinterpret_cast<const T*>(a.o) + b * a.o_sb + h * a.o_sh + (long long)row * a.o_sn + tr * 8;
    const T* dp = reinterpret_cast<const T*>(a.d_o) + b * a.do_sb + h * a.do_sh + (long long)row * a.do_sn + tr * 8;
    const uint4 ro = *reinterpret_cast<const uint4*>(op);
    const uint4 rd = *reinterpret_cast<const uint4*>(dp);
    const float2 a0 = unpack2<T>(ro.x), a1 = unpack2<T>(ro.y), a2 = unpack2<T>(ro.z), a3 = unpack2<T>(ro.w);
    const float2 b0 = unpack2<T>(rd.x), b1 = unpack2<T>(rd.y), b2 = unpack2<T>(rd.z), b3 = unpack2<T>(rd.w);
    dot = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y +
          a3.x * b3.x + a3.y * b3.y;
  }
  for (int m = 1; m < tpr; m <<= 1) dot += __shfl_xor_sync(0xFFFFFFFFu, dot, m);
  if (in && tr == 0) {
    const int qt = row >> 7, r = row & 127;
    float* st = a.stats + ((long long)bh * a.nqt + qt) * 256;
    float c3 = 0.f, dl = 0.f;
    if (valid) {
      c3 = log2f(a.inv_l[(long long)bh * a.Nq + row]) - a.c2;
      dl = dot;
    }
    st[r] = c3;
    st[128 + r] = dl;
  }
}

// ------------------------------------------------------------------------------------------
// 2. main kernel (D = 64)
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
};

template <int D>
struct BwdCfg {
  static constexpr int kTile = 128 * D * 2;       // 16 KB at D = 64
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kTile;
  static constexpr int kOffQ = kOffV + kTile;      // 2 stages
  static constexpr int kOffDO = kOffQ + 2 * kTile; // 2 stages
  static constexpr int kOffDS = kOffDO + 2 * kTile;   // 128 keys x 128 queries 16-bit = 32 KB
  static constexpr int kOffDQ = kOffDS + 32768;       // fp32 staging 128 x D
  static constexpr int kOffStats = kOffDQ + 128 * D * 4;  // 2 stages x 1 KB
  static constexpr int kOffBar = kOffStats + 2048;
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static constexpr int kThreads = 512;
};

template <typename T, int D>
__global__ void __launch_bounds__(512, 1)
fcsa_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const BwdArgs a) {
  static_assert(D == 64, "this kernel is the D = 64 variant");
  using Cfg = BwdCfg<D>;
  constexpr int TILE = Cfg::kTile;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const uint32_t sK = smem_u32(smem + Cfg::kOffK);
  const uint32_t sV = smem_u32(smem + Cfg::kOffV);
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ);
  const uint32_t sDO = smem_u32(smem + Cfg::kOffDO);
  uint8_t* pDS = smem + Cfg::kOffDS;
  const uint32_t sDS = smem_u32(pDS);
  uint8_t* pDQ = smem + Cfg::kOffDQ;
  const float* pStats = reinterpret_cast<const float*>(smem + Cfg::kOffStats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  enum {
    KV_FULL = 0, Q_FULL = 1, Q_EMPTY = 3, DO_FULL = 5, DO_EMPTY = 7, S_FULL = 9, P_FULL = 10,
    DP_FULL = 11, DS_FULL = 12, DS_FREE = 13, DQ_FULL = 14, DQ_EMPTY = 15, DKV_FULL = 16, NBARS = 17
  };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2;

  // ---- work item --------------------------------------------------------------------------
  const int bh_count = a.B * a.H;
  const int jt = blockIdx.x / bh_count;                 // key tile; ascending = heaviest first (causal)
  const int bh = blockIdx.x - jt * bh_count;
  const int b = bh / a.H, h = bh - b * a.H;
  const int hk = (a.kv_heads == 1) ? 0 : h;
  const int off = a.Nk - a.Nq;
  const int key0 = jt * 128;
  int i_lo = 0;
  if (a.causal) {
    const int x = key0 - off - 127;                     // first query row that can see key0 ... (tile granularity)
    i_lo = x <= 0 ? 0 : (x + 127) >> 7;
  }
  const int NI = a.nqt - i_lo;                          // number of query tiles to visit (>= 1)

  // ---- setup ------------------------------------------------------------------------------
  if (warp == 12 && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(BAR(KV_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(Q_FULL + s), 1);
      mbar_init(BAR(Q_EMPTY + s), 1);
      mbar_init(BAR(DO_FULL + s), 1);
      mbar_init(BAR(DO_EMPTY + s), 1);
    }
    mbar_init(BAR(S_FULL), 1);
    mbar_init(BAR(P_FULL), 256);
    mbar_init(BAR(DP_FULL), 1);
    mbar_init(BAR(DS_FULL), 256);
    mbar_init(BAR(DS_FREE), 1);
    mbar_init(BAR(DQ_FULL), 1);
    mbar_init(BAR(DQ_EMPTY), 128);
    mbar_init(BAR(DKV_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

  if (wg == 3) {
    reg_dealloc<64>();
    if (warp == 12) {
      // =============================== TMA producer ===============================
      if (NI > 0 && elect_one()) {
        mbar_expect_tx(BAR(KV_FULL), 2 * TILE);
        tma_load_4d(sK, &tm_k, BAR(KV_FULL), 0, key0, hk, b);
        tma_load_4d(sV, &tm_v, BAR(KV_FULL), 0, key0, hk, b);
        for (int i = 0; i < NI; ++i) {
          const int st = i & 1, qt = i_lo + i;
          const uint32_t par = ((i >> 1) & 1) ^ 1;
          mbar_wait(BAR(Q_EMPTY + st), par);
          mbar_expect_tx(BAR(Q_FULL + st), TILE + 1024);
          tma_load_4d(sQ + st * TILE, &tm_q, BAR(Q_FULL + st), 0, qt * 128, h, b);
          bulk_load_1d(smem_u32(smem + Cfg::kOffStats) + st * 1024,
                       a.stats + ((long long)bh * a.nqt + qt) * 256, 1024, BAR(Q_FULL + st));
          mbar_wait(BAR(DO_EMPTY + st), par);
          mbar_expect_tx(BAR(DO_FULL + st), TILE);
          tma_load_4d(sDO + st * TILE, &tm_do, BAR(DO_FULL + st), 0, qt * 128, h, b);
        }
      }
    } else if (warp == 13) {
      // =============================== MMA issuer =================================
      if (NI > 0 && elect_one()) {
        constexpr uint32_t idesc_s = umma_idesc<T>(128, 128, 0, 0);   // S^T, dP^T
        constexpr uint32_t idesc_ts = umma_idesc<T>(128, D, 0, 1);    // dV, dK (A from TMEM, B MN-major)
        constexpr uint32_t idesc_dq = umma_idesc<T>(128, D, 1, 1);    // dQ (A, B MN-major)
        auto issue_ST = [&](uint32_t d_col, uint32_t a_smem, uint32_t b_smem) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss(tmem + d_col, umma_desc_sw128(a_smem + k * 32, 16, 1024),
                    umma_desc_sw128(b_smem + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        };
        // A operand halves written by compute warpgroup 0 (queries 0..63) and 1 (64..127)
        auto a_col = [](uint32_t base, int kk) { return base + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8); };

        mbar_wait(BAR(KV_FULL), 0);
        mbar_wait(BAR(Q_FULL + 0), 0);
        tc_fence_after();
        issue_ST(TM_S, sK, sQ);
        umma_commit(BAR(S_FULL));
        mbar_wait(BAR(DO_FULL + 0), 0);
        tc_fence_after();
        issue_ST(TM_DP, sV, sDO);
        umma_commit(BAR(DP_FULL));

        for (int i = 0; i < NI; ++i) {
          const int st = i & 1;
          const bool more = (i + 1 < NI);
          // ---- dV += P^T dO ----
          mbar_wait(BAR(P_FULL), i & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(tmem + TM_DV, tmem + a_col(TM_S, kk),
                    umma_desc_sw128(sDO + st * TILE + kk * 2048, 16384, 1024), idesc_ts,
                    (i > 0 || kk > 0) ? 1u : 0u);
          umma_commit(BAR(DO_EMPTY + st));
          // ---- S^T(i+1) ----
          if (more) {
            mbar_wait(BAR(Q_FULL + (st ^ 1)), ((i + 1) >> 1) & 1);
            tc_fence_after();
            issue_ST(TM_S, sK, sQ + (st ^ 1) * TILE);
            umma_commit(BAR(S_FULL));
          }
          // ---- dK += dS^T Q ----
          mbar_wait(BAR(DS_FULL), i & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(tmem + TM_DK, tmem + a_col(TM_DP, kk),
                    umma_desc_sw128(sQ + st * TILE + kk * 2048, 16384, 1024), idesc_ts,
                    (i > 0 || kk > 0) ? 1u : 0u);
          umma_commit(BAR(Q_EMPTY + st));
          // ---- dQ = dS K ----
          mbar_wait(BAR(DQ_EMPTY), (i & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tmem + TM_DQ, umma_desc_sw128(sDS + kk * 2048, 16384, 1024),
                    umma_desc_sw128(sK + kk * 2048, 16384, 1024), idesc_dq, kk > 0 ? 1u : 0u);
          umma_commit(BAR(DQ_FULL));
          umma_commit(BAR(DS_FREE));
          // ---- dP^T(i+1) ----
          if (more) {
            mbar_wait(BAR(DO_FULL + (st ^ 1)), ((i + 1) >> 1) & 1);
            tc_fence_after();
            issue_ST(TM_DP, sV, sDO + (st ^ 1) * TILE);
            umma_commit(BAR(DP_FULL));
          }
        }
        umma_commit(BAR(DKV_FULL));
      }
    }
  } else if (wg == 2) {
    // =============================== dQ reduce warpgroup ============================
    reg_dealloc<96>();
    const int wq = warp & 3;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    uint8_t* my_stage = pDQ + wq * (32 * D * 4);                // 8 KB per warp
    for (int i = 0; i < NI; ++i) {
      const int qt = i_lo + i;
      mbar_wait(BAR(DQ_FULL), i & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_x32(lane_base + TM_DQ, r0);
      tmem_ld_x32(lane_base + TM_DQ + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(BAR(DQ_EMPTY));
      // the previous bulk reduce must have finished reading the staging buffer
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(my_stage + c * 512 + lane * 16) =
            make_uint4(r0[4 * c], r0[4 * c + 1], r0[4 * c + 2], r0[4 * c + 3]);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(my_stage + (8 + c) * 512 + lane * 16) =
            make_uint4(r1[4 * c], r1[4 * c + 1], r1[4 * c + 2], r1[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        float* dst = a.dq_acc + (((long long)bh * a.nqt + qt) * 4 + wq) * (32 * D);
        bulk_reduce_add_f32(dst, smem_u32(my_stage), 32 * D * 4);
        bulk_commit_group();
      }
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  } else {
    // =============================== compute warpgroups =============================
    reg_alloc<176>();
    const int w = wg;                        // 0: queries [0,64) of each tile, 1: [64,128)
    const int wq = warp & 3;
    const int r = wq * 32 + lane;            // key row inside the tile
    const int key_g = key0 + r;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t tS = lane_base + TM_S + 64 * w;
    const uint32_t tDP = lane_base + TM_DP + 64 * w;
    const float c1 = a.c1;
    bool key_ok = key_g < a.Nk;
    if (a.has_mask && key_ok) key_ok = a.mask[(long long)b * a.mask_sb + key_g] != 0;
    const bool tile_key_ragged = (key0 + 127 >= a.Nk) || a.has_mask;

    for (int i = 0; i < NI; ++i) {
      const int st = i & 1, qt = i_lo + i;
      const int row0 = qt * 128;
      const float* c3p = pStats + st * 256 + 64 * w;
      const float* dlp = c3p + 128;
      // visible iff lo <= cc <= hi  (cc = query index inside the tile)
      const bool need_mask = tile_key_ragged || (row0 + 127 >= a.Nq) ||
                             (a.causal && (key0 + 127 > row0 + off));
      int lo = 0, hi = 127;
      if (need_mask) {
        hi = min(127, a.Nq - 1 - row0);
        if (a.causal) lo = max(0, key_g - off - row0);
        if (!key_ok) lo = 1000;
      }

      mbar_wait(BAR(Q_FULL + st), (i >> 1) & 1);     // stats for this tile are in smem
      mbar_wait(BAR(S_FULL), i & 1);
      tc_fence_after();
      float p[64];
      {
        uint32_t s0[32], s1[32];
        tmem_ld_x32(tS, s0);
        tmem_ld_x32(tS + 32, s1);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 k0 = *reinterpret_cast<const float4*>(c3p + c);
          const float4 k1 = *reinterpret_cast<const float4*>(c3p + 32 + c);
          p[c + 0] = ex2_approx(fmaf(__uint_as_float(s0[c + 0]), c1, k0.x));
          p[c + 1] = ex2_approx(fmaf(__uint_as_float(s0[c + 1]), c1, k0.y));
          p[c + 2] = ex2_approx(fmaf(__uint_as_float(s0[c + 2]), c1, k0.z));
          p[c + 3] = ex2_approx(fmaf(__uint_as_float(s0[c + 3]), c1, k0.w));
          p[32 + c + 0] = ex2_approx(fmaf(__uint_as_float(s1[c + 0]), c1, k1.x));
          p[32 + c + 1] = ex2_approx(fmaf(__uint_as_float(s1[c + 1]), c1, k1.y));
          p[32 + c + 2] = ex2_approx(fmaf(__uint_as_float(s1[c + 2]), c1, k1.z));
          p[32 + c + 3] = ex2_approx(fmaf(__uint_as_float(s1[c + 3]), c1, k1.w));
        }
      }
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const int cc = 64 * w + c;
          p[c] = (cc >= lo && cc <= hi) ? p[c] : 0.f;
        }
      }
      // P^T -> TMEM (packed), over the S^T columns this warpgroup has just consumed
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t pk[16];
#pragma unroll
        for (int q2 = 0; q2 < 16; ++q2) pk[q2] = pack2<T>(p[32 * hh + 2 * q2], p[32 * hh + 2 * q2 + 1]);
        tmem_st_x16(tS + 16 * hh, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(BAR(P_FULL));

      // ---- dS = P * (dP - delta) ----
      mbar_wait(BAR(DP_FULL), i & 1);
      tc_fence_after();
      mbar_wait(BAR(DS_FREE), (i & 1) ^ 1);          // dQ(i-1) has finished reading the smem dS
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t d[32];
        tmem_ld_x32(tDP + 32 * hh, d);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 dl = *reinterpret_cast<const float4*>(dlp + 32 * hh + c);
          const float d0 = p[32 * hh + c + 0] * (__uint_as_float(d[c + 0]) - dl.x);
          const float d1 = p[32 * hh + c + 1] * (__uint_as_float(d[c + 1]) - dl.y);
          const float d2 = p[32 * hh + c + 2] * (__uint_as_float(d[c + 2]) - dl.z);
          const float d3 = p[32 * hh + c + 3] * (__uint_as_float(d[c + 3]) - dl.w);
          pk[c / 2] = pack2<T>(d0, d1);
          pk[c / 2 + 1] = pack2<T>(d2, d3);
        }
        tmem_st_x16(tDP + 16 * hh, pk);
        // the same 32 queries -> shared memory, row = key, M(query)-contiguous, 128B swizzle
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(pDS + w * 16384 + sw128_offset(r, 4 * hh + q4)) =
              make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(BAR(DS_FULL));
    }

    // ---- epilogue: warpgroup 0 stores dV, warpgroup 1 stores dK * scale -------------------
    if (NI > 0) {
      mbar_wait(BAR(DKV_FULL), 0);
      tc_fence_after();
    }
    const bool store_ok = key_g < a.Nk;
    const uint32_t tACC = lane_base + (w == 0 ? TM_DV : TM_DK);
    const float mul = (w == 0) ? 1.0f : a.scale;
    const bool shared_kv = (a.kv_heads == 1 && a.H > 1);
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
      if (store_ok) {
        if (!shared_kv) {
          T* base = (w == 0)
              ? reinterpret_cast<T*>(a.dv) + b * a.dv_sb + hk * a.dv_sh + (long long)key_g * a.dv_sn
              : reinterpret_cast<T*>(a.dk) + b * a.dk_sb + hk * a.dk_sh + (long long)key_g * a.dk_sn;
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 o4;
            o4.x = pack2<T>(__uint_as_float(acc[8 * v4 + 0]) * mul, __uint_as_float(acc[8 * v4 + 1]) * mul);
            o4.y = pack2<T>(__uint_as_float(acc[8 * v4 + 2]) * mul, __uint_as_float(acc[8 * v4 + 3]) * mul);
            o4.z = pack2<T>(__uint_as_float(acc[8 * v4 + 4]) * mul, __uint_as_float(acc[8 * v4 + 5]) * mul);
            o4.w = pack2<T>(__uint_as_float(acc[8 * v4 + 6]) * mul, __uint_as_float(acc[8 * v4 + 7]) * mul);
            *reinterpret_cast<uint4*>(base + c * 32 + v4 * 8) = o4;
          }
        } else {
          // keys/values shared by all heads: sum over heads in fp32 (reference: cu:1613-1619)
          float* accp = ((w == 0) ? a.dv_acc : a.dk_acc) + ((long long)b * a.Nk + key_g) * D + c * 32;
#pragma unroll
          for (int x = 0; x < 32; ++x) atomicAdd(accp + x, __uint_as_float(acc[x]) * mul);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// 3. finish kernels
// ------------------------------------------------------------------------------------------
struct DqFinishArgs {
  int B, H, Nq, D, nqt;
  float scale;
  const float* dq_acc;
  void* dq; long long sb, sh, sn;
};

// one thread = 8 consecutive features of one row
template <typename T>
__global__ void __launch_bounds__(256) bwd_dq_finish_kernel(const DqFinishArgs a) {
  const int tpr = a.D >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)a.B * a.H * a.Nq * tpr;
  if (idx >= total) return;
  const int c8 = (int)(idx % tpr);
  const long long rowid = idx / tpr;
  const int row = (int)(rowid % a.Nq);
  const int bh = (int)(rowid / a.Nq);
  const int b = bh / a.H, h = bh % a.H;
  const int qt = row >> 7, r = row & 127, wq = r >> 5, rl = r & 31;
  const float* tile = a.dq_acc + (((long long)bh * a.nqt + qt) * 4 + wq) * (32 * a.D);
  const float4 lo = *reinterpret_cast<const float4*>(tile + (2 * c8) * 128 + rl * 4);
  const float4 hi = *reinterpret_cast<const float4*>(tile + (2 * c8 + 1) * 128 + rl * 4);
  uint4 o4;
  o4.x = pack2<T>(lo.x * a.scale, lo.y * a.scale);
  o4.y = pack2<T>(lo.z * a.scale, lo.w * a.scale);
  o4.z = pack2<T>(hi.x * a.scale, hi.y * a.scale);
  o4.w = pack2<T>(hi.z * a.scale, hi.w * a.scale);
  T* dst = reinterpret_cast<T*>(a.dq) + b * a.sb + h * a.sh + (long long)row * a.sn + c8 * 8;
  *reinterpret_cast<uint4*>(dst) = o4;
}

struct KvFinishArgs {
  int B, Nk, D;
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
  void* workspace;
};

template <typename T, int D>
int run_backward_t(const BwdHostArgs& h, cudaStream_t stream, int* launches, const char** err,
                   cudaError_t* ce) {
  const BwdWorkspace w = bwd_workspace_layout(h.B, h.H, h.kv_heads, h.Nq, h.Nk, D);
  uint8_t* ws = reinterpret_cast<uint8_t*>(h.workspace);
  float* stats = reinterpret_cast<float*>(ws + w.stats_off);
  float* dq_acc = reinterpret_cast<float*>(ws + w.dq_off);
  float* dkv_acc = reinterpret_cast<float*>(ws + w.dkv_off);
  const bool shared_kv = (h.kv_heads == 1 && h.H > 1);
  const float log2e = 1.4426950408889634f;
  cudaError_t e;

  // zero the fp32 accumulators (dq, and dk/dv when shared across heads)
  e = cudaMemsetAsync(dq_acc, 0, w.dkv_off - w.dq_off + (shared_kv ? (size_t)2 * h.B * h.Nk * D * 4 : 0), stream);
  if (e != cudaSuccess) { *err = "cudaMemsetAsync(workspace)"; *ce = e; return FCSA_ERR_CUDA; }

  // 1. preprocess
  {
    PrepArgs pa;
    pa.B = h.B; pa.H = h.H; pa.Nq = h.Nq; pa.D = D; pa.nqt = w.nqt;
    pa.c2 = h.shift * log2e;
    pa.o = h.o.ptr; pa.o_sb = h.o.sb; pa.o_sh = h.o.sh; pa.o_sn = h.o.sn;
    pa.d_o = h.d_o.ptr; pa.do_sb = h.d_o.sb; pa.do_sh = h.d_o.sh; pa.do_sn = h.d_o.sn;
    pa.inv_l = h.inv_l; pa.stats = stats;
    const int rows_per_block = 256 / (D / 8);
    const long long rows = (long long)h.B * h.H * w.nqt * 128;
    const long long grid = (rows + rows_per_block - 1) / rows_per_block;
    bwd_prep_kernel<T><<<(unsigned)grid, 256, 0, stream>>>(pa);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = "backward preprocess launch"; *ce = e; return FCSA_ERR_CUDA; }
    ++*launches;
  }
  // 2. main
  {
    CUtensorMap tq, tk, tv, tdo;
    if (make_tensor_map_bhnd(&tq, h.q.ptr, h.dtype_bf16, h.B, h.H, h.Nq, D, h.q.sb, h.q.sh, h.q.sn, 128) ||
        make_tensor_map_bhnd(&tk, h.k.ptr, h.dtype_bf16, h.B, h.kv_heads, h.Nk, D, h.k.sb, h.k.sh, h.k.sn, 128) ||
        make_tensor_map_bhnd(&tv, h.v.ptr, h.dtype_bf16, h.B, h.kv_heads, h.Nk, D, h.v.sb, h.v.sh, h.v.sn, 128) ||
        make_tensor_map_bhnd(&tdo, h.d_o.ptr, h.dtype_bf16, h.B, h.H, h.Nq, D, h.d_o.sb, h.d_o.sh, h.d_o.sn, 128)) {
      *err = "cuTensorMapEncodeTiled failed (backward)";
      return FCSA_ERR_INVALID;
    }
    BwdArgs a;
    a.B = h.B; a.H = h.H; a.Nq = h.Nq; a.Nk = h.Nk; a.kv_heads = h.kv_heads; a.causal = h.causal;
    a.has_mask = h.mask ? 1 : 0; a.nqt = w.nqt;
    a.c1 = h.scale * log2e; a.scale = h.scale;
    a.mask = h.mask; a.mask_sb = h.mask_sb;
    a.stats = stats; a.dq_acc = dq_acc;
    a.dk_acc = dkv_acc; a.dv_acc = dkv_acc + (size_t)h.B * h.Nk * D;
    a.dk = h.dk.ptr; a.dk_sb = h.dk.sb; a.dk_sh = h.dk.sh; a.dk_sn = h.dk.sn;
    a.dv = h.dv.ptr; a.dv_sb = h.dv.sb; a.dv_sh = h.dv.sh; a.dv_sn = h.dv.sn;
    using Cfg = BwdCfg<D>;
    auto kern = fcsa_bwd_kernel<T, D>;
    static bool attr_set = false;
    if (!attr_set) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
      if (e != cudaSuccess) { *err = "cudaFuncSetAttribute(bwd)"; *ce = e; return FCSA_ERR_CUDA; }
      attr_set = true;
    }
    const long long grid = (long long)((h.Nk + 127) / 128) * h.B * h.H;
    kern<<<(unsigned)grid, Cfg::kThreads, Cfg::kSmem, stream>>>(tq, tk, tv, tdo, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = "backward kernel launch"; *ce = e; return FCSA_ERR_CUDA; }
    ++*launches;
  }
  // 3. finish
  {
    DqFinishArgs fa;
    fa.B = h.B; fa.H = h.H; fa.Nq = h.Nq; fa.D = D; fa.nqt = w.nqt; fa.scale = h.scale;
    fa.dq_acc = dq_acc; fa.dq = h.dq.ptr; fa.sb = h.dq.sb; fa.sh = h.dq.sh; fa.sn = h.dq.sn;
    const long long total = (long long)h.B * h.H * h.Nq * (D / 8);
    bwd_dq_finish_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(fa);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = "dq finish launch"; *ce = e; return FCSA_ERR_CUDA; }
    ++*launches;
    if (shared_kv) {
      for (int which = 0; which < 2; ++which) {
        KvFinishArgs ka;
        ka.B = h.B; ka.Nk = h.Nk; ka.D = D;
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
  }
  return FCSA_OK;
}

inline int run_backward(const BwdHostArgs& h, cudaStream_t stream, int* launches, const char** err,
                        cudaError_t* ce) {
  if (h.D != 64) {
    *err = "backward: head_dim 128 kernel not built yet (64 only)";
    return FCSA_ERR_UNSUPPORTED;
  }
  if (h.dtype_bf16) return run_backward_t<__nv_bfloat16, 64>(h, stream, launches, err, ce);
  return run_backward_t<__half, 64>(h, stream, launches, err, ce);
}

}  // namespace fcsa
