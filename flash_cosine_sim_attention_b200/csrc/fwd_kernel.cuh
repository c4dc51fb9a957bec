// Forward kernel: o = (sum_j exp(scale*q.k_j - shift) v_j) / max(sum_j exp(..), 1e-10)
//
// Replaces the reference's forward_kernel (flash_cosine_sim_attention_cuda.cu:1072-1247).
// Same math (fixed shift instead of a running row max, cu:1216; running row sum, cu:1236;
// final 1/max(l, eps) scaling, cu:1239-1246; bottom-right aligned causal mask, cu:1097/1210;
// key-padding mask, cu:1198-1212) - completely different machine mapping:
//
//   * one work item = 256 query rows of one (batch, head): two 128-row tiles that ping-pong, each with its own
//     "engine" (MMA issuer + 8 softmax warps).  Persistent CTAs, one per SM, walk the items (heaviest first, snake
//     order): TMEM, barriers and the K/V ring live on, the next item's Q/K/V are prefetched, one tile's epilogue
//     runs under the other tile's MMAs
//   * warp 16  : TMA producer   (Q per item and tile, K and V tiles through mbarrier rings)
//   * warps 17,18: tcgen05 issuers, one per query tile (S_t = Q_t K_j^T into TMEM; O_t += P_t V_j with
//                P_t read from TMEM); warp 19 idles
//   * warps 0-7: "softmax" warps of tile 0, warps 8-15 of tile 1.  Per tile two warpgroups, each on
//                one half of the 128 key columns (thread == query row x 64 columns):
//                tcgen05.ld S -> exp2(fma) -> partial row sum in registers -> 16-bit P -> tcgen05.st.
//                Four warps per scheduler all on the exp stage hide each other's TMEM / MUFU latency;
//                the two partial row sums of a row meet once, in the epilogue (there is no running
//                max, so nothing else couples the two halves of a row).
//   * TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D).  Because there is
//     no row max there is no rescaling of O: the accumulator never leaves TMEM until the epilogue.
//     D = 64 : P0 [384,448) P1 [448,512) (packed 16-bit pairs) are separate columns, so the
//              softmax warpgroup hands S_t back as soon as it sits in registers and S_t(j+1)
//              is computed while the exps of tile j are still running;
//     D = 128: no spare columns - P_t overwrites the upper half of S_t and S_t(j+1) is issued
//              behind P_t(j) V_j in the (in-order) tensor pipe.
#pragma once

#include <type_traits>

#include "sm100_primitives.cuh"

#ifndef FCSA_FWD_SVC_REGS_64
#define FCSA_FWD_SVC_REGS_64 64
#endif
#ifndef FCSA_FWD_SVC_REGS_128
#define FCSA_FWD_SVC_REGS_128 96   // D = 128: no re-budgeting - with 64 the MMA issuers spill, and their reloads sit on the serial S -> P -> PV chain (1050 vs 850 us at (1,16,16384,128))
#endif
#ifndef FCSA_POLY_EVERY
#define FCSA_POLY_EVERY 4   // forward, D = 64: 1 of every N exp pairs runs on the FMA pipe (0 = none)
#endif

namespace fcsa {

struct FwdArgs {
  int B, H, Nq, Nk;
  int causal;        // bottom-right aligned: key j visible to query i iff j <= i + (Nk - Nq)
  int has_mask;      // key padding mask (B, Nk), nonzero = keep
  int kv_heads;      // 1 => keys/values shared by all heads, else == H
  int n_qblk;        // ceil(Nq / 256)
  float c1;          // scale * log2(e)
  float c2;          // shift * log2(e)
  const uint8_t* mask;
  long long mask_sb;
  void* o;
  long long o_sb, o_sh, o_sn;   // element strides of o (feature dim contiguous)
  int o_f32;                    // 1: o is float32 (strides in float32 elements), else the operand type
  float* inv_l;                 // (B, H, Nq) fp32, contiguous
  // additive bias on the logits (reference py:312, cu:1168,1214): element type = q's, [b][h][i][j] with
  // element strides (bias_sb = 0 when the bias has no batch dimension); rows are 16-byte aligned and
  // hold at least ceil8(Nk) elements.  Only read by the BIAS instantiation.
  const void* bias;
  long long bias_sb, bias_sh, bias_sn;
  const float* bias_amax;       // optional device scalar: shift += max(*bias_amax, 0)  (fp16 range guard)
};

template <int D>
struct FwdCfg {
  static constexpr int kTile = 128 * D * 2;          // bytes of one 128-row operand tile
  static constexpr int kKS = (D == 64) ? 3 : 2;       // K ring depth
  static constexpr int kVS = (D == 64) ? 3 : 2;       // V ring depth
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = 2 * kTile;
  static constexpr int kOffV = kOffK + kKS * kTile;
  static constexpr int kOffL = kOffV + kVS * kTile;   // partial row sums: 2 tiles x 2 halves x 128 floats
  // D = 64: 2 KB per softmax warp to turn the epilogue's row-per-lane stores into coalesced ones
  static constexpr bool kStageO = (D == 64);
  static constexpr int kOffStage = kOffL + 4096;      // (the row sums are double-buffered over work items)
  static constexpr int kOffBar = kOffStage + (kStageO ? 16 * 2048 : 0);
  static constexpr int kSmem = kOffBar + 256 + 1024;  // + alignment slack
  static constexpr int kThreads = 640;                // 16 softmax warps + 4 service warps
};

template <typename T, int D, bool BIAS = false>
__global__ void __launch_bounds__(640, 1)
fcsa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FwdArgs a) {
  using Cfg = FwdCfg<D>;
  constexpr int KS = Cfg::kKS, VS = Cfg::kVS, TILE = Cfg::kTile;
  constexpr int DCH = D / 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ);
  const uint32_t sK = smem_u32(smem + Cfg::kOffK);
  const uint32_t sV = smem_u32(smem + Cfg::kOffV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  // barrier indices
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int Q_FULL = 0, Q_EMPTY = 2, K_FULL = 4, K_EMPTY = K_FULL + KS, V_FULL = K_EMPTY + KS,
                V_EMPTY = V_FULL + VS, S_FULL = V_EMPTY + VS, P_FULL = S_FULL + 2,
                O_FULL = P_FULL + 2, S_FREE = O_FULL + 2, P_FREE = S_FREE + 2, O_FREE = P_FREE + 2,
                NBARS = O_FREE + 2;
  static_assert(NBARS * 8 + 4 <= 256, "barrier area");
  constexpr bool PSEP = (D == 64);     // P in its own TMEM columns (see header comment)
  constexpr int kPolyEvery = (D == 64) ? FCSA_POLY_EVERY : 8;   // 1 of every N exp pairs is emulated on the FMA pipe
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // ---- work items -------------------------------------------------------------------
  // One item = 256 query rows of one (batch, head).  Items are numbered heaviest first (causal: the
  // lowest query blocks see the fewest keys) and dealt to the CTAs of the grid in snake order
  // (round r: item r*G + c, or r*G + G-1-c on odd rounds), which balances as well as the hardware's
  // own longest-first dispatch of one CTA per item.  With a grid of one CTA per SM the CTA is
  // persistent: barriers, TMEM and the K/V ring live on across items, the next item's Q / K / V are
  // prefetched under the current item's last tiles and one tile's epilogue runs under the other tile's
  // MMAs.  (A grid of n_items CTAs degenerates to one item per CTA.)
  const int bh_count = a.B * a.H;
  const int n_items = a.n_qblk * bh_count;
  const int off = a.Nk - a.Nq;
  const int nkt = (a.Nk + 127) >> 7;
  struct Item {
    int b, h, hk, m0, n_t[2], NT;
  };
  auto item_index = [&](int r) -> int {
    const int G = gridDim.x, c = blockIdx.x;
    const int idx = r * G + ((r & 1) ? (G - 1 - c) : c);
    return idx < n_items ? idx : -1;
  };
  auto load_item = [&](int idx) -> Item {
    Item it;
    const int rank = idx / bh_count;
    const int bh = idx - rank * bh_count;
    const int qblk = a.causal ? (a.n_qblk - 1 - rank) : rank;   // heaviest causal blocks first
    it.b = bh / a.H;
    it.h = bh - it.b * a.H;
    it.hk = (a.kv_heads == 1) ? 0 : it.h;
    it.m0 = qblk * 256;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int row_lo = it.m0 + 128 * t;
      if (row_lo >= a.Nq) {
        it.n_t[t] = 0;
      } else if (!a.causal) {
        it.n_t[t] = nkt;
      } else {
        const int row_hi = min(row_lo + 127, a.Nq - 1);
        const int last_col = row_hi + off;
        it.n_t[t] = last_col < 0 ? 0 : min(nkt, (last_col >> 7) + 1);
      }
    }
    it.NT = max(it.n_t[0], it.n_t[1]);
    return it;
  };

  // ---- one-time setup ------------------------------------------------------------------
  if (warp == 16 && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int i = 0; i < KS; ++i) {
      mbar_init(BAR(K_FULL + i), 1);
      mbar_init(BAR(K_EMPTY + i), 2);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(BAR(V_FULL + i), 1);
      mbar_init(BAR(V_EMPTY + i), 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(BAR(Q_FULL + t), 1);
      mbar_init(BAR(Q_EMPTY + t), 1);
      mbar_init(BAR(O_FREE + t), 256);
      mbar_init(BAR(S_FULL + t), 1);
      mbar_init(BAR(P_FULL + t), 256);
      mbar_init(BAR(O_FULL + t), 1);
      mbar_init(BAR(S_FREE + t), 256);
      mbar_init(BAR(P_FREE + t), 1);
    }
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
#ifdef FCSA_WATCHDOG
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("barrier 0 at smem 0x%x (8 bytes each)\n", bar0);
#endif
  pdl_wait();          // q, k (normalised by the previous kernel), v, mask are read from here on

  // register split: 4 softmax warpgroups x kSmxRegs + the service warpgroup x kSvcRegs = 5 x 96, the pool the CTA was
  // launched with (640 threads x 96)
  constexpr int kSvcRegs = (D == 64) ? FCSA_FWD_SVC_REGS_64 : FCSA_FWD_SVC_REGS_128;
  constexpr int kSmxRegs = (5 * 96 - kSvcRegs) / 4;
  static_assert(kSmxRegs % 8 == 0, "setmaxnreg wants multiples of 8");
  if constexpr (kSvcRegs < 96) {
    if (warp >= 16) reg_dealloc<kSvcRegs>();
  }
  if (warp == 16) {
    // =============================== TMA producer ===============================
    // (elect_one, not lane == 0: ptxas then keeps descriptors/addresses in uniform registers
    //  instead of wrapping every UTMALDG/UTCHMMA in a divergence "waterfall" loop)
    if (elect_one()) {
      int g = 0;                                  // key tiles loaded so far (ring position)
      for (int k = 0;; ++k) {
        const int idx = item_index(k);
        if (idx < 0) break;
        const Item it = load_item(idx);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (k > 0) mbar_wait(BAR(Q_EMPTY + t), (k - 1) & 1);   // the previous item's last S_t has completed
          mbar_expect_tx(BAR(Q_FULL + t), TILE);
#pragma unroll
          for (int ch = 0; ch < DCH; ++ch)
            tma_load_4d(sQ + t * TILE + ch * 16384, &tm_q, BAR(Q_FULL + t), ch * 64, it.m0 + 128 * t, it.h, it.b);
        }
        for (int j = 0; j < it.NT; ++j, ++g) {
          const int ks = g % KS, vs = g % VS;
          mbar_wait(BAR(K_EMPTY + ks), ((g / KS) & 1) ^ 1);
          mbar_expect_tx(BAR(K_FULL + ks), TILE);
#pragma unroll
          for (int ch = 0; ch < DCH; ++ch)
            tma_load_4d(sK + ks * TILE + ch * 16384, &tm_k, BAR(K_FULL + ks), ch * 64, j * 128, it.hk, it.b);
          mbar_wait(BAR(V_EMPTY + vs), ((g / VS) & 1) ^ 1);
          mbar_expect_tx(BAR(V_FULL + vs), TILE);
#pragma unroll
          for (int ch = 0; ch < DCH; ++ch)
            tma_load_4d(sV + vs * TILE + ch * 16384, &tm_v, BAR(V_FULL + vs), ch * 64, j * 128, it.hk, it.b);
        }
      }
    }
  } else if (warp == 17 || warp == 18) {
    // =============================== MMA issuers ================================
    // One issuing thread per query tile (warp 17: tile 0, warp 18: tile 1).  Each follows only its
    // own softmax warpgroup (S_t(j+1) when S_t(j) is in registers, P_t(j) V_j when P_t(j) is stored),
    // so neither tile ever waits behind the other's barriers; the tensor pipe interleaves the two
    // instruction streams.  K / V ring slots are released by both (barrier count 2).
    const int t = warp - 17;
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc<T>(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc<T>(128, D, 0, 1);
      int g = 0;                                  // ring position of the item's first key tile
      int c = 0;                                  // iterations of THIS tile so far (parity of its S / P hand-shakes)
      for (int k = 0;; ++k) {
        const int idx = item_index(k);
        if (idx < 0) break;
        const Item it = load_item(idx);
        const int nt = t ? it.n_t[1] : it.n_t[0], NT = it.NT;
        // gj = ring position of key tile j of this item; `last`: no further S_t in this item -> Q_t may be replaced
        auto issue_S = [&](int gj, bool last) {
          const int ks = gj % KS;
          mbar_wait(BAR(K_FULL + ks), (gj / KS) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t o = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_ss(tmem + t * 128, umma_desc_sw128(sQ + t * TILE + o, 16, 1024),
                    umma_desc_sw128(sK + ks * TILE + o, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(BAR(S_FULL + t));
          umma_commit(BAR(K_EMPTY + ks));          // this tile's share of the release of K_j
          if (last) umma_commit(BAR(Q_EMPTY + t));
        };
        if (nt > 0) {
          mbar_wait(BAR(Q_FULL + t), k & 1);
          // PSEP: S_t of the previous item's last iteration must sit in registers before it is overwritten
          // (!PSEP: this S goes behind the previous P_t V in the in-order pipe, and P_FULL implied the reads)
          if (PSEP && c > 0) mbar_wait(BAR(S_FREE + t), (c - 1) & 1);
          tc_fence_after();
          // The two softmax warpgroups share the MUFU pipe: started half a tile apart, one computes at
          // full rate while the other is in its per-tile overhead (TMEM load/store, barriers).
#ifndef FCSA_FWD_NO_STAGGER
          if (k == 0 && t == 1 && it.n_t[0] > 0) mbar_wait(BAR(P_FULL + 0), 0);
#endif
          issue_S(g, nt == 1);
        } else {
          // nothing to compute for this tile: keep its per-item phases moving, in step with the producer
          // (an issuer running two items ahead would alias the parity the producer waits for)
          // O_FULL: only once the softmax warps are through with the previous item (O_FREE).  Without that an
          // issuer running through a series of empty items flips O_FULL faster than the softmax warps look at
          // it (parity aliasing), and a plain arrive could overtake the previous item's tcgen05.commit, which
          // is still in flight.  (Q_EMPTY needs no such care: Q_FULL(k) is only loaded after the previous
          // item's Q_EMPTY phase, so producer and issuer pace each other.)
          mbar_wait(BAR(Q_FULL + t), k & 1);
          mbar_arrive(BAR(Q_EMPTY + t));
          if (k > 0) mbar_wait(BAR(O_FREE + t), (k - 1) & 1);
          mbar_arrive(BAR(O_FULL + t));
        }
        for (int j = 0; j < NT; ++j) {
          const int vs = (g + j) % VS;
          if (j < nt) {
            if (PSEP && j + 1 < nt) {
              mbar_wait(BAR(S_FREE + t), c & 1);   // S_t(j) sits in registers: produce S_t(j+1) now
              FCSA_TR(0, j, t);
              tc_fence_after();
              issue_S(g + j + 1, j + 2 == nt);
            }
            mbar_wait(BAR(P_FULL + t), c & 1);
            FCSA_TR(0, j, 2 + 2 * t);
            if (j == 0 && k > 0) mbar_wait(BAR(O_FREE + t), (k - 1) & 1);   // the previous item's O_t has been read out
            mbar_wait(BAR(V_FULL + vs), ((g + j) / VS) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              umma_ts(tmem + 256 + t * D, tmem + (PSEP ? 384 + t * 64 : t * 128 + 64) + kk * 8,
                      umma_desc_sw128(sV + vs * TILE + kk * 2048, 16384, 1024), idesc_o,
                      (j > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(BAR(V_EMPTY + vs));
            if (PSEP) umma_commit(BAR(P_FREE + t));
            if (j + 1 < nt) {
              if (!PSEP) issue_S(g + j + 1, j + 2 == nt);   // P_t aliases S_t: S_t(j+1) goes behind P_t(j) V_j
            } else {
              umma_commit(BAR(O_FULL + t));
            }
            ++c;
          } else {
            // this tile has no work on key tile j (causal: tile 0 ends one key tile before tile 1;
            // or the tile is past the end of q): still release the ring slots the other tile uses
            // (paced by the fills: one arrival per refill of the slot, never two in one phase)
            mbar_wait(BAR(K_FULL + (g + j) % KS), ((g + j) / KS) & 1);
            mbar_arrive(BAR(K_EMPTY + (g + j) % KS));
            mbar_wait(BAR(V_FULL + vs), ((g + j) / VS) & 1);
            mbar_arrive(BAR(V_EMPTY + vs));
          }
        }
        g += NT;
      }
    }
  } else if (warp < 16) {
    // =============================== softmax warpgroups =========================
    if constexpr (kSmxRegs > 96) reg_alloc<kSmxRegs>();
    const int t = warp >> 3;                 // which 128-row tile
    const int half = (warp >> 2) & 1;        // which 64 of the 128 key columns of every tile
    const int wq = warp & 3;                 // TMEM lane quarter this warp may touch
    const int r = wq * 32 + lane;            // row inside the tile
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t tS = lane_base + t * 128 + 64 * half;
    const uint32_t tP = (PSEP ? lane_base + 384 + t * 64 : lane_base + t * 128 + 64) + 32 * half;
    const uint32_t tO = lane_base + 256 + t * D;
    float nc2 = -a.c2;
    if constexpr (BIAS) {
      if (a.bias_amax != nullptr) nc2 -= fmaxf(__ldg(a.bias_amax), 0.f) * 1.4426950408889634f;
    }
    const float c1 = a.c1;
    const bool tr_lane = (half == 0 && wq == 0 && lane == 0);
    int cc = 0;                              // iterations of this tile so far, over all items (barrier parities)
    for (int k = 0;; ++k) {
    const int idx = item_index(k);
    if (idx < 0) break;
    const Item it = load_item(idx);
    const int tl = t;                        // which 128 rows of the item this tile engine takes
    const int b = it.b, h = it.h, m0 = it.m0, nt = tl ? it.n_t[1] : it.n_t[0];
    const int row_g = m0 + 128 * tl + r;     // global query row
    // bias row of this query (clamped for the padding rows of the last tile, which are never stored)
    const T* brow = nullptr;
    if constexpr (BIAS)
      brow = reinterpret_cast<const T*>(a.bias) + (long long)b * a.bias_sb + (long long)h * a.bias_sh +
             (long long)min(row_g, a.Nq - 1) * a.bias_sn;
    // 32 bias values of chunk c (tile-relative columns [32 (2 half + c), +32)) as 16 packed words, already
    // turned into the addend of the exponent: bias * log2(e) - c2
    auto bias_chunk = [&](int col0, int c, float2 (&bb)[16]) {
      const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f), c2v = make_float2(nc2, nc2);
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        const int col = col0 + (2 * half + c) * 32 + 8 * v4;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (col < a.Nk) w = ldg_stream128(brow + col);
        bb[4 * v4 + 0] = __ffma2_rn(unpack2<T>(w.x), l2e, c2v);
        bb[4 * v4 + 1] = __ffma2_rn(unpack2<T>(w.y), l2e, c2v);
        bb[4 * v4 + 2] = __ffma2_rn(unpack2<T>(w.z), l2e, c2v);
        bb[4 * v4 + 3] = __ffma2_rn(unpack2<T>(w.w), l2e, c2v);
      }
    };
    float l = 0.f;
    float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);   // two partial row-sum chains

    for (int j = 0; j < nt; ++j, ++cc) {
      if (tr_lane) FCSA_TR(1 + t, j, 0);
      mbar_wait(BAR(S_FULL + t), cc & 1);
      if (tr_lane) FCSA_TR(1 + t, j, 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_x32(tS + 0, s0);
      tmem_ld_x32(tS + 32, s1);
      tmem_ld_wait();
      if (tr_lane) FCSA_TR(1 + t, j, 2);
      tc_fence_before();
      mbar_arrive(BAR(S_FREE + t));            // PSEP: S_t may be overwritten by tile j+1 now
      // before the first store of P(j):
      //   PSEP : P_t(j-1) V must have finished reading the P columns
      //   !PSEP: P_t overwrites the upper half of S_t - the other warpgroup's columns: it must hold them
      //          in registers (S_FREE doubles as that rendezvous; the issuer does not wait on it)
      // Waited for as late as possible - after the first chunk of exps - so it rarely costs anything.
      auto p_cols_free = [&]() {
        if (PSEP) {
          if (cc > 0) mbar_wait(BAR(P_FREE + t), (cc - 1) & 1);
        } else {
          mbar_wait(BAR(S_FREE + t), cc & 1);
        }
      };
      if (tr_lane) FCSA_TR(1 + t, j, 3);

      const int col0 = j * 128;
      const bool need_mask = a.has_mask || (col0 + 127 >= a.Nk) ||
                             (a.causal && (col0 + 127 > m0 + 128 * tl + off));
      if (!need_mask) {
        // packed f32x2 math: one FFMA2 / FADD2 per element pair (halves the FMA-pipe instruction
        // count next to the MUFU-bound exps); two independent row-sum chains per thread
        const float2 c1v = make_float2(c1, c1), c2v = make_float2(nc2, nc2);
        auto chunk = [&](const uint32_t(&s)[32], int c) {
          uint32_t pk[16];
          float2 bb[BIAS ? 16 : 1];
          if constexpr (BIAS) bias_chunk(col0, c, bb);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c1v,
                                        BIAS ? bb[BIAS ? i : 0] : c2v);
            // every kPolyEvery-th pair goes to the FMA pipe instead of the MUFU
            const float2 pp = (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1)
                                  ? ex2_poly2(x)
                                  : make_float2(ex2_approx(x.x), ex2_approx(x.y));
            if (i & 1) l2b = __fadd2_rn(l2b, pp);
            else l2a = __fadd2_rn(l2a, pp);
            pk[i] = pack2<T>(pp.x, pp.y);
          }
          if (c == 0) p_cols_free();
          tmem_st_x16(tP + c * 16, pk);
        };
        chunk(s0, 0);
        chunk(s1, 1);
      } else {
        // Masked tiles (diagonal, ragged end, key-padding mask): the same packed math plus one select
        // per element, driven by a per-thread visibility word per 32-column chunk:
        // visible iff column <= lim (causal / ragged end) and key-mask bit set
        int lim = a.Nk - 1;
        if (a.causal) lim = min(lim, row_g + off);
        lim -= col0;
        uint32_t vis[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int c0 = (2 * half + w) * 32;                       // first column of the chunk, tile-relative
          uint32_t word = 0xFFFFFFFFu;
          if (a.has_mask) {
            const int col = col0 + c0 + lane;
            const uint8_t mv = (col < a.Nk) ? a.mask[(long long)b * a.mask_sb + col] : uint8_t(0);
            word = __ballot_sync(0xFFFFFFFFu, mv != 0);
          }
          const int nvis = lim - c0 + 1;                            // columns of this chunk at or below lim
          const uint32_t range = nvis >= 32 ? 0xFFFFFFFFu : (nvis <= 0 ? 0u : ((1u << nvis) - 1u));
          vis[w] = word & range;
        }
        const float2 c1v = make_float2(c1, c1), c2v = make_float2(nc2, nc2);
        auto chunk = [&](const uint32_t(&s)[32], int c, uint32_t v) {
          uint32_t pk[16];
          float2 bb[BIAS ? 16 : 1];
          if constexpr (BIAS) bias_chunk(col0, c, bb);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c1v,
                                        BIAS ? bb[BIAS ? i : 0] : c2v);
            const float p0 = ((v >> (2 * i)) & 1u) ? ex2_approx(x.x) : 0.f;
            const float p1 = ((v >> (2 * i + 1)) & 1u) ? ex2_approx(x.y) : 0.f;
            l += p0 + p1;
            pk[i] = pack2<T>(p0, p1);
          }
          if (c == 0) p_cols_free();
          tmem_st_x16(tP + c * 16, pk);
        };
        chunk(s0, 0, vis[0]);
        chunk(s1, 1, vis[1]);
      }
      if (tr_lane) FCSA_TR(1 + t, j, 4);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(BAR(P_FULL + t));
      if (tr_lane) FCSA_TR(1 + t, j, 5);
    }

    // ---- epilogue: O * 1/max(l, eps) -> global ----------------------------------------
    // The two column halves of a row exchange their partial sums through shared memory, then each
    // stores half of the D output features.  The clamp only guards rows with no visible key (l == 0,
    // O == 0 -> o = 0; reference: cu:1239 uses 1e-10, too large here because shift = scale*groups
    // makes legitimately tiny row sums).
    l += (l2a.x + l2a.y) + (l2b.x + l2b.y);
    float* lbuf = reinterpret_cast<float*>(smem + Cfg::kOffL) + (k & 1) * 512 + t * 256;   // double-buffered over items
    lbuf[half * 128 + r] = l;
    named_bar_sync(1 + t, 256);
    l += lbuf[(half ^ 1) * 128 + r];
    const float inv = 1.0f / fmaxf(l, 1e-37f);
    const bool row_ok = row_g < a.Nq;
    const long long o_off = (long long)b * a.o_sb + (long long)h * a.o_sh + (long long)row_g * a.o_sn;
    T* orow = reinterpret_cast<T*>(a.o) + o_off;
    float* orow32 = reinterpret_cast<float*>(a.o) + o_off;
    const bool o_f32 = a.o_f32 != 0;
    constexpr int CPH = D / 64;              // 32-column chunks of O per half
    mbar_wait(BAR(O_FULL + t), k & 1);
    if (nt > 0) {
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < CPH; ++ch) {
        const int c = half * CPH + ch;
        uint32_t acc[32];
        tmem_ld_x32(tO + c * 32, acc);
        tmem_ld_wait();
        if (ch == CPH - 1) {                   // O_t is in registers: the next item's first P_t V may overwrite it
          tc_fence_before();
          mbar_arrive(BAR(O_FREE + t));
        }
        if constexpr (Cfg::kStageO) {
          // this warp's 32 query rows x 32 features through its 2 KB staging buffer (see warp_store_rows64)
          const uint32_t stage = smem_u32(smem + Cfg::kOffStage) + warp * 2048;
          const int row_w = m0 + 128 * tl + wq * 32;                // first query row of this warp
          const int valid = min(32, max(0, a.Nq - row_w));
          const long long w_off = (long long)b * a.o_sb + (long long)h * a.o_sh + (long long)row_w * a.o_sn + c * 32;
          uint32_t w16[16];
          if (o_f32) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
              for (int x = 0; x < 16; ++x) w16[x] = __float_as_uint(__uint_as_float(acc[16 * hf + x]) * inv);
              warp_store_rows64(stage, lane, w16, reinterpret_cast<uint8_t*>(reinterpret_cast<float*>(a.o) + w_off + 16 * hf),
                                a.o_sn * 4, valid);
            }
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x)
              w16[x] = pack2<T>(__uint_as_float(acc[2 * x]) * inv, __uint_as_float(acc[2 * x + 1]) * inv);
            warp_store_rows64(stage, lane, w16, reinterpret_cast<uint8_t*>(reinterpret_cast<T*>(a.o) + w_off), a.o_sn * 2, valid);
          }
        } else if (row_ok && o_f32) {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(orow32 + c * 32 + v * 4) =
                make_float4(__uint_as_float(acc[4 * v + 0]) * inv, __uint_as_float(acc[4 * v + 1]) * inv,
                            __uint_as_float(acc[4 * v + 2]) * inv, __uint_as_float(acc[4 * v + 3]) * inv);
        } else if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 w;
            w.x = pack2<T>(__uint_as_float(acc[8 * v + 0]) * inv, __uint_as_float(acc[8 * v + 1]) * inv);
            w.y = pack2<T>(__uint_as_float(acc[8 * v + 2]) * inv, __uint_as_float(acc[8 * v + 3]) * inv);
            w.z = pack2<T>(__uint_as_float(acc[8 * v + 4]) * inv, __uint_as_float(acc[8 * v + 5]) * inv);
            w.w = pack2<T>(__uint_as_float(acc[8 * v + 6]) * inv, __uint_as_float(acc[8 * v + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + v * 8) = w;
          }
        }
      }
    } else {
      mbar_arrive(BAR(O_FREE + t));
      if (row_ok && o_f32) {
#pragma unroll
        for (int v = 0; v < D / 8; ++v)
          *reinterpret_cast<float4*>(orow32 + half * (D / 2) + v * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (row_ok) {
#pragma unroll
        for (int v = 0; v < D / 16; ++v)
          *reinterpret_cast<uint4*>(orow + half * (D / 2) + v * 8) = make_uint4(0, 0, 0, 0);
      }
    }
    if (half == 0 && row_ok && a.inv_l != nullptr)
      a.inv_l[((long long)b * a.H + h) * a.Nq + row_g] = inv;
    }   // items
  }

  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem, 512);
}

}  // namespace fcsa
