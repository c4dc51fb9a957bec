// sm_100a building blocks used by every kernel in this library: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 MMA / TMEM load-store / alloc, UMMA shared-memory
// and instruction descriptors.  Everything is inline PTX; nothing here depends on
// CUTLASS/CuTe or torch.
//
// This header replaces the reference's whole tile library
// (flash_cosine_sim_attention_cuda.cu:89-1067: mem::shared_fragment,
// rowsum_accumulator, layout::*, mma::warp_tile) - none of it is reused.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Optional in-kernel timeline (tests/cuda/trace_*.cu build with -DFCSA_TRACE): one lane per role
// of one CTA stamps clock64() at pipeline events.  Compiled out of the product library.
#ifdef FCSA_TRACE
#ifndef FCSA_TRACE_CTA
#define FCSA_TRACE_CTA 0
#endif
__device__ long long g_fcsa_trace[8][48][8];
#define FCSA_TR(role, it, slot)                                                       \
  do {                                                                                \
    if (blockIdx.x == FCSA_TRACE_CTA && (it) < 48) g_fcsa_trace[role][it][slot] = clock64(); \
  } while (0)
#else
#define FCSA_TR(role, it, slot) do { } while (0)
#endif

// Per-CTA timeline (test builds only, -DFCSA_CTA_TIMELINE): clock64 of up to 8 events per CTA plus the SM id,
// to see how consecutive CTAs of one SM follow each other.
#ifdef FCSA_CTA_TIMELINE
__device__ long long g_fcsa_cta_t[4096][10];
#define FCSA_CTA_T(cond, slot)                                                        \
  do {                                                                                \
    if ((cond) && blockIdx.x < 4096) {                                                \
      g_fcsa_cta_t[blockIdx.x][slot] = clock64();                                     \
      if ((slot) == 0) {                                                              \
        unsigned smid_;                                                               \
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                            \
        g_fcsa_cta_t[blockIdx.x][8] = smid_;                                          \
        unsigned long long gt_;                                                       \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                       \
        g_fcsa_cta_t[blockIdx.x][9] = (long long)gt_;                                 \
      }                                                                               \
    }                                                                                 \
  } while (0)
// the same, one row per work item of a persistent kernel
#define FCSA_ITEM_T(cond, item, slot)                                                 \
  do {                                                                                \
    if ((cond) && (item) < 4096) {                                                    \
      g_fcsa_cta_t[item][slot] = clock64();                                           \
      if ((slot) == 0) {                                                              \
        unsigned smid_;                                                               \
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                            \
        g_fcsa_cta_t[item][8] = smid_;                                                \
      }                                                                               \
    }                                                                                 \
  } while (0)
#else
#define FCSA_CTA_T(cond, slot) do { } while (0)
#define FCSA_ITEM_T(cond, item, slot) do { } while (0)
#endif

namespace fcsa {

// ----------------------------------------------------------------------------------
// generic helpers
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a system-defined interval; a
// scheduler that polls several barriers must not).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin until the phase with the given parity has completed.  try_wait itself suspends
// the thread for a hardware-defined interval, so this is not a hot spin.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef FCSA_WATCHDOG
  // debug builds: report the barrier a thread is stuck on (shared-memory address) and stop the kernel
  const long long t_start = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t_start > 2000000000LL) {
      printf("stuck: block %d thread %d barrier smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (UMMA/TMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// explicit shared-state-space accesses (a generic pointer makes ptxas emit LD.E / ST.E, which
// go through address translation and cost extra wavefronts on broadcast reads)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}

// One warp stores 32 rows x 64 bytes to global memory, row r at dst + r * row_stride_bytes, where lane r holds
// row r in registers (w[16]).  Stored directly, each instruction would touch 32 different 128-byte lines with
// 16 bytes each - the LSU serialises that into 32 wavefronts.  Going through 2 KB of shared memory (16-byte
// chunks XOR-swizzled so that neither side has bank conflicts) every store instruction covers 8 rows x 64
// contiguous bytes.  Rows >= valid_rows are not stored.  The caller owns `stage` (this warp only).
__device__ __forceinline__ void warp_store_rows64(uint32_t stage, int lane, const uint32_t (&w)[16], uint8_t* dst,
                                                  long long row_stride_bytes, int valid_rows) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    sts128(stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = 8 * k + (lane >> 2);
    const float4 v = lds128f(stage + row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
    if (row < valid_rows) *reinterpret_cast<float4*>(dst + row * row_stride_bytes + c * 16) = v;
  }
  __syncwarp();      // every lane has read its chunks: the stage may be rewritten
}

// Programmatic dependent launch: every kernel of a step is launched with the
// programmatic-stream-serialization attribute.  `pdl_launch_dependents` (first thing in a kernel)
// lets the next kernel of the stream start being scheduled as soon as all CTAs of this one have
// started; `pdl_wait` blocks until the previous kernel has completed and its writes are visible -
// it must precede the first access to global memory that kernel may have produced.  Between the
// two a kernel runs its private prologue (barrier init, TMEM allocation, descriptor prefetch),
// which thereby overlaps the previous kernel's tail instead of adding to the launch gap.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// 16-byte read of data that is touched once: read-only path, no L1 allocation
__device__ __forceinline__ uint4 ldg_stream128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// release / acquire on a shared-memory word (CTA scope): hand-off of plain counters between warps
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// named barrier among a subset of the CTA's warps
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 4-D tiled load: coordinates are (c0 = innermost element index, c1, c2, c3).
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// 1-D bulk copy global -> shared (bytes multiple of 16), completion on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
      : "memory");
}

// register re-budgeting between warpgroups (all warps of a warpgroup must execute it)
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// 1-D bulk reduce-add (fp32) shared -> global; completion tracked by the bulk async-group.
__device__ __forceinline__ void bulk_reduce_add_f32(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile(
      "cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
      ::"l"(gdst), "r"(smem_src), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------
// TMEM allocation (one warp, .sync.aligned)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64-bit), SWIZZLE_128B flavour.  Field layout:
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version = 1 (sm_100)
//   [61,64) layout type (2 = SWIZZLE_128B)
// All our operand tiles are what a TMA SWIZZLE_128B box of 64 16-bit elements x R rows
// leaves in shared memory: row r at byte r*128, XOR-swizzled inside 1024-byte atoms.
//   K-major  operand (rows index M or N, the 64 elements are K):  SBO = 1024 (next 8 rows),
//            LBO unused; advancing K by 16 elements = +32 bytes on the start address.
//   MN-major operand (rows index K, the 64 elements are M or N):  SBO = 1024 (next 8 K rows),
//            LBO = byte distance to the next 64-wide M/N chunk; advancing K by 16 rows = +2048.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// K-major operand that is exactly ONE K = 16 step wide: what a TMA SWIZZLE_32B box of 16 16-bit
// elements x R rows leaves in shared memory (row r at byte r*32, XOR-swizzled inside 256-byte atoms).
// SBO = 256 (next 8 rows); layout type 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}

// Instruction descriptor for kind::f16 (fp16/bf16 inputs, fp32 accumulate).
//   [4,6) D format (1 = f32)   [7,10) A format   [10,13) B format (0 = f16, 1 = bf16)
//   [15] A major  [16] B major (0 = K-major, 1 = MN-major)   [17,23) N>>3   [24,29) M>>4
template <typename T>
struct umma_fmt;
template <>
struct umma_fmt<__half> {
  static constexpr uint32_t v = 0;
};
template <>
struct umma_fmt<__nv_bfloat16> {
  static constexpr uint32_t v = 1;
};

template <typename T>
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                  uint32_t b_mn_major) {
  return (1u << 4) | (umma_fmt<T>::v << 7) | (umma_fmt<T>::v << 10) | (a_mn_major << 15) |
         (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]     (A must be K-major: lane = row, 2 x 16-bit per column)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives (count 1) once every tcgen05 op issued so far by this thread has completed.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// ----------------------------------------------------------------------------------
// TMEM <-> registers.  32x32b shape: warp w of a warpgroup owns lanes 32*(w%4)..+31,
// thread t <-> lane, register j <-> column (taddr.col + j).
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7])
      : "memory");
}

// ----------------------------------------------------------------------------------
// numeric helpers
// ----------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair of x <= 0 on the FMA / ALU pipes (no MUFU): x = n + r with n = round(x),
// r in [-0.5, 0.5]; cubic minimax for 2^r (max relative error 1.6e-4, below half an ulp of the
// 16-bit P it feeds); the exponent is inserted with an integer shift-add.  The attention kernels
// are MUFU-bound at head dim 64 (16384 exps vs 524 MMA cycles per 128x128 tile), so a fraction of
// the exps is computed this way to unload the MUFU pipe.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));      // 1.5 * 2^23: low bits = round(x)
  const float2 n = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(r, make_float2(5.676588789e-02f, 5.676588789e-02f), make_float2(2.427372634e-01f, 2.427372634e-01f));
  p = __ffma2_rn(p, r, make_float2(6.929193139e-01f, 6.929193139e-01f));
  p = __ffma2_rn(p, r, make_float2(9.999317527e-01f, 9.999317527e-01f));
  float2 y;
  y.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  y.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return y;
}

// pack two fp32 into one 32-bit word of 16-bit values, `lo` in bits [0,16)
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  uint32_t r;
  // satfinite: an out-of-range value becomes +-65504 instead of inf (fp16 has no headroom for
  // exp(scale * q.k) when grouped l2norm lets q.k exceed 1)
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <>
__device__ __forceinline__ float to_float<__half>(__half v) {
  return __half2float(v);
}
// unpack a 32-bit word of two 16-bit values
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t w);
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  __half2 h = *reinterpret_cast<__half2*>(&w);
  return __half22float2(h);
}

// Byte offset of the 16-byte chunk (row r, chunk c of 8) inside a SWIZZLE_128B tile whose
// rows are 128 bytes: chunk index is XORed with (row mod 8).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) {
  return r * 128u + ((c ^ (r & 7u)) << 4);
}

// warp_store_rows64 with the staging chunks laid out like 4 consecutive 16-byte columns (c0 .. c0+3, c0 a multiple
// of 4) of a 128-byte-swizzled tile at `tile`: 32 rows r0 .. r0+31.  The backward uses it with the chunks of its
// dS^T staging tile that this very warp writes during the main loop, so the epilogue needs no buffer of its own
// and no other warp is involved.
__device__ __forceinline__ void warp_store_rows64_sw128(uint32_t tile, int r0, int c0, int lane, const uint32_t (&w)[16],
                                                        uint8_t* dst, long long row_stride_bytes, int valid_rows) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    sts128(tile + sw128_offset(r0 + lane, c0 + c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = 8 * k + (lane >> 2);
    const float4 v = lds128f(tile + sw128_offset(r0 + row, c0 + c));
    if (row < valid_rows) *reinterpret_cast<float4*>(dst + row * row_stride_bytes + c * 16) = v;
  }
  __syncwarp();
}

}  // namespace fcsa
