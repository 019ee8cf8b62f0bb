// Shared device/host helpers for the b2m sm_100a kernels.
// PTX wrappers for mbarrier, cp.async, bulk copy (TMA engine), tcgen05 (UMMA) and TMEM.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/b2m.h"

#define B2M_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;      \
  } while (0)

namespace b2m {

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// 64-bit packed coordinate key: [b:16 | x:16 | y:16 | z:16], each biased by 32768 so that
// the signed range [-32768, 32767] is representable. Lexicographic (b,x,y,z) order == integer
// order of the key, which is what makes sorted strided maps identical to np.unique(axis=0).
// ---------------------------------------------------------------------------------------------
#define B2M_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define B2M_COORD_BIAS 32768

__host__ __device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return (unsigned)(b + B2M_COORD_BIAS) < 65535u && (unsigned)(x + B2M_COORD_BIAS) < 65536u &&
         (unsigned)(y + B2M_COORD_BIAS) < 65536u && (unsigned)(z + B2M_COORD_BIAS) < 65536u;
}
__host__ __device__ __forceinline__ uint64_t pack_key(int b, int x, int y, int z) {
  return ((uint64_t)(uint16_t)(b + B2M_COORD_BIAS) << 48) | ((uint64_t)(uint16_t)(x + B2M_COORD_BIAS) << 32) |
         ((uint64_t)(uint16_t)(y + B2M_COORD_BIAS) << 16) | (uint64_t)(uint16_t)(z + B2M_COORD_BIAS);
}
__host__ __device__ __forceinline__ void unpack_key(uint64_t k, int& b, int& x, int& y, int& z) {
  b = (int)((k >> 48) & 0xFFFF) - B2M_COORD_BIAS;
  x = (int)((k >> 32) & 0xFFFF) - B2M_COORD_BIAS;
  y = (int)((k >> 16) & 0xFFFF) - B2M_COORD_BIAS;
  z = (int)(k & 0xFFFF) - B2M_COORD_BIAS;
}
// murmur3 finaliser
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
#ifdef B2M_DEBUG_WAIT
// debug builds: a timed-out wait first records {block, thread, barrier address, parity, tag} in mapped host memory
__device__ unsigned int* g_b2m_dbg = nullptr;
#define B2M_WAIT_TAG(t) (t)
// debug builds: SM-clock stamps of block (0,0,0) at a few points of a kernel, written to the mapped host buffer
// ([4096 + id]; tools/small_conv_trace.py). Lane 0 of the calling warp writes; posted store, no read-back.
__device__ __forceinline__ void b2m_trace(int id) {
  if (g_b2m_dbg && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (threadIdx.x & 31) == 0)
    *reinterpret_cast<volatile unsigned int*>(g_b2m_dbg + 4096 + id) = (unsigned int)clock64() | 1u;
}
#define B2M_TRACE(id) b2m_trace(id)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 20)) {
      if (g_b2m_dbg && (threadIdx.x & 31) == 0) {
        const unsigned int slot = atomicAdd(g_b2m_dbg, 1u);
        if (slot < 255) {
          volatile unsigned int* e = g_b2m_dbg + 8 + slot * 8;
          e[0] = blockIdx.x; e[1] = threadIdx.x; e[2] = bar; e[3] = parity; e[4] = tag; e[5] = blockIdx.y;
        }
        __threadfence_system();
      }
      __nanosleep(2000000);
      __trap();
    }
  }
}
#else
#define B2M_TRACE(id) ((void)0)
// Every lane of the (converged) calling warp polls. Measured alternatives: lane 0 polling with
// the rest of the warp parked at __syncwarp is 1.3-1.5x SLOWER, a __nanosleep back-off in the non-critical waiters
// changes nothing: polling pressure on the mbarrier unit is not what makes the MMA issuer's stage take ~1000 cycles.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0) {
  (void)tag;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { __trap(); }
  }
}
#endif
// The issuer's waits almost always find the barrier complete: test once on the fall-through path (a taken branch
// costs a lone warp ~30 cycles, tools/umma_issue_bench.cu) and only then enter the bounded polling loop.
__device__ __forceinline__ void mbar_wait_likely(uint32_t bar, uint32_t parity, uint32_t tag = 0) {
  if (__builtin_expect(mbar_try_wait(bar, parity) != 0u, 1)) return;
  mbar_wait(bar, parity, tag);
}
// One pipeline stage of the forward / dgrad kernel in ONE straight-line block (no branches on the issue path):
// chunk 0 at (a_lo, b_lo) and an optional chunk 1 at (a_lo + 16 KB, b_lo + b_chunk), each either a full 64-wide SW128
// chunk (4 K = 16 steps, +32 bytes per step) or the 32-wide SW64 remainder (2 steps), then the two commits that
// release the A slot and the weight slot. c0_rem != 0: chunk 0 is the 32-wide one; f1: 0 = no chunk 1, 1 = full,
// 2 = 32-wide. Executed by the one elected thread.
__device__ __forceinline__ void umma_stage_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t b_chunk,
                                                uint32_t idesc, uint32_t accumulate, uint32_t c0_rem, uint32_t f1,
                                                uint32_t hi128, uint32_t hi64, uint32_t bar_a, uint32_t bar_b) {
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, pt, p0f, p1, p1f;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 hi0, hi1, a1, b1, ta, tb;\n\t"
      "setp.ne.b32 pacc, %5, 0;\n\t"
      "setp.eq.b32 pt, %0, %0;\n\t"
      "setp.eq.b32 p0f, %6, 0;\n\t"
      "setp.ne.b32 p1, %7, 0;\n\t"
      "setp.eq.b32 p1f, %7, 1;\n\t"
      "selp.b32 hi0, %8, %9, p0f;\n\t"
      "selp.b32 hi1, %8, %9, p1f;\n\t"
      "mov.b64 da, {%1, hi0};\n\t"
      "mov.b64 db, {%2, hi0};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pacc;\n\t"
      "add.u32 ta, %1, 2;\n\t"
      "add.u32 tb, %2, 2;\n\t"
      "mov.b64 da, {ta, hi0};\n\t"
      "mov.b64 db, {tb, hi0};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 ta, %1, 4;\n\t"
      "add.u32 tb, %2, 4;\n\t"
      "mov.b64 da, {ta, hi0};\n\t"
      "mov.b64 db, {tb, hi0};\n\t"
      "@p0f tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 ta, %1, 6;\n\t"
      "add.u32 tb, %2, 6;\n\t"
      "mov.b64 da, {ta, hi0};\n\t"
      "mov.b64 db, {tb, hi0};\n\t"
      "@p0f tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 a1, %1, 1024;\n\t"
      "add.u32 b1, %2, %3;\n\t"
      "mov.b64 da, {a1, hi1};\n\t"
      "mov.b64 db, {b1, hi1};\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 ta, a1, 2;\n\t"
      "add.u32 tb, b1, 2;\n\t"
      "mov.b64 da, {ta, hi1};\n\t"
      "mov.b64 db, {tb, hi1};\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 ta, a1, 4;\n\t"
      "add.u32 tb, b1, 4;\n\t"
      "mov.b64 da, {ta, hi1};\n\t"
      "mov.b64 db, {tb, hi1};\n\t"
      "@p1f tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 ta, a1, 6;\n\t"
      "add.u32 tb, b1, 6;\n\t"
      "mov.b64 da, {ta, hi1};\n\t"
      "mov.b64 db, {tb, hi1};\n\t"
      "@p1f tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(b_chunk), "r"(idesc), "r"(accumulate), "r"(c0_rem), "r"(f1), "r"(hi128),
        "r"(hi64), "r"(bar_a), "r"(bar_b) : "memory");
}
// cp.async 16 B global->shared, zero-filled when src_bytes == 0
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the same with an ignore-src predicate: `zero` != 0 writes 16 zero bytes and makes no memory request (LDGSTS.ZFILL)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool zero) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t}"
      ::"r"(dst), "l"(src), "r"((uint32_t)zero) : "memory");
}
// mbarrier arrive that fires when all prior cp.async of this thread have landed (counts as a normal arrival)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// 1-D bulk copy global->shared through the TMA engine, completes on an mbarrier (tx bytes)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// L2 prefetch of a contiguous global range through the TMA engine (no shared-memory destination, nothing to wait for)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA row gather (sm_100a): 4 rows (r0..r3, any order, out-of-range rows are zero-filled) x one box of columns
// starting at `col` of a 2-D tensor map whose box is {columns, 1}; the rows land back to back in shared memory
// with the map's swizzle; completes `4 * box bytes` on the mbarrier.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, uint32_t bar, int col, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ int4 ld_nc_int4(const int32_t* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- TMEM / tcgen05 ----
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same with the descriptors given as (lo, hi) 32-bit halves: the hot loops only add to the low word.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
// high word of a swizzled descriptor: SBO (16-byte units) | version 1 (bit 46) | layout type (bits 61-63) chosen by
// the row bytes of the tile: 128 -> SWIZZLE_128B (2), 64 -> SWIZZLE_64B (4), 32 -> SWIZZLE_32B (6)
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
// one lane of the (converged) warp; the same lane every time for a full warp, so MMAs and their commits pair up
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// low word: start address | LBO, both in 16-byte units
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane, reg j = column j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, 128-byte swizzle. Fields in 16-byte units.
//   bits [0,14) start address, [16,30) leading byte offset, [32,46) stride byte offset,
//   [46,48) version (1 on sm_100), [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Generic swizzled descriptor: row_bytes in {128, 64, 32} selects SWIZZLE_128B / 64B / 32B (layout type 2 / 4 / 6).
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
// UMMA instruction descriptor for kind::f16 with BF16 A/B and FP32 accumulator.
//   [4,6) c_format=1 (F32), [7,10) a_format=1 (BF16), [10,13) b_format=1 (BF16),
//   [15] a_major (0=K,1=MN), [16] b_major, [17,23) N>>3, [24,29) M>>4.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void red_shared_f64(uint32_t addr, double v) {
  asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

}  // namespace b2m
