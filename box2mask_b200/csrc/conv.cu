// Sparse convolution on sm_100a: TMA row gather -> tcgen05.mma (UMMA) implicit GEMM with TMEM accumulators.
//
// Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward+backward
// (reference call sites: /root/reference/models/detection_net.py:235-337, models/resnet.py:70-83).
//
// Both kernels consume a SORTED kernel map (b2m_kernel_map_sort): output rows are visited in `order`
// (rows grouped by their neighbour-occupancy bit mask inside blocks of rows), nbr is already permuted to
// that order and group_mask[g] says which offsets occur in each 64-row group, so whole (tile, offset)
// blocks without pairs are never touched. Tables have a row pitch of b2m_map_pitch(n_out) (-1 padded).
//
// Feature rows are fetched by the TMA engine in its row-gather mode (cp.async.bulk.tensor.2d ...
// tile::gather4): one instruction moves 4 arbitrary rows x one column box (<= 64 bf16) of the [rows, channels]
// tensor into shared memory with the swizzle the UMMA descriptors expect; rows with index -1 (no neighbour) and
// columns past the channel count are zero-filled by the hardware at no memory traffic. The box width follows
// the reduction width: 64 channels -> SWIZZLE_128B tiles, 32 -> SWIZZLE_64B, 16 -> SWIZZLE_32B.
// An 8-channel input (the 6-channel colour+normal input of conv0p1s1, models/detection_net.py:37, padded to 8) is
// one 16-byte piece per (row, offset): those are fetched with cp.async (LDGSTS) instead, 8 (forward) / 16
// (wgrad) offsets side by side in one SW128 tile, because a TMA request per 16 bytes would be request-bound.
//
//  conv_fwd_kernel   persistent, warp-specialised, output-stationary implicit GEMM.
//      work item  = T (1 or 2) tiles of 128 output rows x <=256 output columns, accumulators in TMEM,
//                   double-buffered across work items so the epilogue overlaps the next main loop.
//      A stage    = 128 gathered rows x one or two 64-wide chunks (the 32-wide remainder counts as a chunk) of the
//                   reduction of one kernel offset; for c_red in {16, 32} a stage holds 4 / 2 offsets as separate
//                   SW32 / SW64 sub-tiles. The host picks chunks per stage and ring depths (b2m_conv_forward).
//      producers  = 12 warps in groups of one warp per chunk of a stage; stage s belongs to group s % groups: one
//                   16-byte index load + one gather4 per lane.
//      B loader   = one thread: 1-D bulk copies (TMA engine) of the pre-swizzled weight slices of a stage; a weight
//                   slot is shared by the T tiles of the work item.
//      MMA issuers = one elected thread per tile of the work item: tcgen05.mma.cta_group::1.kind::f16, M=128,
//                   N=ntile, K=16. Its per-stage latency chain (~0.4 us) is what bounds the kernel.
//      epilogue   = 4 warps: tcgen05.ld -> per-warp fp32 staging -> BatchNorm column statistics (fp64 atomics)
//                   -> bf16 row stores (scattered through `order`).
//      The same kernel computes dgrad (weights packed transposed / mirrored).
//  conv_wgrad_kernel dW[k] = X_gathered^T * dY.  M = input channels (several offsets packed into the 128 M
//      rows when c_in <= 64), N = c_out, reduction over output rows in 64-row stages, both operands
//      MN-major (a gathered row IS a run of M / N elements). One CTA owns up to G accumulators (G*N <= 512 TMEM
//      columns) = G offset groups and a range of row groups; a dY stage is shared by the G gathers. fp32 vector
//      atomics into dW at the end, transposed through shared memory so that a warp instruction touches 4 lines.
#ifdef B2M_DEBUG_BUILD
#define B2M_DEBUG_WAIT
#endif
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#ifdef B2M_LEAN_ISSUER
#define B2M_ISSUER_WAIT(bar, parity, tag) mbar_wait_likely(bar, parity, tag)
#else
#define B2M_ISSUER_WAIT(bar, parity, tag) mbar_wait(bar, parity, tag)
#endif

namespace b2m {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
// Reduction layout of B per offset group kg:
//   c_red in {8, 16, 32} ("flat"): 64 / c_red consecutive offsets share ONE 64-wide SW128 slice [c_n][64].
//   otherwise: c_red is cut into chunks of 64 (SW128 slices [c_n][64]); a remainder of exactly 32 becomes a
//   SW64 slice [c_n][32]; remainders of 16 / 48 are zero-padded to a full chunk.
// A slice row is 128 (64) bytes; its 16-byte groups are XOR-swizzled with n & 7 ((n >> 1) & 3).
__host__ __device__ inline int conv_kpack(int c_red) { return (c_red == 8 || c_red == 16 || c_red == 32) ? 64 / c_red : 1; }
__host__ __device__ inline int conv_rem(int c_red) { return (conv_kpack(c_red) == 1 && (c_red % 64) == 32) ? 1 : 0; }
__host__ __device__ inline int conv_nfull(int c_red) {
  if (conv_kpack(c_red) > 1) return 1;
  return c_red / 64 + (((c_red % 64) == 16 || (c_red % 64) == 48) ? 1 : 0);
}

// one 16-byte group (8 bf16) of the packed image of one kernel; c_in_src <= c_in: input channels >= c_in_src are zero
// (the 6-channel network input padded to 8)
__device__ __forceinline__ void pack_group(const float* __restrict__ w, int kvol, int c_in_src, int c_in, int c_out, int mode,
                                           uint16_t* __restrict__ packed, int64_t gid) {
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  const int kpack = conv_kpack(c_red), nfull = conv_nfull(c_red), rem = conv_rem(c_red);
  const int64_t groups_kg = (int64_t)c_n * (nfull * 8 + rem * 4);  // 16-byte groups per offset group
  const int kg = (int)(gid / groups_kg);
  int64_t r = gid % groups_kg;
  int n, g, chunk;
  if (r < (int64_t)c_n * nfull * 8) {
    chunk = (int)(r / ((int64_t)c_n * 8));
    r %= (int64_t)c_n * 8;
    n = (int)(r / 8);
    g = (int)(r % 8) ^ (n & 7);
  } else {
    chunk = nfull;
    r -= (int64_t)c_n * nfull * 8;
    n = (int)(r / 4);
    g = (int)(r % 4) ^ ((n >> 1) & 3);
  }
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int p = g * 8 + e;  // position inside the slice
    int k, ch;
    if (kpack > 1) { k = kg * kpack + p / c_red; ch = p % c_red; } else { k = kg; ch = chunk * 64 + p; }
    float f = 0.f;
    if (k < kvol && ch < c_red) {
      const int ksrc = (mode == 1) ? (kvol - 1 - k) : k;
      const int ci = (mode == 0) ? ch : n, co = (mode == 0) ? n : ch;
      if (ci < c_in_src) f = w[((int64_t)ksrc * c_in_src + ci) * c_out + co];
    }
    v[e] = __float2bfloat16_rn(f);
  }
  // threads write consecutive 16-byte groups: the packed image is exactly gid * 16 bytes in
  *reinterpret_cast<uint4*>(packed + gid * 8) = *reinterpret_cast<const uint4*>(v);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int kvol, int c_in, int c_out, int mode,
                                    uint16_t* __restrict__ packed, int64_t total) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid < total) pack_group(w, kvol, c_in, c_in, c_out, mode, packed, gid);
}

// All kernels of a network in one launch. Job j: src[j] fp32 [kvol, c_in_src, c_out] -> dst[j]; meta[j] = {kvol,
// c_in_src, c_in, c_out, mode}; prefix[j] = first global 16-byte group of job j (prefix[n_jobs] = total).
__global__ void pack_weights_batched_kernel(const float* const* __restrict__ src, uint16_t* const* __restrict__ dst,
                                            const int32_t* __restrict__ meta, const int64_t* __restrict__ prefix,
                                            int n_jobs) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= prefix[n_jobs]) return;
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {                      // last job whose prefix <= gid
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= gid) lo = mid; else hi = mid - 1;
  }
  const int32_t* m = meta + 5 * lo;
  pack_group(src[lo], m[0], m[1], m[2], m[3], m[4], dst[lo], gid - prefix[lo]);
}

__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct MaskBits { uint32_t w0, w1, w2, w3; };
__device__ __forceinline__ uint32_t mask_word(const MaskBits& m, int i) { return i == 0 ? m.w0 : (i == 1 ? m.w1 : (i == 2 ? m.w2 : m.w3)); }
// the `cnt` (1, 2, 4 or 8; divides 32, k0 % cnt == 0) mask bits starting at offset k0
__device__ __forceinline__ uint32_t mask_bits(const MaskBits& m, int k0, int cnt) {
  return (mask_word(m, k0 >> 5) >> (k0 & 31)) & ((1u << cnt) - 1u);
}
__device__ __forceinline__ MaskBits mask_zero() { MaskBits m; m.w0 = m.w1 = m.w2 = m.w3 = 0; return m; }
__device__ __forceinline__ MaskBits mask_load(const uint32_t* __restrict__ gmask, int64_t g, int mwords) {
  MaskBits m = mask_zero();
  const uint32_t* p = gmask + g * mwords;
  m.w0 = __ldg(p);
  if (mwords > 1) m.w1 = __ldg(p + 1);
  if (mwords > 2) m.w2 = __ldg(p + 2);
  if (mwords > 3) m.w3 = __ldg(p + 3);
  return m;
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 4;
// Gather warps; A stage s is produced by warp s % kFwdProd. Issuing one gather4 per lane costs the issuing warp
// ~76 cycles per lane (ELECT + R2UR loop around UTMALDG; tools/tma_gather_bench.cu) whatever the box size, and
// scales linearly with the number of warps, so the TMA engine is fed by many warps.
constexpr int kFwdProd = 12;
constexpr int kFwdMma = 2;                                     // one MMA issuer warp per tile of a work item
// Warp w runs on scheduler sub-partition w % 4, and tcgen05.mma / tcgen05.commit queue up in that sub-partition's
// memory-instruction (MIO) queue behind the cp.async of any gather warp living there (ncu: a third of the issuers' time
// was `mio` stall on UTCHMMA / UTCBAR). So sub-partition 0 holds no gather warp: warps 0-3 epilogue (TMEM lane quarter =
// warp % 4), warps 4 and 8 the MMA issuers, warp 12 the weight loader, warp 16 idle, every other warp >= 5 a gather warp.
constexpr int kFwdWarps = 20;
constexpr int kFwdThreads = kFwdWarps * 32;                     // 640
constexpr int kFwdIssuer0 = 4, kFwdIssuer1 = 8, kFwdLoader = 12, kFwdIdle = 16;
__device__ __forceinline__ int fwd_gather_index(int warp) { return ((warp >> 2) - 1) * 3 + (warp & 3) - 1; }   // 0 .. 11
constexpr int kTileM = 128;
constexpr int kStagePitch = 33;
constexpr int kASlotBytes = kTileM * 128;

#ifdef B2M_DEBUG_BUILD
#define B2M_ABLATE(a, bit) (((a).ablate >> (bit)) & 1)
#else
#define B2M_ABLATE(a, bit) 0
#endif

struct FwdArgs {
  const uint16_t* x; const int32_t* nbr; const int32_t* order; const uint32_t* gmask; const uint8_t* w;
  uint16_t* y; double* colsum;
  int64_t n_out, n_pitch;
  int c_red, kvol, c_n, ntile, T, kpack, nkg, nfull, rem, wa, mwords, colstride, n_tiles, n_work;
  int a_slots, b_slots, b_bytes, kg_bytes;
  int cps, ncg, a_slot_bytes;   // KPACK == 1: 64-wide chunks per A stage (1 or 2), stages per offset, bytes of an A slot
  int off_b, off_stage, off_csum, off_bars, tmem_cols;
  int ablate;   // debug builds (B2M_ABLATE env): 1 = no A gathers, 2 = no MMAs, 4 = no epilogue stores, 8 = no B copies
  // fused epilogue (all optional): v = acc * scale[col] + shift[col] (+ residual[row][col]) (ReLU); statistics are
  // taken of v. Eval-mode BatchNorm folded into the convolution, bias + ReLU of the MLP heads, fp32 logits.
  const float* ep_scale; const float* ep_shift; const uint16_t* ep_res; int ep_relu;
  float* y32; int c_store;      // fp32 output [n_out, c_store] (c_store <= c_n real columns) instead of the bf16 y
  // offsets split over gridDim.z CTAs (levels with few row tiles): slice z accumulates offset groups
  // [z * kg_per, (z + 1) * kg_per) and writes its fp32 partial tile to part[z][n_out][c_n]; conv_finalize_kernel sums
  // the slices in a fixed order (deterministic), applies the epilogue and takes the statistics.
  float* part; int ksplit, kg_per;
  int off_ep;
  int lean_off; // 1 = general MMA issue loop even where the lean one applies (b2m_set_option B2M_OPT_ISSUER, tests)
  int ldgsts;   // KPACK == 1: A rows fetched by cp.async (LDGSTS) from every lane of the gather warps instead of TMA gather4
  // dgrad with the BatchNorm-backward reduction of the PRODUCER of this gradient fused into the epilogue (bnr_x != nullptr):
  // the statistics written to `colsum` are then (sum g, sum g * xhat) with g = bf16(v) gated by the producer's ReLU mask and
  // xhat = (bnr_x - mean) * invstd - what b2m_bn_backward_reduce computes in a pass of its own over v, x and the mask.
  const uint16_t* bnr_x; const uint8_t* bnr_mask; const float* bnr_mean; const float* bnr_invstd;
};

// bits [lo, hi) of a 128-bit mask
__device__ __forceinline__ uint32_t range_word(int lo, int hi, int wd) {
  const int b0 = max(lo - 32 * wd, 0), b1 = min(hi - 32 * wd, 32);
  if (b1 <= b0) return 0u;
  const uint32_t upto = (b1 == 32) ? 0xFFFFFFFFu : ((1u << b1) - 1u);
  return upto & ~((1u << b0) - 1u);       // b0 < 32 here
}
// the kernel offsets this CTA accumulates: all of them, or slice blockIdx.z of the offset groups (a.ksplit > 1)
__device__ __forceinline__ MaskBits fwd_slice_mask(const FwdArgs& a) {
  MaskBits m;
  if (a.ksplit <= 1) { m.w0 = m.w1 = m.w2 = m.w3 = 0xFFFFFFFFu; return m; }
  const int lo = (int)blockIdx.z * a.kg_per * a.kpack, hi = min(a.nkg, ((int)blockIdx.z + 1) * a.kg_per) * a.kpack;
  m.w0 = range_word(lo, hi, 0); m.w1 = range_word(lo, hi, 1); m.w2 = range_word(lo, hi, 2); m.w3 = range_word(lo, hi, 3);
  return m;
}
// Offsets present in a 128-row tile, restricted to this CTA's slice. EVERY role (gather warps, weight loader, MMA
// issuers, epilogue) derives its stage sequence from this one function, which is what keeps them in step.
__device__ __forceinline__ MaskBits fwd_tile_mask(const FwdArgs& a, int tile, const MaskBits& sl) {
  MaskBits m = mask_zero();
  if (tile < a.n_tiles) {
    if (a.gmask == nullptr) { m.w0 = 1u; return m; }  // identity map (kvol == 1)
    const int64_t ngroups = (a.n_out + 63) / 64;
    const int64_t g0 = 2 * (int64_t)tile;
    m = mask_load(a.gmask, g0, a.mwords);
    if (g0 + 1 < ngroups) {
      const MaskBits m2 = mask_load(a.gmask, g0 + 1, a.mwords);
      m.w0 |= m2.w0; m.w1 |= m2.w1; m.w2 |= m2.w2; m.w3 |= m2.w3;
    }
    m.w0 &= sl.w0; m.w1 &= sl.w1; m.w2 &= sl.w2; m.w3 &= sl.w3;
  }
  return m;
}

// next offset group >= from that has a pair in the union mask u of the work item's tiles (nkg if none)
template <int KPACK>
__device__ __forceinline__ int next_group(const MaskBits& u, int from, int nkg, int mwords) {
  if (KPACK == 1) {
    for (int wd = from >> 5; wd < mwords; ++wd) {
      uint32_t bits = mask_word(u, wd);
      if (wd == (from >> 5)) bits &= 0xFFFFFFFFu << (from & 31);
      if (bits) return (wd << 5) + __ffs(bits) - 1;
    }
    return nkg;
  } else {
    for (int kg = from; kg < nkg; ++kg)
      if (mask_bits(u, kg * KPACK, KPACK)) return kg;
    return nkg;
  }
}
__device__ __forceinline__ MaskBits mask_or(const MaskBits& a, const MaskBits& b) {
  MaskBits m; m.w0 = a.w0 | b.w0; m.w1 = a.w1 | b.w1; m.w2 = a.w2 | b.w2; m.w3 = a.w3 | b.w3; return m;
}

struct Ring {
  int slot; uint32_t phase; int n;
  __device__ __forceinline__ void init(int n_) { slot = 0; phase = 0; n = n_; }
  __device__ __forceinline__ void next() { if (++slot == n) { slot = 0; phase ^= 1u; } }
  __device__ __forceinline__ void advance(int k) { slot += k; while (slot >= n) { slot -= n; phase ^= 1u; } }
  __device__ __forceinline__ Ring at(int k) const { Ring r = *this; r.advance(k); return r; }
};

// the 4 (2) K = 16 steps of one 64-wide SW128 (32-wide SW64) chunk: +32 bytes per step in both descriptors
__device__ __forceinline__ void umma_chunk4(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  umma_bf16_lohi(d, a_lo, hi, b_lo, hi, idesc, acc);
  umma_bf16_lohi(d, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
  umma_bf16_lohi(d, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
  umma_bf16_lohi(d, a_lo + 6, hi, b_lo + 6, hi, idesc, 1u);
}
__device__ __forceinline__ void umma_chunk2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  umma_bf16_lohi(d, a_lo, hi, b_lo, hi, idesc, acc);
  umma_bf16_lohi(d, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
}


// ------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) row gathers into UMMA tiles
// ------------------------------------------------------------------------------------------------
// One gather4 costs the issuing warp ~76 cycles PER LANE (the operands travel to uniform registers one lane at a
// time), so a warp needs ~2400 cycles to request the 128 rows of a chunk and a ring slot turns around in ~5700 cycles:
// with 3 slots per ring that, not bandwidth, set the stage rate of round 1. A cp.async is an ordinary SIMT
// instruction: a warp requests 4 rows x 128 bytes (8 lanes per row, so every request is a full 128-byte line piece)
// per instruction and the whole chunk in ~32 instructions + address arithmetic; completion is reported to the stage's
// mbarrier by cp.async.mbarrier.arrive.noinc from every lane (fire and forget). The writes go through the generic
// proxy: the MMA issuer executes fence.proxy.async after its wait on the stage barrier.
//
// R[u] holds the gathered row index (-1 = none: zero-filled, no memory request) of tile row 32 * u + lane.
// 128 rows x 64 channels -> SW128 tile (row r at r * 128, 16-byte piece j at (j ^ (r & 7)) * 16);
// `pieces` = 16-byte pieces of the chunk that exist in the tensor (8, or fewer for a zero-padded last chunk).
template <int NROWS>
__device__ __forceinline__ void ldgsts_rows_sw128(uint32_t dst, const uint8_t* __restrict__ src, uint32_t row_bytes,
                                                  const int (&R)[NROWS / 32], int lane, int pieces) {
  const int j = lane & 7, s = lane >> 3;
  const uint32_t d_even = dst + (uint32_t)(s * 128 + ((j ^ s) << 4));
  const uint32_t d_odd = dst + (uint32_t)((4 + s) * 128 + ((j ^ (4 + s)) << 4));
  const uint8_t* src_lane = src + j * 16;
#ifdef B2M_LDGSTS_LEAN_ADDR
  asm volatile("" : "+l"(src_lane));     // one opaque 64-bit base: the row address is a single IMAD.WIDE
#endif
  const bool colok = j < pieces;
#pragma unroll
  for (int i = 0; i < NROWS / 4; ++i) {
    const int r = __shfl_sync(0xFFFFFFFFu, R[i >> 3], ((4 * i) & 31) + s);
    const bool ok = colok && r >= 0;
    // (a row without a neighbour reads nothing: the address only has to be well formed)
    cp_async16_zfill(((i & 1) ? d_odd : d_even) + (uint32_t)((i >> 1) * 1024),
                     src_lane + (uint64_t)(uint32_t)max(r, 0) * (uint64_t)row_bytes, !ok);
  }
}
// NROWS rows x 32 channels -> SW64 tile (row r at r * 64, 16-byte piece j at (j ^ ((r >> 1) & 3)) * 16)
template <int NROWS>
__device__ __forceinline__ void ldgsts_rows_sw64(uint32_t dst, const uint8_t* __restrict__ src, uint32_t row_bytes,
                                                 const int (&R)[NROWS / 32], int lane) {
  const int j = lane & 3, s = lane >> 2;
  const uint32_t d0 = dst + (uint32_t)(s * 64 + ((j ^ ((lane >> 3) & 3)) << 4));
  const uint8_t* src_lane = src + j * 16;
#ifdef B2M_LDGSTS_LEAN_ADDR
  asm volatile("" : "+l"(src_lane));
#endif
#pragma unroll
  for (int i = 0; i < NROWS / 8; ++i) {
    const int r = __shfl_sync(0xFFFFFFFFu, R[i >> 2], ((8 * i) & 31) + s);
    cp_async16_zfill(d0 + (uint32_t)(i * 512), src_lane + (uint64_t)(uint32_t)max(r, 0) * (uint64_t)row_bytes, r < 0);
  }
}

// Rolled variants (one loop per index register, ~1 KB of SASS per instance instead of ~3.7 KB) for the wgrad kernel's
// optional cp.async mode. NOT for the forward kernel: how fast a gather warp gets its requests out sets the turnaround of
// a ring slot, and with these loops (shuffle -> address -> copy serialised per iteration, ~1300 cycles per chunk) the
// forward kernel measured 0.61 ms on k27 96->96 against 0.48 ms with the unrolled bursts above (TMA gather4: 0.59 ms).
template <int NROWS>
__device__ __forceinline__ void ldgsts_rows_sw128_rolled(uint32_t dst, const uint8_t* __restrict__ src, uint32_t row_bytes,
                                                  const int (&R)[NROWS / 32], int lane, int pieces) {
  const int j = lane & 7, s = lane >> 3;
  const uint32_t d_even = dst + (uint32_t)(s * 128 + ((j ^ s) << 4));
  const uint32_t d_odd = dst + (uint32_t)((4 + s) * 128 + ((j ^ (4 + s)) << 4));
  const uint8_t* src_lane = src + j * 16;
  const bool colok = j < pieces;
#pragma unroll
  for (int u = 0; u < NROWS / 32; ++u) {
    const int Ru = R[u];
#pragma unroll 2
    for (int ii = 0; ii < 8; ++ii) {                 // rows 32 u + 4 ii + s
      const int r = __shfl_sync(0xFFFFFFFFu, Ru, 4 * ii + s);
      const bool ok = colok && r >= 0;
      // (a row without a neighbour reads nothing: the address only has to be well formed)
      cp_async16_zfill(((ii & 1) ? d_odd : d_even) + (uint32_t)((4 * u + (ii >> 1)) * 1024),
                       src_lane + (uint64_t)(uint32_t)max(r, 0) * (uint64_t)row_bytes, !ok);
    }
  }
}
// NROWS rows x 32 channels -> SW64 tile (row r at r * 64, 16-byte piece j at (j ^ ((r >> 1) & 3)) * 16)
template <int NROWS>
__device__ __forceinline__ void ldgsts_rows_sw64_rolled(uint32_t dst, const uint8_t* __restrict__ src, uint32_t row_bytes,
                                                 const int (&R)[NROWS / 32], int lane) {
  const int j = lane & 3, s = lane >> 2;
  const uint32_t d0 = dst + (uint32_t)(s * 64 + ((j ^ ((lane >> 3) & 3)) << 4));
  const uint8_t* src_lane = src + j * 16;
#pragma unroll
  for (int u = 0; u < NROWS / 32; ++u) {
    const int Ru = R[u];
#pragma unroll 2
    for (int ii = 0; ii < 4; ++ii) {                 // rows 32 u + 8 ii + s
      const int r = __shfl_sync(0xFFFFFFFFu, Ru, 8 * ii + s);
      cp_async16_zfill(d0 + (uint32_t)((4 * u + ii) * 512), src_lane + (uint64_t)(uint32_t)max(r, 0) * (uint64_t)row_bytes, r < 0);
    }
  }
}

// KPACK = offsets per A stage: 1 (c_red >= 48: 64-wide chunks), 2 / 4 (c_red 32 / 16: SW64 / SW32 sub-tiles),
// 8 (c_red 8: cp.async path). A template parameter so that the single-thread MMA issue loop has no mode branches.
// BNR: the dgrad instantiation whose epilogue also takes the BatchNorm-backward reduction of the producer layer (a.bnr_*);
// a template parameter so that the plain instantiations keep exactly the code (and registers) they had.
template <int KPACK, bool BNR = false>
__global__ void __launch_bounds__(kFwdThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_rem, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SA = a.a_slots, SB = a.b_slots;
  const uint32_t bars = smem_base + a.off_bars;
  const uint32_t a_full = bars, a_empty = bars + 8 * SA;
  const uint32_t b_full = bars + 16 * SA, b_empty = b_full + 8 * SB;
  const uint32_t acc_full = b_empty + 8 * SB, acc_empty = acc_full + 16;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + a.off_bars + 16 * SA + 16 * SB + 32);
  const int n0 = blockIdx.y * a.ntile;
  const int nch = a.nfull + a.rem;
  const int nst = (KPACK == 1) ? a.ncg : nch;     // pipeline stages (and weight slots) per offset group
  const MaskBits sl = fwd_slice_mask(a);
  if (warp == 0) B2M_TRACE(0);

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(a_full + 8 * s, KPACK == 1 ? (a.ldgsts == 1 ? 32 * a.cps : a.cps) : ((KPACK == 2 && a.ldgsts) ? 32 : 1)); mbar_init(a_empty + 8 * s, 1); }
    // every tile's MMA issuer releases a B slot / completes an accumulator set: T arrivals each
    for (int s = 0; s < SB; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, a.T); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + 8 * s, a.T); mbar_init(acc_empty + 8 * s, kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 6) { tmem_alloc(smem_u32(tmem_ptr_s), (uint32_t)a.tmem_cols); tmem_relinquish(); }
  if (warp == 7 && lane == 0) { tma_prefetch_desc(&tm_main); tma_prefetch_desc(&tm_rem); }
  if (warp < kEpiWarps) {
    for (int i = tid; i < 2 * a.ntile; i += kEpiWarps * 32) reinterpret_cast<double*>(smem + a.off_csum)[i] = 0.0;
    if (a.ep_scale != nullptr || a.ep_shift != nullptr) {
      float* ep = reinterpret_cast<float*>(smem + a.off_ep);
      for (int i = tid; i < a.ntile; i += kEpiWarps * 32) {
        ep[i] = a.ep_scale ? __ldg(a.ep_scale + n0 + i) : 1.f;
        ep[a.ntile + i] = a.ep_shift ? __ldg(a.ep_shift + n0 + i) : 0.f;
      }
    } else if (BNR) {
      float* ep = reinterpret_cast<float*>(smem + a.off_ep);
      for (int i = tid; i < a.ntile; i += kEpiWarps * 32) {
        ep[i] = __ldg(a.bnr_mean + n0 + i);
        ep[a.ntile + i] = __ldg(a.bnr_invstd + n0 + i);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (warp == 0) B2M_TRACE(1);

  if (warp < kEpiWarps) {
    // ================= epilogue warps: TMEM -> registers -> fused epilogue -> statistics + row stores ====
    // A lane owns one output row: its 32 (16) accumulator columns of a chunk go to global memory straight from
    // registers (64 contiguous bytes per row); the transposed copy in shared memory only feeds the per-column
    // sum / sum of squares, which accumulate in fp64 in shared memory and reach colsum once per CTA.
    // Fused epilogue (optional): v = acc * scale[col] + shift[col] (+ residual[row][col]) (ReLU) - eval-mode BatchNorm
    // folded into the convolution, bias + ReLU of the MLP heads. Split mode (a.part): the raw fp32 accumulators go to
    // this slice's partial tile instead and conv_finalize_kernel does the rest.
    const uint32_t stage_a = smem_base + a.off_stage + warp * (32 * kStagePitch * 4);
    const uint32_t csum_a = smem_base + a.off_csum;
    const uint32_t ep_a = smem_base + a.off_ep;              // [2][ntile] floats: scale, shift of this CTA's columns
    const bool ep_affine = (a.ep_scale != nullptr) || (a.ep_shift != nullptr);
    const bool split = a.part != nullptr;
    float* part = split ? a.part + (int64_t)blockIdx.z * a.n_out * a.c_n : nullptr;
    int wi = 0;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
      const int par = wi & 1;
      mbar_wait(acc_full + 8 * par, (uint32_t)(wi >> 1) & 1u, 1);
      tc_fence_after();
      if (warp == 0 && wi == 0) B2M_TRACE(30);
      for (int t = 0; t < a.T; ++t) {
        const int tile = w * a.T + t;
        if (tile >= a.n_tiles) break;
        const MaskBits m = fwd_tile_mask(a, tile, sl);
        const bool has_acc = (m.w0 | m.w1 | m.w2 | m.w3) != 0;
        const int64_t pos = (int64_t)tile * kTileM + warp * 32 + lane;  // this lane's row (position in `order`)
        int32_t orow = -1;
        if (pos < a.n_out) orow = a.order ? __ldg(a.order + pos) : (int32_t)pos;
        const int64_t rbase = (int64_t)(orow >= 0 ? orow : 0) * a.c_n + n0;
        const int nchunk32 = (a.ntile + 31) >> 5;
        for (int cc = 0; cc < nchunk32; ++cc) {
          const int cw = min(32, a.ntile - cc * 32);   // 32 or 16
          uint32_t v[32];
          // fused BatchNorm-backward reduction: the producer's rows and ReLU gate are requested before anything else of
          // the chunk, so that their latency overlaps the accumulator load and the residual add
          uint4 xr[4];
          uint32_t gate = 0xFFFFFFFFu;
          if (BNR) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              xr[q] = (orow >= 0 && q * 8 < cw) ? __ldg(reinterpret_cast<const uint4*>(a.bnr_x + rbase + cc * 32 + q * 8))
                                                : make_uint4(0u, 0u, 0u, 0u);
            if (a.bnr_mask != nullptr && orow >= 0) {
              const uint8_t* mp = a.bnr_mask + ((int64_t)orow * a.c_n + n0 + cc * 32) / 8;
              if (((a.c_n | n0) & 31) == 0) {
                gate = __ldg(reinterpret_cast<const uint32_t*>(mp));
              } else {
                gate = 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q * 8 < cw) gate |= (uint32_t)__ldg(mp + q) << (8 * q);
              }
            }
          }
          if (has_acc) {
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                   (uint32_t)((par * a.T + t) * a.colstride + cc * 32);
            if (cw == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (split) {
            // raw fp32 accumulators of this offset slice: 128 contiguous bytes per row
            if (orow >= 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q * 4 < cw)
                  *reinterpret_cast<uint4*>(part + rbase + cc * 32 + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
            continue;
          }
          if (ep_affine) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < cw)
                v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), ld_shared_f32(ep_a + (cc * 32 + j) * 4),
                                            ld_shared_f32(ep_a + (a.ntile + cc * 32 + j) * 4)));
          }
          if (a.ep_res != nullptr && orow >= 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q * 8 < cw) {
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(a.ep_res + rbase + cc * 32 + q * 8));
                const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(rb[e]);
                  v[q * 8 + 2 * e] = __float_as_uint(__uint_as_float(v[q * 8 + 2 * e]) + f.x);
                  v[q * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(v[q * 8 + 2 * e + 1]) + f.y);
                }
              }
            }
          }
          if (a.ep_relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
          }
          if (ep_affine && orow < 0) {      // padding rows of the last tile must not enter the statistics
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (orow >= 0 && !B2M_ABLATE(a, 2)) {
            if (a.y32 != nullptr) {
              // fp32 rows of c_store real columns (class logits): scalar stores, the row pitch is not 16-byte aligned
              float* yr = a.y32 + (int64_t)orow * a.c_store;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = n0 + cc * 32 + j;
                if (j < cw && col < a.c_store) yr[col] = __uint_as_float(v[j]);
              }
            } else {
              uint16_t* yrow = a.y + rbase;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (q * 8 < cw) {
                  uint4 o;
                  __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    ob[e] = __floats2bfloat162_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
                  *reinterpret_cast<uint4*>(yrow + cc * 32 + q * 8) = o;
                }
              }
            }
          }
          if (a.colsum != nullptr && !B2M_ABLATE(a, 2)) {
            if (BNR) {
              // BatchNorm-backward reduction of the tensor this gradient belongs to: g = the bf16 value just stored, gated
              // by the producer's ReLU mask; first quantity g, second g * xhat (two staging rounds through the same tile)
              // g = the bf16 value just stored, gated; the column sums of g and of g * x go through the staging tile one
              // after the other (xhat = (x - mean) * invstd is applied to the SUMS when they leave the CTA)
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = (orow >= 0 && ((gate >> j) & 1u))
                           ? __float_as_uint(__bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j])))) : 0u;
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < cw) st_shared_f32(stage_a + (lane * kStagePitch + j) * 4, __uint_as_float(v[j]));
              __syncwarp();
              if (lane < cw) {
#pragma unroll
                for (int r = 0; r < 32; ++r) s1 += ld_shared_f32(stage_a + (r * kStagePitch + lane) * 4);
              }
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (q * 8 < cw) {
                  const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&xr[q]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(rb[e]);
                    st_shared_f32(stage_a + (lane * kStagePitch + q * 8 + 2 * e) * 4, __uint_as_float(v[q * 8 + 2 * e]) * f.x);
                    st_shared_f32(stage_a + (lane * kStagePitch + q * 8 + 2 * e + 1) * 4,
                                  __uint_as_float(v[q * 8 + 2 * e + 1]) * f.y);
                  }
                }
              }
              __syncwarp();
              if (lane < cw) {
#pragma unroll
                for (int r = 0; r < 32; ++r) s2 += ld_shared_f32(stage_a + (r * kStagePitch + lane) * 4);
                red_shared_f64(csum_a + (cc * 32 + lane) * 8, (double)s1);
                red_shared_f64(csum_a + (a.ntile + cc * 32 + lane) * 8, (double)s2);
              }
            } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < cw) st_shared_f32(stage_a + (lane * kStagePitch + j) * 4, __uint_as_float(v[j]));
            __syncwarp();
            if (lane < cw) {
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const float f = ld_shared_f32(stage_a + (r * kStagePitch + lane) * 4);
                s1 += f;
                s2 = fmaf(f, f, s2);
              }
              red_shared_f64(csum_a + (cc * 32 + lane) * 8, (double)s1);
              red_shared_f64(csum_a + (a.ntile + cc * 32 + lane) * 8, (double)s2);
            }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * par);
      if (warp == 0 && wi == 0) B2M_TRACE(31);
    }
    if (warp == 0) B2M_TRACE(32);
    if (a.colsum != nullptr && !split) {
      named_bar_sync(1, kEpiWarps * 32);   // all four epilogue warps have added their last tile
      const double* cs = reinterpret_cast<const double*>(smem + a.off_csum);
      if (BNR) {
        // (sum g, sum g * x) -> (sum g, sum g * xhat): xhat = (x - mean) * invstd
        const float* ep = reinterpret_cast<const float*>(smem + a.off_ep);
        for (int i = tid; i < a.ntile; i += kEpiWarps * 32) {
          atomicAdd(a.colsum + n0 + i, cs[i]);
          atomicAdd(a.colsum + a.c_n + n0 + i, (double)ep[a.ntile + i] * (cs[a.ntile + i] - (double)ep[i] * cs[i]));
        }
      } else {
      for (int i = tid; i < 2 * a.ntile; i += kEpiWarps * 32) {
        const int col = (i < a.ntile) ? (n0 + i) : (a.c_n + n0 + i - a.ntile);
        atomicAdd(a.colsum + col, cs[i]);
      }
      }
    }
    if (warp == 0) B2M_TRACE(33);
  } else if (warp == kFwdIssuer0 || warp == kFwdIssuer1) {
    // ================= MMA issuers: warp 4 owns tile 0 of every work item, warp 5 tile 1 (T == 2) =================
    // The single-thread issue loop is the critical path of this kernel (~90 instructions per stage), so the two
    // tiles of a work item, which accumulate into different TMEM columns, get an issuer each. The whole warp runs
    // the loop (warp-uniform values live in uniform registers); one elected lane issues MMAs and commits.
    const int me = (warp == kFwdIssuer1) ? 1 : 0;
    if (me < a.T) {
      const bool lead = elect_one_sync();
      const uint32_t idesc = umma_idesc_bf16(kTileM, a.ntile, 0, 0);
      const uint32_t hi128 = umma_desc_hi(1024, 128), hi64 = umma_desc_hi(512, 64), hi32 = umma_desc_hi(256, 32);
      const uint32_t hia = (KPACK == 2) ? hi64 : (KPACK == 4 ? hi32 : hi128);
      // Each tile of the work item has a PRIVATE ring of A slots (SA / T each) fed by its own producer warps, so an
      // issuer only ever meets its own stages (a parity wait is sound only for a waiter that sees every phase of a
      // barrier). The weight slices (B ring) are shared: both issuers wait for and release every slice.
      const int SAr = SA / a.T, abase = me * SAr;
      Ring ra, rb;
      ra.init(SAr); rb.init(SB);
      // KPACK == 1: the low descriptor words of the current A / B slots travel with the rings (no per-stage address
      // arithmetic on the issue path: every instruction of this loop is ~10 cycles of a lone warp's critical path).
      // Stage forms: every chunk is a full 64-wide SW128 tile (4 K-steps; a 16- or 48-wide tail is zero in both
      // operands) except the very last chunk of an offset when c_red % 64 == 32 (2 K-steps, SW64).
      const uint32_t a_ring_lo = umma_desc_lo(smem_base + abase * a.a_slot_bytes, 16);
      const uint32_t b_ring_lo = umma_desc_lo(smem_base + a.off_b, 16);
      const uint32_t a_step = (uint32_t)(a.a_slot_bytes >> 4), b_step = (uint32_t)(a.b_bytes >> 4);
      const uint32_t b_chunk = (uint32_t)((a.ntile * 128) >> 4);
      uint32_t a_cur = a_ring_lo, b_cur = b_ring_lo;
      const bool two = (KPACK == 1) && a.cps == 2;
      const int last_n = nch - (nst - 1) * ((KPACK == 1) ? a.cps : 1);   // chunks of an offset's last stage
      const bool has_rem = a.rem != 0;
      // which of the four stage forms the LAST stage of an offset has; every other stage is two (one) full chunks
      uint32_t last_c0_rem = (has_rem && last_n == 1) ? 1u : 0u, last_c1 = (two && last_n == 2) ? (has_rem ? 2u : 1u) : 0u;
      uint32_t st_a = a_step, st_b = b_step, st_c = b_chunk, st_last = (uint32_t)(nst - 1);
      // keep them in registers: re-deriving them from the kernel parameters costs constant-bank loads on the issue path
      asm volatile("" : "+r"(last_c0_rem), "+r"(last_c1), "+r"(st_a), "+r"(st_b), "+r"(st_c), "+r"(st_last));
#ifndef B2M_DEBUG_BUILD
      if (KPACK == 1 && a.mwords == 1 && !a.lean_off) {
        // ---- lean issue loop (kernel volumes <= 32: the offsets of a tile are ONE mask word) ----
        // ncu's instruction-level samples of the general loop below (profiles/r2_ncu_fwd_96_cpasync_issuer.txt): 37 % of
        // the issuer's time went into loop control (mask word selection, bit scans, kernel parameters re-read from the
        // constant bank), 17 % into the two barrier waits taken one after the other, 8 % into the proxy fence, 13 % into
        // the MMAs. Here: the issuer never needs the offset NUMBER (what a slot holds is the producers' business), only
        // whether its own tile takes part in each set bit of the work item's union mask; every loop-invariant lives in a
        // register; the waits on the weight slot and the A slot are issued back to back; the masks of the next work
        // item are loaded while this one runs; MMAs and commits of a stage are one straight-line block.
        const uint32_t nst_r = (uint32_t)nst, T_r = (uint32_t)a.T, SAr_r = (uint32_t)SAr, SB_r = (uint32_t)SB;
        const uint32_t two_f = two ? 1u : 0u, fence_f = (a.ldgsts == 1) ? 1u : 0u;
        const uint32_t af0 = a_full + 8 * abase, ae0 = a_empty + 8 * abase;
        const uint32_t slice_w = sl.w0;
        const int n_tiles_r = a.n_tiles, n_work_r = a.n_work, stride_r = (int)gridDim.x;
        const uint32_t* gm = a.gmask;
        const int64_t ngroups = (a.n_out + 63) / 64;
        const uint32_t colstride_r = (uint32_t)a.colstride;
        // masks of tiles (tile0, tile0 + 1) of a work item; the same values every other role derives (fwd_tile_mask)
        auto load_masks = [&](int w, uint32_t& q0, uint32_t& q1) {
          q0 = 0u; q1 = 0u;
          if (w >= n_work_r) return;
          const int t0 = w * (int)T_r;
          if (gm == nullptr) { q0 = (t0 < n_tiles_r) ? 1u : 0u; q1 = (T_r > 1 && t0 + 1 < n_tiles_r) ? 1u : 0u; return; }
          const int64_t g0 = 2 * (int64_t)t0;
          if (t0 < n_tiles_r) { q0 = __ldg(gm + g0); if (g0 + 1 < ngroups) q0 |= __ldg(gm + g0 + 1); }
          if (T_r > 1 && t0 + 1 < n_tiles_r) { q1 = __ldg(gm + g0 + 2); if (g0 + 3 < ngroups) q1 |= __ldg(gm + g0 + 3); }
          q0 &= slice_w; q1 &= slice_w;
        };
        uint32_t ra_slot = 0, ra_phase = 0, rb_slot = 0, rb_phase = 0;
        uint32_t a_fa = af0, a_ea = ae0, b_fa = b_full, b_ea = b_empty;
        uint32_t q0, q1;
        load_masks((int)blockIdx.x, q0, q1);
        uint32_t wi = 0;
        for (int w = blockIdx.x; w < n_work_r; w += stride_r, ++wi) {
          const uint32_t par = wi & 1u;
          uint32_t n0m, n1m;
          load_masks(w + stride_r, n0m, n1m);                      // in flight during this work item
          mbar_wait(acc_empty + 8 * par, ((wi >> 1) & 1u) ^ 1u, 2);
          tc_fence_after();
          const uint32_t mine = me ? q1 : q0;
          uint32_t mu = q0 | q1;
          const uint32_t d = tmem_base + (par * T_r + (uint32_t)me) * colstride_r;
          uint32_t acc = 0;
          while (mu) {
            const uint32_t bit = mu & (0u - mu);
            mu ^= bit;
            const bool has = (mine & bit) != 0u;
            for (uint32_t c = 0; c < nst_r; ++c) {
              if (has) {
                const uint32_t okb = mbar_try_wait(b_fa, rb_phase), oka = mbar_try_wait(a_fa, ra_phase);
                if (!okb) mbar_wait(b_fa, rb_phase, 4);
                if (!oka) mbar_wait(a_fa, ra_phase, 5);
                if (fence_f) fence_proxy_async();      // cp.async wrote the slot through the generic proxy
                tc_fence_after();
                if (lead) {
                  const bool last = (c == st_last);
                  umma_stage_bf16(d, a_cur, b_cur, st_c, idesc, acc, last ? last_c0_rem : 0u, last ? last_c1 : two_f,
                                  hi128, hi64, a_ea, b_ea);
                }
                acc = 1u;
                const bool wrap = (++ra_slot == SAr_r);
                ra_slot = wrap ? 0u : ra_slot;
                ra_phase ^= wrap ? 1u : 0u;
                a_cur = wrap ? a_ring_lo : a_cur + st_a;
                a_fa = wrap ? af0 : a_fa + 8;
                a_ea = wrap ? ae0 : a_ea + 8;
              } else {
                // the other tile uses this offset, this one does not: wait for the slice, then release it (see below)
                mbar_wait(b_fa, rb_phase, 8);
                if (lead) mbar_arrive(b_ea);
              }
              const bool wrapb = (++rb_slot == SB_r);
              rb_slot = wrapb ? 0u : rb_slot;
              rb_phase ^= wrapb ? 1u : 0u;
              b_cur = wrapb ? b_ring_lo : b_cur + st_b;
              b_fa = wrapb ? b_full : b_fa + 8;
              b_ea = wrapb ? b_empty : b_ea + 8;
            }
          }
          if (lead) umma_commit(acc_full + 8 * par);
          q0 = n0m; q1 = n1m;
        }
      } else
#endif
      {
      int wi = 0;
      int nstage = 0;      // debug trace only
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
        const int par = wi & 1;
        mbar_wait(acc_empty + 8 * par, ((uint32_t)(wi >> 1) & 1u) ^ 1u, 2);
        tc_fence_after();
        const MaskBits m0 = fwd_tile_mask(a, w * a.T, sl);
        const MaskBits m1 = (a.T > 1) ? fwd_tile_mask(a, w * a.T + 1, sl) : mask_zero();
        const MaskBits mu = mask_or(m0, m1);
        const uint32_t d = tmem_base + (uint32_t)((par * a.T + me) * a.colstride);
        uint32_t acc = 0;
        for (int kg = next_group<KPACK>(mu, 0, a.nkg, a.mwords); kg < a.nkg; kg = next_group<KPACK>(mu, kg + 1, a.nkg, a.mwords)) {
          const uint32_t s0 = mask_bits(m0, kg * KPACK, KPACK), s1 = mask_bits(m1, kg * KPACK, KPACK);
          const uint32_t sub = me ? s1 : s0;
          for (int c = 0; c < nst; ++c) {
            if (sub) {
              const int aslot = abase + ra.slot;
              if (me == 0 && nstage >= 32 && nstage < 44) B2M_TRACE(100 + (nstage - 32) * 4);
              B2M_ISSUER_WAIT(b_full + 8 * rb.slot, rb.phase, 4);
              if (me == 0 && nstage == 0) B2M_TRACE(19);
              if (me == 0 && nstage >= 32 && nstage < 44) B2M_TRACE(101 + (nstage - 32) * 4);
              B2M_ISSUER_WAIT(a_full + 8 * aslot, ra.phase, 5);
              if ((KPACK == 1 && a.ldgsts == 1) || (KPACK == 2 && a.ldgsts)) fence_proxy_async();   // cp.async wrote the slot through the generic proxy
              tc_fence_after();
              if (me == 0 && nstage >= 32 && nstage < 44) B2M_TRACE(102 + (nstage - 32) * 4);
              if (me == 0 && nstage < 16) B2M_TRACE(40 + nstage);
              if (me == 0 && (nstage & 15) == 0 && nstage < 256) B2M_TRACE(60 + (nstage >> 4));
              ++nstage;
              if (lead) {
                const uint32_t b_lo = (KPACK == 1) ? 0u : umma_desc_lo(smem_base + a.off_b + rb.slot * a.b_bytes, 16);
                const uint32_t a_lo = (KPACK == 1) ? 0u : umma_desc_lo(smem_base + aslot * a.a_slot_bytes, 16);
                if (KPACK == 1) {
                  // a stage holds one or two chunks of this offset: sub-tile j of the A slot (16 KB apart) against
                  // weight slice j of the B slot (ntile * 128 bytes apart)
                  const bool last = ((uint32_t)c == st_last);
#ifdef B2M_LEAN_ISSUER
                  if (!B2M_ABLATE(a, 1)) {
                    // the whole stage, MMAs and both commits, as one straight-line block (common.cuh)
                    umma_stage_bf16(d, a_cur, b_cur, st_c, idesc, acc, last ? last_c0_rem : 0u,
                                    last ? last_c1 : (two ? 1u : 0u), hi128, hi64, a_empty + 8 * aslot, b_empty + 8 * rb.slot);
                  } else {
                    umma_commit(a_empty + 8 * aslot);
                    umma_commit(b_empty + 8 * rb.slot);
                  }
#else
                  if (B2M_ABLATE(a, 1)) {
                  } else {
                    if (last && last_c0_rem) umma_chunk2(d, a_cur, b_cur, hi64, idesc, acc);
                    else umma_chunk4(d, a_cur, b_cur, hi128, idesc, acc);
                    const uint32_t f1 = last ? last_c1 : (two ? 1u : 0u);      // 0 = no second chunk, 1 = full, 2 = 32-wide
                    if (f1 == 1u) umma_chunk4(d, a_cur + (kASlotBytes >> 4), b_cur + st_c, hi128, idesc, 1u);
                    else if (f1 == 2u) umma_chunk2(d, a_cur + (kASlotBytes >> 4), b_cur + st_c, hi64, idesc, 1u);
                  }
#endif
                } else if (KPACK == 8) {
                  // 8 offsets x 8 channels side by side in one SW128 tile: a K=16 step covers a pair of offsets
                  uint32_t ac = acc;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    if ((sub >> (2 * ks)) & 3u) {
                      umma_bf16_lohi(d, a_lo + 2 * ks, hi128, b_lo + 2 * ks, hi128, idesc, ac);
                      ac = 1u;
                    }
                  }
                } else {
                  constexpr uint32_t wa = 128 / KPACK;               // sub-tile row bytes: 64 (SW64) or 32 (SW32)
                  uint32_t ac = acc;
#pragma unroll
                  for (int j = 0; j < KPACK; ++j) {
                    if ((sub >> j) & 1u) {
#pragma unroll
                      for (uint32_t ks = 0; ks < wa / 32; ++ks) {
                        umma_bf16_lohi(d, a_lo + ((j * kTileM * wa) >> 4) + 2 * ks, hia, b_lo + ((j * wa) >> 4) + 2 * ks, hi128, idesc, ac);
                        ac = 1u;
                      }
                    }
                  }
                }
#ifdef B2M_LEAN_ISSUER
                if (KPACK != 1)
#endif
                {
                  umma_commit(a_empty + 8 * aslot);
                  umma_commit(b_empty + 8 * rb.slot);          // arrives once this tile's MMAs on the slice are done
                }
              }
              if (me == 0 && nstage > 32 && nstage <= 44) B2M_TRACE(103 + (nstage - 33) * 4);
              acc = 1u;
              ra.next();
              a_cur = (ra.slot == 0) ? a_ring_lo : a_cur + st_a;
            } else {
              // The other tile uses this offset group, this one does not: release the B slot. Waiting for the slice
              // first keeps this warp from running a whole slot use ahead: b_empty counts arrivals, it cannot tell
              // two arrivals of one warp from one arrival of each.
              mbar_wait(b_full + 8 * rb.slot, rb.phase, 8);
              if (lead) mbar_arrive(b_empty + 8 * rb.slot);
            }
            rb.next();
            b_cur = (rb.slot == 0) ? b_ring_lo : b_cur + st_b;
          }
        }
        if (lead) umma_commit(acc_full + 8 * par);
        if (me == 0 && wi == 0) B2M_TRACE(21);
      }
      }
    }
    __syncwarp();
  } else if (warp == kFwdLoader) {
    // ================= B loader: bulk copies of pre-swizzled weight slices (warp-uniform, elected lane issues) ====
    const bool lead = elect_one_sync();
    Ring rb;
    rb.init(SB);
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      const MaskBits m0 = fwd_tile_mask(a, w * a.T, sl);
      const MaskBits m1 = (a.T > 1) ? fwd_tile_mask(a, w * a.T + 1, sl) : mask_zero();
      const MaskBits mu = mask_or(m0, m1);
      for (int kg = next_group<KPACK>(mu, 0, a.nkg, a.mwords); kg < a.nkg; kg = next_group<KPACK>(mu, kg + 1, a.nkg, a.mwords)) {
        for (int c = 0; c < nst; ++c) {
          mbar_wait(b_empty + 8 * rb.slot, rb.phase ^ 1u, 9);
          if (lead) {
            const uint32_t b_s = smem_base + a.off_b + rb.slot * a.b_bytes;
            const int per = (KPACK == 1) ? a.cps : 1;            // weight slices of this stage
            const int ch0 = c * per, ch1 = min(nch, ch0 + per);
            uint32_t total = 0;
            for (int ch = ch0; ch < ch1; ++ch) total += (uint32_t)(a.ntile * ((ch < a.nfull) ? 128 : 64));
            if (B2M_ABLATE(a, 3)) {
              mbar_arrive(b_full + 8 * rb.slot);
            } else {
              mbar_arrive_expect_tx(b_full + 8 * rb.slot, total);
              for (int ch = ch0; ch < ch1; ++ch) {
                const int wb = (ch < a.nfull) ? 128 : 64;
                const uint8_t* src = a.w + (int64_t)kg * a.kg_bytes + (int64_t)min(ch, a.nfull) * a.c_n * 128 + (int64_t)n0 * wb;
                bulk_g2s(b_s + (uint32_t)((ch - ch0) * a.ntile * 128), src, (uint32_t)(a.ntile * wb), b_full + 8 * rb.slot);
              }
            }
          }
          rb.next();
        }
      }
    }
    __syncwarp();
  } else if (warp != kFwdIdle) {
    // ================= gather warps: A stage s is produced by warp s % kFwdProd =================
    // Stage s is produced by warp s % np with np <= SA: a warp then never runs more than one use of a slot ahead
    // of the consumer, which is what waiting on an mbarrier phase PARITY requires.
    // With T == 2 the warps are split between the two tiles' private rings (see the MMA issuers). Within a ring,
    // stage s is produced by warp s % np with np <= ring size: a warp then never runs more than one use of a slot
    // ahead of the consumer, which is what waiting on an mbarrier phase PARITY requires.
    // KPACK == 1: the a.cps chunks of a stage are gathered by a.cps warps (a producer group), one chunk each, so that
    // the ~1.2 us a warp needs to issue its 32 gather4s does not grow with the stage.
    const int NS = (KPACK == 1) ? a.cps : 1;
    const int gi = fwd_gather_index(warp);
    const int pw = gi / NS;                                     // producer group
    const int part = gi % NS;                                   // the chunk of the stage this warp gathers
    const int groups = (kFwdProd / NS) / a.T;                   // producer groups per ring
    const int t = (a.T > 1 && pw >= groups) ? 1 : 0;            // the tile (ring) this warp feeds
    const int p = pw - t * groups;
    const int SAr = SA / a.T, abase = t * SAr;
    const int np = min(groups, SAr);
    const int q_lo = 0, q_hi = 32;                               // every warp gathers all 32 row quads of the tile
    Ring ra;
    ra.init(SAr);
    int turn = 0;  // stage counter modulo np
    const bool has_work = (p < np && pw < groups * a.T);   // spare warps have no stages
    if (KPACK == 1) {
      // 64-wide-chunk mode, software-pipelined: the index quad of this warp's NEXT stage is loaded before the current
      // stage waits for its slot, so the ~1 us index-load latency is off the slot's turn-around path (with every unit
      // of real work removed the kernel was still bound by that loop: ablation in DESIGN.md section 3).
      struct St { int w, kg, c, slot; uint32_t phase; bool valid; };
      int w = has_work ? (int)blockIdx.x : a.n_work;
      MaskBits mt = mask_zero();
      int kg = a.nkg, d = 0;
      Ring ra0 = ra;
      bool fresh = true;                 // work item w not entered yet
      auto next = [&]() -> St {
        St s; s.valid = false; s.w = 0; s.kg = 0; s.c = 0; s.slot = 0; s.phase = 0;
        while (w < a.n_work) {
          if (fresh) {
            mt = fwd_tile_mask(a, w * a.T + t, sl);
            kg = next_group<1>(mt, 0, a.nkg, a.mwords);
            d = p - turn; if (d < 0) d += np;
            fresh = false;
          }
          if (kg < a.nkg) {
            if (d < nst) {
              const Ring rs = ra0.at(d);
              s.valid = true; s.w = w; s.kg = kg; s.c = d; s.slot = abase + rs.slot; s.phase = rs.phase;
              d += np;
              return s;
            }
            ra0 = ra0.at(nst);
            turn += nst; while (turn >= np) turn -= np;
            kg = next_group<1>(mt, kg + 1, a.nkg, a.mwords);
            d = p - turn; if (d < 0) d += np;
            continue;
          }
          w += gridDim.x;
          fresh = true;
        }
        return s;
      };
      const int quad = q_lo + lane;                         // this lane gathers rows 4*quad .. 4*quad+3 of the tile
      const bool on = quad < q_hi;
      auto load_idx = [&](const St& s) -> int4 {
        int4 idx = make_int4(-1, -1, -1, -1);
        if (s.valid && on) {
          const int64_t pos0 = (int64_t)(s.w * a.T + t) * kTileM + 4 * quad;
          if (a.nbr) {
            idx = ld_nc_int4(a.nbr + (int64_t)s.kg * a.n_pitch + pos0);
          } else {
            idx.x = pos0 < a.n_out ? (int)pos0 : -1;
            idx.y = pos0 + 1 < a.n_out ? (int)pos0 + 1 : -1;
            idx.z = pos0 + 2 < a.n_out ? (int)pos0 + 2 : -1;
            idx.w = pos0 + 3 < a.n_out ? (int)pos0 + 3 : -1;
          }
        }
        return idx;
      };
      St cur = next();
      if (a.ldgsts) {
        // cp.async mode: lane l holds the row indices of tile rows l, 32 + l, 64 + l, 96 + l (four coalesced 128-byte
        // loads per warp), loaded one stage ahead like the index quads of the TMA mode
        auto load_rows = [&](const St& s, int (&R)[4]) {
#pragma unroll
          for (int u = 0; u < 4; ++u) R[u] = -1;
          if (s.valid) {
            const int64_t pos0 = (int64_t)(s.w * a.T + t) * kTileM + lane;
            if (a.nbr) {
              const int32_t* pr = a.nbr + (int64_t)s.kg * a.n_pitch + pos0;
#pragma unroll
              for (int u = 0; u < 4; ++u) R[u] = __ldg(pr + 32 * u);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) R[u] = (pos0 + 32 * u < a.n_out) ? (int)(pos0 + 32 * u) : -1;
            }
          }
        };
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(a.x);
        const uint32_t row_bytes = (uint32_t)a.c_red * 2u;
        int R[4];
        load_rows(cur, R);
        while (cur.valid) {
          const St nxt = next();
          int Rn[4];
          load_rows(nxt, Rn);                                  // in flight while this stage waits for its slot
          const int ch = cur.c * NS + part;                    // this warp's chunk of the stage (may not exist)
          const uint32_t a_s = smem_base + cur.slot * a.a_slot_bytes + part * kASlotBytes;
          const uint32_t full = a_full + 8 * cur.slot;
          mbar_wait(a_empty + 8 * cur.slot, cur.phase ^ 1u, 10);
          if (ch < nch && !B2M_ABLATE(a, 0)) {
            if (ch < a.nfull) {
              const int left = (a.c_red - ch * 64) >> 3;       // 16-byte pieces of the tensor in this chunk
              ldgsts_rows_sw128<128>(a_s, xb + ch * 128, row_bytes, R, lane, left < 8 ? left : 8);
            } else {
              ldgsts_rows_sw64<128>(a_s, xb + ch * 128, row_bytes, R, lane);
            }
          }
          if (a.ldgsts == 1) {
            cp_async_mbar_arrive_noinc(full);                  // every lane: the barrier counts 32 arrivals per warp
          } else {
            // mode 2: the warp waits for its copies, makes them visible to the async proxy itself and arrives once, so
            // that the MMA issuer (the critical path) needs no proxy fence; the ring depth still bounds what is in
            // flight because there are as many producer groups as slots
            cp_async_wait_all();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) R[u] = Rn[u];
          cur = nxt;
        }
      }
      int4 idx = load_idx(cur);
      int ntr = 0;       // debug trace only
      while (cur.valid) {
        const St nxt = next();
        const int4 idx_next = load_idx(nxt);               // in flight while this stage waits for its slot
        const int ch = cur.c * NS + part;                      // this warp's chunk of the stage (may not exist)
        const bool has = ch < nch;
        const uint32_t a_s = smem_base + cur.slot * a.a_slot_bytes + part * kASlotBytes;
        const uint32_t full = a_full + 8 * cur.slot;
        const uint32_t wc = (ch < a.nfull) ? 128u : 64u;
        mbar_wait(a_empty + 8 * cur.slot, cur.phase ^ 1u, 10);
        if (pw == 0 && ntr < 4) B2M_TRACE(10 + 2 * ntr);
        if (lane == 0) {
          if (B2M_ABLATE(a, 0) || !has) mbar_arrive(full); else mbar_arrive_expect_tx(full, (uint32_t)(q_hi - q_lo) * 4u * wc);
        }
        __syncwarp();
        if (on && has && !B2M_ABLATE(a, 0))
          tma_gather4(a_s + quad * 4 * wc, (ch < a.nfull) ? &tm_main : &tm_rem, full, ch * 64, idx.x, idx.y, idx.z, idx.w);
        __syncwarp();
        if (pw == 0 && ntr < 4) B2M_TRACE(11 + 2 * ntr);
        ++ntr;
        cur = nxt;
        idx = idx_next;
      }
    } else
    for (int w = has_work ? (int)blockIdx.x : a.n_work; w < a.n_work; w += gridDim.x) {
      const MaskBits mt = fwd_tile_mask(a, w * a.T + t, sl);
      for (int kg = next_group<KPACK>(mt, 0, a.nkg, a.mwords); kg < a.nkg; kg = next_group<KPACK>(mt, kg + 1, a.nkg, a.mwords)) {
        const uint32_t sub = mask_bits(mt, kg * KPACK, KPACK);
        // jump straight to the chunks of this offset group whose turn is this warp's (stage index = turn + d)
        const int n_kg = nch;
        int d0 = p - turn;
        if (d0 < 0) d0 += np;
        const Ring ra0 = ra;
        for (int d = d0; d < n_kg; d += np) {
            const int c = d;
            const Ring rs = ra0.at(d);
            const int aslot = abase + rs.slot;
            {
              const int quad = (KPACK == 1) ? q_lo + lane : lane;                 // this lane gathers rows 4*quad .. 4*quad+3
              const int64_t pos0 = (int64_t)(w * a.T + t) * kTileM + 4 * quad;
              const uint32_t a_s = smem_base + aslot * a.a_slot_bytes;
              const uint32_t full = a_full + 8 * aslot;
              if (KPACK == 1) {
                const bool on = quad < q_hi;
                int4 idx = make_int4(-1, -1, -1, -1);
                if (on) {
                  if (a.nbr) {
                    idx = ld_nc_int4(a.nbr + (int64_t)kg * a.n_pitch + pos0);
                  } else {
                    idx.x = pos0 < a.n_out ? (int)pos0 : -1;
                    idx.y = pos0 + 1 < a.n_out ? (int)pos0 + 1 : -1;
                    idx.z = pos0 + 2 < a.n_out ? (int)pos0 + 2 : -1;
                    idx.w = pos0 + 3 < a.n_out ? (int)pos0 + 3 : -1;
                  }
                }
                const uint32_t wc = (c < a.nfull) ? 128u : 64u;
                mbar_wait(a_empty + 8 * aslot, rs.phase ^ 1u, 10);
                if (lane == 0) {
                  if (B2M_ABLATE(a, 0)) mbar_arrive(full); else mbar_arrive_expect_tx(full, (uint32_t)(q_hi - q_lo) * 4u * wc);
                }
                __syncwarp();
                if (on && !B2M_ABLATE(a, 0)) tma_gather4(a_s + quad * 4 * wc, (c < a.nfull) ? &tm_main : &tm_rem, full, c * 64, idx.x, idx.y, idx.z, idx.w);
              } else if (KPACK == 8) {
                // cp.async path: lane = (offset j of the group, row quad rq); 8 x (one 16-byte index load + 4 copies)
                const int j = lane & 7, rq = lane >> 3;
                const int k = kg * 8 + j;
                const bool on = ((sub >> j) & 1u) != 0;
                const int32_t* nb = a.nbr ? a.nbr + (int64_t)k * a.n_pitch + (int64_t)(w * a.T + t) * kTileM : nullptr;
                mbar_wait(a_empty + 8 * aslot, rs.phase ^ 1u);
#pragma unroll 2
                for (int it = 0; it < 8; ++it) {
                  const int r0 = 16 * it + 4 * rq;
                  int4 idx = make_int4(-1, -1, -1, -1);
                  if (on) {
                    if (nb) {
                      idx = ld_nc_int4(nb + r0);
                    } else {
                      const int64_t q0 = (int64_t)(w * a.T + t) * kTileM + r0;
                      idx.x = q0 < a.n_out ? (int)q0 : -1;
                      idx.y = q0 + 1 < a.n_out ? (int)q0 + 1 : -1;
                      idx.z = q0 + 2 < a.n_out ? (int)q0 + 2 : -1;
                      idx.w = q0 + 3 < a.n_out ? (int)q0 + 3 : -1;
                    }
                  }
                  const int iv[4] = {idx.x, idx.y, idx.z, idx.w};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int row = r0 + u;
                    const uint32_t dst = a_s + row * 128 + ((j ^ (row & 7)) << 4);
                    if (iv[u] >= 0) cp_async16(dst, a.x + (int64_t)iv[u] * 8, 16u);
                    else st_shared_zero16(dst);
                  }
                }
                cp_async_wait_all();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
              } else if (KPACK == 2 && a.ldgsts) {
                // 32-channel rows (64 bytes) by cp.async: 4 lanes per row, 8 rows per instruction, 16 instructions per
                // offset instead of 32 gather4s at ~76 issue cycles each (wgrad with 32-channel operands went from 0.213
                // to 0.082 ms that way). Lane l holds the neighbour indices of rows l, l + 32, l + 64, l + 96.
                const int64_t tile0 = (int64_t)(w * a.T + t) * kTileM;
                int R[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int64_t pos = tile0 + 32 * u + lane;
                    if (!((sub >> j) & 1u)) R[j][u] = -1;
                    else if (a.nbr) R[j][u] = __ldg(a.nbr + (int64_t)(kg * 2 + j) * a.n_pitch + pos);
                    else R[j][u] = pos < a.n_out ? (int)pos : -1;
                  }
                }
                mbar_wait(a_empty + 8 * aslot, rs.phase ^ 1u);
                const uint8_t* xb = reinterpret_cast<const uint8_t*>(a.x);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  if ((sub >> j) & 1u) ldgsts_rows_sw64<128>(a_s + j * kTileM * 64, xb, 64u, R[j], lane);
                cp_async_mbar_arrive_noinc(full);                  // every lane: the barrier counts 32 arrivals
              } else {
                constexpr uint32_t wa = 128 / KPACK;
                int4 idx[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < KPACK && ((sub >> j) & 1u)) {
                    if (a.nbr) {
                      idx[j] = ld_nc_int4(a.nbr + (int64_t)(kg * KPACK + j) * a.n_pitch + pos0);
                    } else {  // identity map (kvol == 1): only j == 0 is ever set
                      idx[j].x = pos0 < a.n_out ? (int)pos0 : -1;
                      idx[j].y = pos0 + 1 < a.n_out ? (int)pos0 + 1 : -1;
                      idx[j].z = pos0 + 2 < a.n_out ? (int)pos0 + 2 : -1;
                      idx[j].w = pos0 + 3 < a.n_out ? (int)pos0 + 3 : -1;
                    }
                  }
                mbar_wait(a_empty + 8 * aslot, rs.phase ^ 1u);
                if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)__popc(sub) * kTileM * wa);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < KPACK && ((sub >> j) & 1u))
                    tma_gather4(a_s + j * kTileM * wa + lane * 4 * wa, &tm_main, full, 0, idx[j].x, idx[j].y, idx[j].z, idx[j].w);
              }
            }
        }
        ra = ra0.at(n_kg);
        turn += n_kg;
        while (turn >= np) turn -= np;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) B2M_TRACE(2);
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize of an offset-split convolution
// ------------------------------------------------------------------------------------------------
// v[row][col] = sum over slices z (ascending: deterministic) of part[z][row][col]; then the same fused epilogue as
// conv_fwd_kernel (scale/shift, residual, ReLU), bf16 (or fp32) store and the per-column sum / sum of squares of v.
// Thread (rl, g) owns the 8 columns of group g for rows rl, rl + rows-per-pass, ... (coalesced 32-byte fp32 pieces).
constexpr int kFinThreads = 256;
__global__ void __launch_bounds__(kFinThreads)
conv_finalize_kernel(const float* __restrict__ part, int nslices, int64_t n, int c, const float* __restrict__ scale,
                     const float* __restrict__ shift, const uint16_t* __restrict__ residual, int relu,
                     uint16_t* __restrict__ y, float* __restrict__ y32, int c_store, double* __restrict__ colsum,
                     const uint16_t* __restrict__ bnr_x, const uint8_t* __restrict__ bnr_mask,
                     const float* __restrict__ bnr_mean, const float* __restrict__ bnr_invstd) {
  extern __shared__ float fin_sh[];   // [rows per pass][c][2]
  const int G = c / 8;
  const int rpp = kFinThreads / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  float s1[8], s2[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[i] = 0.f; s2[i] = 0.f;
    // (with bnr_x - the fused BatchNorm-backward reduction, see conv_fwd_kernel - they hold mean / invstd instead)
    sc[i] = bnr_x ? bnr_mean[g * 8 + i] : (scale ? scale[g * 8 + i] : 1.f);
    sh[i] = bnr_x ? bnr_invstd[g * 8 + i] : (shift ? shift[g * 8 + i] : 0.f);
  }
  const bool affine = !bnr_x && (scale != nullptr || shift != nullptr);
  if (rl < rpp) {
    for (int64_t r = (int64_t)blockIdx.x * rpp + rl; r < n; r += (int64_t)gridDim.x * rpp) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll 4
      for (int z = 0; z < nslices; ++z) {
        const float4* p = reinterpret_cast<const float4*>(part + ((int64_t)z * n + r) * c + g * 8);
        const float4 a0 = __ldg(p), a1 = __ldg(p + 1);
        v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
        v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
      }
      if (affine) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
      }
      if (residual) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(residual + r * c) + g);
        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(rb[e]);
          v[2 * e] += f.x; v[2 * e + 1] += f.y;
        }
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (bnr_x) {
        const uint4 xx = __ldg(reinterpret_cast<const uint4*>(bnr_x + r * c) + g);
        const unsigned gate = bnr_mask ? __ldg(bnr_mask + r * G + g) : 0xFFu;
        const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&xx);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(xb[e]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 2 * e + h;
            const float gq = ((gate >> i) & 1u) ? __bfloat162float(__float2bfloat16_rn(v[i])) : 0.f;
            s1[i] += gq;
            s2[i] = fmaf(gq, ((h ? f.y : f.x) - sc[i]) * sh[i], s2[i]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1[i] += v[i]; s2[i] = fmaf(v[i], v[i], s2[i]); }
      }
      if (y32) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (g * 8 + i < c_store) y32[r * c_store + g * 8 + i] = v[i];
      } else {
        uint4 o;
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        reinterpret_cast<uint4*>(y + r * c)[g] = o;
      }
    }
  }
  if (colsum == nullptr) return;
  if (rl < rpp) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      fin_sh[(rl * c + g * 8 + i) * 2] = s1[i];
      fin_sh[(rl * c + g * 8 + i) * 2 + 1] = s2[i];
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < c; col += kFinThreads) {
    float t1 = 0.f, t2 = 0.f;
    for (int r = 0; r < rpp; ++r) { t1 += fin_sh[(r * c + col) * 2]; t2 += fin_sh[(r * c + col) * 2 + 1]; }
    atomicAdd(colsum + col, (double)t1);
    atomicAdd(colsum + c + col, (double)t2);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad kernel
// ------------------------------------------------------------------------------------------------
constexpr int kWgEpi = 4;                                       // warps 0-3: epilogue (TMEM lane quarters)
#ifdef B2M_WG_REMAP
// experiment: 20 warps, no gather warp on the issuer's scheduler sub-partition (warps 8, 12, 16 idle), like the forward kernel
constexpr int kWgAProd = 8;
constexpr int kWgBProd = 4;
constexpr int kWgThreads = 640;
#define B2M_WG_SW128 ldgsts_rows_sw128
#define B2M_WG_SW64 ldgsts_rows_sw64
#else
constexpr int kWgAProd = 10;                                    // gather warps for the X operand
constexpr int kWgBProd = 4;                                     // gather warps for the dY operand
constexpr int kWgThreads = (kWgEpi + 1 + kWgBProd + kWgAProd) * 32;   // 608
#define B2M_WG_SW128 ldgsts_rows_sw128_rolled
#define B2M_WG_SW64 ldgsts_rows_sw64_rolled
#endif
constexpr int kWgRows = 64;
constexpr int kWgASlotBytes = 128 * kWgRows * 2;                // 16 KB: M = 128 x 64 reduction rows

struct WgArgs {
  const uint16_t* x; const uint16_t* dy; const int32_t* nbr; const int32_t* order; const uint32_t* gmask; float* dw;
  int64_t n_out, n_pitch;
  int c_in, c_out, kvol, mwords, cpad, pk, G, colstride, groups_per_cta;
  int wa, nab;    // X operand: row bytes of a block (128 / 64 / 32) and blocks per A stage (nab * wa / 2 == 128 M rows)
  int wb, nbb;    // dY operand: row bytes of a block and number of column blocks
  int a_slots, b_slots, b_bytes, off_b, off_bars, tmem_cols;
  int store;      // 1: one CTA owns each dW element (no row splits): plain stores, zeros for offsets without pairs
  int lda, ldb;   // X / dY rows fetched by cp.async from every lane of the gather warps instead of TMA gather4
  int nwa, nwb;   // warps per producer group of the X / dY operand (2 in cp.async mode when the stage has >= 2 blocks)
  int ncols;      // gridDim.x: accumulator q of column `col` holds offset slot col + ncols * q (interleaved assignment)
  float* part;    // != nullptr: row split y writes its partial dW to part[y][kvol * c_in * c_out] with plain stores
                  // (summed in a fixed order by wgrad_reduce_kernel: deterministic, no contended atomics)
};

template <int ROWS>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy, const WgArgs a) {
  constexpr int kSlotA = 128 * ROWS * 2;     // bytes of an A slot: M = 128 channels x ROWS reduction rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SA = a.a_slots, SB = a.b_slots;
  const uint32_t bars = smem_base + a.off_bars;
  const uint32_t a_full = bars, a_empty = bars + 8 * SA;
  const uint32_t b_full = bars + 16 * SA, b_empty = b_full + 8 * SB;
  const uint32_t accum_bar = b_empty + 8 * SB;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + a.off_bars + 16 * SA + 16 * SB + 16);
  uint32_t* used_s = tmem_ptr_s + 1;

  const int col = blockIdx.x;             // which G offset groups
  const int mt = blockIdx.z;              // M tile (input-channel block of 128) when c_in > 128
  const int64_t total_groups = (a.n_out + ROWS - 1) / ROWS;
  // Row groups are dealt round-robin to the gridDim.y CTAs of a column: all CTAs then sweep the arrays together, so
  // the rows one CTA gathers (and the dY stage its column neighbours re-read) are still in L2 when the others need them.
  // With a contiguous range per CTA the 24 x 6 CTAs of a full-resolution launch streamed 24 distant regions and
  // drifted apart (L2 hit rate 57 %, 2 GB of DRAM reads per launch against 0.5 GB of operands).
  const int64_t g_begin = (int64_t)blockIdx.y;
  const int64_t g_end = total_groups;
  const int64_t g_step = (int64_t)gridDim.y;
  // Which offset slots (slot = pk consecutive kernel offsets) this CTA's accumulators hold. The work of a slot is the
  // number of row groups in which it occurs, and that differs a lot between offsets (the centre offset occurs in every
  // group, a corner offset in 16-40 % of them; the sorted order even breaks the k <-> 26 - k symmetry): with contiguous
  // runs of 5 offsets per column the busiest column of a ScanNet-shape 3^3 map had 1.51x the mean work (interleaved
  // assignment: 1.32x), i.e. a third of the launch was CTAs waiting for the slowest column, and the columns drifted
  // apart while sweeping the rows (ncu: L2 hit rate 51 %, 1.66 GB of DRAM reads for 0.47 GB of operands). So every CTA
  // counts the groups per slot from the map's group masks (a 76 KB read for 1.2 M rows) and runs the same deterministic
  // longest-processing-time assignment of slots to columns (<= G per column): 1.06x on that map.
  const int kslots_all = (a.kvol + a.pk - 1) / a.pk;
  __shared__ int s_cnt[128];
  __shared__ int s_koff[16];
  __shared__ int s_nq;
  // (small launches are latency bound, not balance bound, and the serial assignment below costs tens of microseconds
  // for 27 slots in 14 columns: they take the interleaved assignment slot = col + ncols * q)
  const bool balance = a.gmask != nullptr && a.ncols > 1 && a.n_out >= 32768 && kslots_all <= 32 && a.ncols <= 32;
  if (balance) {
    for (int i = tid; i < 128; i += kWgThreads) s_cnt[i] = 0;
    __syncthreads();
    const int64_t ng = (a.n_out + 63) / 64;
    const uint32_t pkb = (a.pk >= 32) ? 0xFFFFFFFFu : ((1u << a.pk) - 1u);
    for (int64_t g = tid; g < ng; g += kWgThreads) {
      const MaskBits m = mask_load(a.gmask, g, a.mwords);
      for (int sl = 0; sl < kslots_all; ++sl) {
        const int k0 = sl * a.pk;
        if ((mask_word(m, k0 >> 5) >> (k0 & 31)) & pkb) atomicAdd(&s_cnt[sl], 1);
      }
    }
    __syncthreads();
    if (warp == 0) {
      // longest-processing-time assignment by one warp: lane = slot for the arg-max, lane = column for the arg-min
      int cnt = lane < kslots_all ? s_cnt[lane] : -1;
      int load = 0, fill = 0, nq_mine = 0;
      for (int it = 0; it < kslots_all; ++it) {
        const int best = __reduce_max_sync(0xFFFFFFFFu, cnt < 0 ? -1 : ((cnt << 5) | (31 - lane)));   // ties: lowest slot
        const int slot = 31 - (best & 31), bc = best >> 5;
        if (lane == slot) cnt = -1;
        const int tgt = __reduce_min_sync(0xFFFFFFFFu, (lane < a.ncols && fill < a.G) ? ((load << 5) | lane) : 0x7FFFFFFF) & 31;
        if (lane == tgt) { load += bc; fill += 1; }
        if (tgt == col) { if (lane == 0) s_koff[nq_mine] = slot; ++nq_mine; }
      }
      if (lane == 0) s_nq = nq_mine;
    }
  } else if (tid == 0) {
    int nq_mine = 0;
    for (int sl = col; sl < kslots_all; sl += a.ncols) s_koff[nq_mine++] = sl;
    s_nq = nq_mine;
  }
  __syncthreads();
  const int nq = s_nq;                       // accumulators of this CTA that hold real offsets (<= G)
  auto k_first = [&](int q) -> int { return s_koff[q] * a.pk; };   // first kernel offset of accumulator q
  if (warp == 0) B2M_TRACE(0);

  if (warp == kWgEpi && lane == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(a_full + 8 * s, a.lda ? 32 * a.nwa : a.nwa); mbar_init(a_empty + 8 * s, 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(b_full + 8 * s, a.ldb ? 32 * a.nwb : a.nwb); mbar_init(b_empty + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == kWgEpi + 1) { tmem_alloc(smem_u32(tmem_ptr_s), (uint32_t)a.tmem_cols); tmem_relinquish(); }
  if (warp == kWgEpi + 1 + kWgBProd && lane == 0) { tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_dy); }
  if (tid == 0) *used_s = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (warp == 0) B2M_TRACE(1);
#ifdef B2M_WG_REMAP
  const int gidx = (warp >= 5 && (warp & 3)) ? fwd_gather_index(warp) : -1;     // 0 .. 11 over the warps off sub-partition 0
  const bool is_bprod = gidx >= 0 && gidx < kWgBProd, is_aprod = gidx >= kWgBProd;
  const int bidx = gidx, aidx = gidx - kWgBProd;
#else
  const bool is_bprod = warp > kWgEpi && warp <= kWgEpi + kWgBProd, is_aprod = warp > kWgEpi + kWgBProd;
  const int bidx = warp - (kWgEpi + 1), aidx = warp - (kWgEpi + 1 + kWgBProd);
#endif

  // offsets present in reduction group g (ROWS rows = ROWS / 64 groups of the sorted map's masks)
  const int64_t ngroups64 = (a.n_out + 63) / 64;
  auto group_mask = [&](int64_t g) -> MaskBits {
    if (a.gmask == nullptr) { MaskBits m = mask_zero(); m.w0 = 1u; return m; }
    MaskBits m = mask_load(a.gmask, g * (ROWS / 64), a.mwords);
    if (ROWS == 128 && 2 * g + 1 < ngroups64) m = mask_or(m, mask_load(a.gmask, 2 * g + 1, a.mwords));
    return m;
  };
  // bit q set iff accumulator q (offsets k_first(q) .. + pk - 1) has a pair in a row group with mask m
  const uint32_t pk_bits = (a.pk >= 32) ? 0xFFFFFFFFu : ((1u << a.pk) - 1u);
  auto acc_mask = [&](const MaskBits& m) -> uint32_t {
    if (a.gmask == nullptr) return 1u;
    uint32_t out = 0;
    if (a.mwords == 1) {                   // kernel volumes <= 32: one mask word
      for (int q = 0; q < nq; ++q) out |= (((m.w0 >> k_first(q)) & pk_bits) != 0u ? 1u : 0u) << q;
      return out;
    }
    for (int q = 0; q < nq; ++q) {         // pk in {1,2,4,8,16} divides 32 and k_first % pk == 0: one word per accumulator
      const int k0 = k_first(q);
      out |= (((mask_word(m, k0 >> 5) >> (k0 & 31)) & pk_bits) != 0u ? 1u : 0u) << q;
    }
    return out;
  };

  if (warp < kWgEpi) {
    // ================= epilogue: TMEM -> fp32 vector atomics into dw =================
    mbar_wait(accum_bar, 0, 20);
    tc_fence_after();
    if (warp == 0) B2M_TRACE(30);
    const uint32_t used = *reinterpret_cast<volatile uint32_t*>(used_s);
    // A lane holds 32 consecutive output columns of ONE M row (TMEM lane == M row == input channel), and rows of dw
    // are c_out floats apart: atomics issued straight from that layout touch 32 different 128-byte lines per warp
    // instruction (measured: 40 us for the 2 x 128 x 256 accumulators of a 256->256 layer, the whole cost of a
    // deep-level launch). Each 32x32 tile is therefore transposed through shared memory (the A ring is idle: every
    // stage has been consumed once accum_bar completes) so that 8 lanes cover the 128 contiguous bytes of a row and
    // a warp instruction touches 4 full lines.
    const uint32_t tr = smem_base + warp * (32 * kStagePitch * 4);
    const int r_sub = lane >> 3, c4 = (lane & 7) * 4;
    const int nchunk32 = (a.c_out + 31) / 32;
    const bool plain = a.store || a.part != nullptr;     // this CTA owns the elements it writes
    float* dwo = a.part ? a.part + (int64_t)blockIdx.y * ((int64_t)a.kvol * a.c_in * a.c_out) : a.dw;
    for (int q = 0; q < nq; ++q) {
      const bool has = ((used >> q) & 1u) != 0;
      if (!has && !plain) continue;
      for (int cc = 0; cc < nchunk32; ++cc) {
        const int cw = min(32, a.c_out - cc * 32);
        uint32_t v[32];
        if (has) {
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * a.colstride + cc * 32);
          if (cw == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < cw) st_shared_f32(tr + (lane * kStagePitch + j) * 4, __uint_as_float(v[j]));
        __syncwarp();
        if (c4 < cw) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = 4 * j + r_sub;                  // row of the 32x32 tile
            const int mrow = warp * 32 + r;               // M row
            const int slot = mrow / a.cpad;               // which packed offset
            const int ci = mrow % a.cpad + mt * 128;
            const int k = k_first(q) + slot;
            if (slot < a.pk && k < a.kvol && ci < a.c_in) {
              const uint32_t src = tr + (r * kStagePitch + c4) * 4;
              const float4 val = make_float4(ld_shared_f32(src), ld_shared_f32(src + 4), ld_shared_f32(src + 8),
                                             ld_shared_f32(src + 12));
              float4* dst = reinterpret_cast<float4*>(dwo + ((int64_t)k * a.c_in + ci) * a.c_out + cc * 32 + c4);
              if (plain) *dst = val; else atomicAdd(dst, val);
            }
          }
        }
        __syncwarp();
      }
    }
    if (warp == 0) B2M_TRACE(31);
  } else if (warp == kWgEpi) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues; see conv_fwd_kernel) =================
    const bool lead = elect_one_sync();
    const uint32_t idesc = umma_idesc_bf16(128, a.c_out, 1, 1);
    const uint32_t wa = (uint32_t)a.wa, wb = (uint32_t)a.wb;
    // MN-major: LBO = stride between blocks (64 rows x row bytes each), SBO = stride between 8-row K groups
    const uint32_t hi_a = umma_desc_hi(8 * wa, wa), hi_b = umma_desc_hi(8 * wb, wb);
    const uint32_t lbo_a = ((ROWS * wa) >> 4) << 16, lbo_b = ((ROWS * wb) >> 4) << 16;
    const uint32_t kstep_a = (16 * wa) >> 4, kstep_b = (16 * wb) >> 4;     // descriptor units per K=16 step
    uint32_t used = 0;
    Ring ra, rb;
    ra.init(SA); rb.init(SB);
    MaskBits m_next = group_mask(g_begin < g_end ? g_begin : 0);
    for (int64_t g = g_begin; g < g_end; g += g_step) {
      const MaskBits m = m_next;
      if (g + g_step < g_end) m_next = group_mask(g + g_step);   // prefetched: the load latency overlaps this group's work
      const uint32_t qm = acc_mask(m);
      if (!qm) continue;
      mbar_wait(b_full + 8 * rb.slot, rb.phase, 21);
      if (a.ldb) fence_proxy_async();      // cp.async wrote the slot through the generic proxy
      if (g == g_begin) B2M_TRACE(19);
      const uint32_t b_lo = (((smem_base + a.off_b + rb.slot * a.b_bytes) >> 4) & 0x3FFFu) | lbo_b;
#pragma unroll 1
      for (uint32_t rem = qm; rem; rem &= rem - 1) {
        const int q = __ffs(rem) - 1;
        mbar_wait(a_full + 8 * ra.slot, ra.phase, 22);
        if (a.lda) fence_proxy_async();
        tc_fence_after();
        if (g == g_begin) B2M_TRACE(20);
        if (lead) {
          const uint32_t a_lo = (((smem_base + ra.slot * kSlotA) >> 4) & 0x3FFFu) | lbo_a;
          const uint32_t d = tmem_base + (uint32_t)(q * a.colstride);
          uint32_t acc = (used >> q) & 1u;
#pragma unroll
          for (int ks = 0; ks < ROWS / 16; ++ks) {
            umma_bf16_lohi(d, a_lo + ks * kstep_a, hi_a, b_lo + ks * kstep_b, hi_b, idesc, acc);
            acc = 1u;
          }
          umma_commit(a_empty + 8 * ra.slot);
        }
        used |= 1u << q;
        ra.next();
      }
      if (lead) umma_commit(b_empty + 8 * rb.slot);
      rb.next();
    }
    if (lead) {
      *reinterpret_cast<volatile uint32_t*>(used_s) = used;
      __threadfence_block();
      umma_commit(accum_bar);
    }
    B2M_TRACE(21);
    __syncwarp();
  } else if (is_bprod) {
    // ================= dY producers (B operand, MN-major): gather rows order[g*64 ..]; stage s by warp s % kWgBProd ====
    // cp.async mode with several column blocks: a stage is gathered by a GROUP of a.nwb warps, warp `partb` of the
    // group taking the blocks b % nwb == partb, so that the ~8 instructions per 512 bytes a warp spends do not
    // serialise a whole 32 KB stage on one warp (the stage barrier counts 32 arrivals per warp of the group)
    const int pb = bidx / a.nwb, partb = bidx % a.nwb;
    const int npb = min(kWgBProd / a.nwb, SB);   // <= SB, see conv_fwd_kernel
    Ring rb;
    rb.init(SB);
    int turn = 0;
    int Rb_next[ROWS / 32];
    bool rb_have = false;
    int64_t rb_g = -1;
    (void)rb_g;
    auto load_order = [&](int64_t gq, int (&R)[ROWS / 32]) {
      const int64_t p0 = gq * ROWS + lane;
#pragma unroll
      for (int u = 0; u < ROWS / 32; ++u) {
        if (a.order) R[u] = (p0 + 32 * u < a.n_pitch) ? __ldg(a.order + p0 + 32 * u) : -1;
        else R[u] = p0 + 32 * u < a.n_out ? (int)(p0 + 32 * u) : -1;
      }
    };
    const int items = a.nbb * 16;   // (column block, 4-row quad)
    MaskBits m_next = group_mask(g_begin < g_end ? g_begin : 0);
    for (int64_t g = g_begin; g < g_end; g += g_step) {
      const MaskBits m = m_next;
      if (g + g_step < g_end) m_next = group_mask(g + g_step);   // prefetched: the load latency overlaps this group's work
      if (!acc_mask(m)) continue;
      const bool mine = (turn == pb);
      turn = (turn + 1 == npb) ? 0 : turn + 1;
      if (!mine) { rb.next(); continue; }
      const uint32_t b_s = smem_base + a.off_b + rb.slot * a.b_bytes;
      const uint32_t full = b_full + 8 * rb.slot;
      if (a.ldb) {
        // cp.async mode: lane l holds the rows at positions l, 32 + l, ... of the group; the indices of this warp's NEXT
        // group were loaded one stage ago (Rb_next below)
        int R[ROWS / 32];
        if (rb_have) {
#pragma unroll
          for (int u = 0; u < ROWS / 32; ++u) R[u] = Rb_next[u];
        } else {
          load_order(g, R);
        }
        // look ahead: the next group this warp gathers
        {
          int64_t g2 = g + g_step;
          int t2 = turn;                      // turn already advanced past this stage
          rb_have = false;
          while (g2 < g_end) {
            if (acc_mask(group_mask(g2))) {
              if (t2 == pb) { load_order(g2, Rb_next); rb_have = true; rb_g = g2; break; }
              t2 = (t2 + 1 == npb) ? 0 : t2 + 1;
            }
            g2 += g_step;
          }
        }
        mbar_wait(b_empty + 8 * rb.slot, rb.phase ^ 1u, 23);
        const uint8_t* dyb = reinterpret_cast<const uint8_t*>(a.dy);
        const uint32_t row_bytes = (uint32_t)a.c_out * 2u;
        for (int blk = partb; blk < a.nbb; blk += a.nwb) {
          if (a.wb == 128) {
            const int left = (a.c_out - blk * 64) >> 3;
            B2M_WG_SW128<ROWS>(b_s + blk * (ROWS * 128), dyb + blk * 128, row_bytes, R, lane, left < 8 ? left : 8);
          } else {
            B2M_WG_SW64<ROWS>(b_s + blk * (ROWS * 64), dyb + blk * 64, row_bytes, R, lane);
          }
        }
        cp_async_mbar_arrive_noinc(full);
        rb.next();
        continue;
      }
      if (ROWS == 128) {
        // TMA gathers of a 128-row stage: one gather4 per lane and column block, blocks dealt to the a.nwb warps of the group
        const int64_t p128 = g * ROWS + 4 * lane;
        int4 idx;
        if (a.order) {
          idx = ld_nc_int4(a.order + p128);
        } else {
          idx.x = p128 < a.n_out ? (int)p128 : -1;
          idx.y = p128 + 1 < a.n_out ? (int)p128 + 1 : -1;
          idx.z = p128 + 2 < a.n_out ? (int)p128 + 2 : -1;
          idx.w = p128 + 3 < a.n_out ? (int)p128 + 3 : -1;
        }
        int nmine = 0;
        for (int blk = partb; blk < a.nbb; blk += a.nwb) ++nmine;
        mbar_wait(b_empty + 8 * rb.slot, rb.phase ^ 1u, 23);
        if (lane == 0) {
          if (nmine) mbar_arrive_expect_tx(full, (uint32_t)(nmine * ROWS * a.wb)); else mbar_arrive(full);
        }
        __syncwarp();
        for (int blk = partb; blk < a.nbb; blk += a.nwb)
          tma_gather4(b_s + blk * (ROWS * a.wb) + lane * 4 * a.wb, &tm_dy, full, blk * 64, idx.x, idx.y, idx.z, idx.w);
        __syncwarp();
        rb.next();
        continue;
      }
      const int64_t pos0 = g * ROWS + 4 * (lane & 15);
      int4 idx;
      if (a.order) {
        idx = ld_nc_int4(a.order + pos0);
      } else {
        idx.x = pos0 < a.n_out ? (int)pos0 : -1;
        idx.y = pos0 + 1 < a.n_out ? (int)pos0 + 1 : -1;
        idx.z = pos0 + 2 < a.n_out ? (int)pos0 + 2 : -1;
        idx.w = pos0 + 3 < a.n_out ? (int)pos0 + 3 : -1;
      }
      mbar_wait(b_empty + 8 * rb.slot, rb.phase ^ 1u, 23);
      if (g == g_begin) B2M_TRACE(10);
      if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)(a.nbb * ROWS * a.wb));
      __syncwarp();
      for (int it = lane; it < items; it += 32) {
        const int blk = it >> 4;
        tma_gather4(b_s + blk * (ROWS * a.wb) + (it & 15) * 4 * a.wb, &tm_dy, full, blk * 64, idx.x, idx.y, idx.z, idx.w);
      }
      __syncwarp();
      if (g == g_begin) B2M_TRACE(11);
      rb.next();
    }
  } else if (is_aprod) {
    // ================= gather warps (A operand = X rows, MN-major): A stage s is produced by warp s % kWgAProd ====
    const int p = aidx / a.nwa, part = aidx % a.nwa;   // group, warp in it
    const int np = min(kWgAProd / a.nwa, SA);   // <= SA, see conv_fwd_kernel
    Ring ra;
    ra.init(SA);
    int turn = 0;
    const int items = a.nab * 16;   // (block, 4-row quad); pk > 1: block == packed offset slot, else channel block
    if (a.lda && a.cpad != 8) {
      // ---- cp.async mode, software-pipelined: the row indices of this warp's NEXT stage are loaded (streamed from
      // DRAM: ~1-2 us) before the current stage waits for its slot, like the forward kernel's gather warps. Stages are
      // enumerated in the order every role uses: row groups ascending, within a group the set bits of acc_mask
      // ascending; stage n lives in ring slot n % SA and belongs to producer group n % np.
      struct St { bool valid; int64_t g; int q; int slot; uint32_t phase; };
      int64_t ge = g_begin - g_step;
      uint32_t rem = 0;
      int n_mod = 0;
      Ring ring;
      ring.init(SA);
      auto next = [&]() -> St {
        St st; st.valid = false; st.g = 0; st.q = 0; st.slot = 0; st.phase = 0;
        if (p >= np) return st;
        while (true) {
          if (rem == 0) {
            ge += g_step;
            if (ge >= g_end) return st;
            rem = acc_mask(group_mask(ge));
            continue;
          }
          const int q = __ffs(rem) - 1;
          rem &= rem - 1;
          const bool mine = (n_mod == p);
          n_mod = (n_mod + 1 == np) ? 0 : n_mod + 1;
          if (mine) { st.valid = true; st.g = ge; st.q = q; st.slot = ring.slot; st.phase = ring.phase; ring.next(); return st; }
          ring.next();
        }
      };
      auto load_rows = [&](const St& st, int (&R)[4][ROWS / 32]) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int u = 0; u < ROWS / 32; ++u) R[b][u] = -1;
        if (!st.valid) return;
        const int k0 = k_first(st.q);
        const int nslots = min(a.pk, a.kvol - k0);
        const int64_t p0 = st.g * ROWS + lane;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
          for (int u = 0; u < ROWS / 32; ++u) {
            if (b < (a.pk > 1 ? nslots : 1) && p0 + 32 * u < a.n_pitch) {
              if (a.nbr) R[b][u] = __ldg(a.nbr + (int64_t)(k0 + b) * a.n_pitch + p0 + 32 * u);
              else R[b][u] = p0 + 32 * u < a.n_out ? (int)(p0 + 32 * u) : -1;
            }
          }
        }
        if (a.pk == 1) {                 // the column blocks of one offset share its rows
#pragma unroll
          for (int b = 1; b < 4; ++b)
#pragma unroll
            for (int u = 0; u < ROWS / 32; ++u) R[b][u] = R[0][u];
        }
      };
      const uint8_t* xb = reinterpret_cast<const uint8_t*>(a.x);
      const uint32_t row_bytes = (uint32_t)a.c_in * 2u;
      St cur = next();
      int R[4][ROWS / 32];
      load_rows(cur, R);
      while (cur.valid) {
        const St nxt = next();
        int Rn[4][ROWS / 32];
        load_rows(nxt, Rn);                                   // in flight while this stage waits for its slot
        const uint32_t a_s = smem_base + cur.slot * kSlotA;
        const int nslots = min(a.pk, a.kvol - k_first(cur.q));
        mbar_wait(a_empty + 8 * cur.slot, cur.phase ^ 1u, 24);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b < a.nab && !(a.pk > 1 && b >= nslots) && (b % a.nwa) == part) {
            if (a.wa == 128) {
              const int col0 = (a.pk > 1) ? 0 : mt * 128 + b * 64;         // first channel of the block
              const int left = (a.c_in - col0) >> 3;
              B2M_WG_SW128<ROWS>(a_s + b * (ROWS * 128), xb + col0 * 2, row_bytes, R[b], lane,
                                         left < 8 ? (left > 0 ? left : 0) : 8);
            } else {
              B2M_WG_SW64<ROWS>(a_s + b * (ROWS * 64), xb, row_bytes, R[b], lane);
            }
          }
        }
        cp_async_mbar_arrive_noinc(a_full + 8 * cur.slot);
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int u = 0; u < ROWS / 32; ++u) R[b][u] = Rn[b][u];
        cur = nxt;
      }
    } else {
    MaskBits m_next = group_mask(g_begin < g_end ? g_begin : 0);
    for (int64_t g = g_begin; g < g_end; g += g_step) {
      const MaskBits m = m_next;
      if (g + g_step < g_end) m_next = group_mask(g + g_step);   // prefetched: the load latency overlaps this group's work
      // The stages of this row group are the set bits of qm in ascending order; jump to the ones whose turn is this
      // warp's instead of walking all of them.
      const uint32_t qm = (p < np) ? acc_mask(m) : 0u;
      const int n_g = __popc(qm);
      int d0 = p - turn;
      if (d0 < 0) d0 += np;
      const Ring ra0 = ra;
      for (int d = d0; d < n_g; d += np) {
        uint32_t rem = qm;
        for (int i = 0; i < d; ++i) rem &= rem - 1;     // drop the d lowest set bits
        const int q = __ffs(rem) - 1;
        ra = ra0.at(d);
        const bool mine = true;
        if (mine && a.cpad == 8) {
          // cp.async path (8-channel input): lane = (offset slot j of 16, row half rh); one 16-byte piece per (row, offset)
          const uint32_t a_s = smem_base + ra.slot * kSlotA;
          const int j = lane & 15, rh = lane >> 4;
          const int k = k_first(q) + j;
          const int32_t* nb = (a.nbr && k < a.kvol) ? a.nbr + (int64_t)k * a.n_pitch + g * ROWS : nullptr;
          mbar_wait(a_empty + 8 * ra.slot, ra.phase ^ 1u, 24);
#pragma unroll 2
          for (int it = 0; it < 8; ++it) {
            const int r0 = 8 * it + 4 * rh;
            int4 idx = make_int4(-1, -1, -1, -1);
            if (nb) {
              idx = ld_nc_int4(nb + r0);
            } else if (!a.nbr && k < a.kvol) {
              const int64_t q0 = g * ROWS + r0;
              idx.x = q0 < a.n_out ? (int)q0 : -1;
              idx.y = q0 + 1 < a.n_out ? (int)q0 + 1 : -1;
              idx.z = q0 + 2 < a.n_out ? (int)q0 + 2 : -1;
              idx.w = q0 + 3 < a.n_out ? (int)q0 + 3 : -1;
            }
            const int iv[4] = {idx.x, idx.y, idx.z, idx.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int row = r0 + u;
              const uint32_t dst = a_s + (j >> 3) * (ROWS * 128) + row * 128 + (((j & 7) ^ (row & 7)) << 4);
              if (iv[u] >= 0) cp_async16(dst, a.x + (int64_t)iv[u] * 8, 16u);
              else st_shared_zero16(dst);
            }
          }
          cp_async_wait_all();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full + 8 * ra.slot);
        } else if (mine && ROWS == 128) {
          // TMA gathers of a 128-row stage: one gather4 per lane and block (32 quads of rows); the blocks of a stage are
          // dealt to the a.nwa warps of the producer group (block b to warp b % nwa), each of which announces its own
          // bytes on the stage barrier (count nwa), so that issuing a 32 KB stage does not take one warp ~5000 cycles
          const uint32_t a_s = smem_base + ra.slot * kSlotA;
          const uint32_t full = a_full + 8 * ra.slot;
          const int k0 = k_first(q);
          const int nslots = min(a.pk, a.kvol - k0);
          const int64_t pos0 = g * ROWS + 4 * lane;
          int4 idx[4];
          int nmine = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int b = part + j * a.nwa;
            if (b < a.nab && !(a.pk > 1 && b >= nslots)) {
              ++nmine;
              if (a.nbr) {
                idx[j] = ld_nc_int4(a.nbr + (int64_t)((a.pk > 1) ? k0 + b : k0) * a.n_pitch + pos0);
              } else {
                idx[j].x = pos0 < a.n_out ? (int)pos0 : -1;
                idx[j].y = pos0 + 1 < a.n_out ? (int)pos0 + 1 : -1;
                idx[j].z = pos0 + 2 < a.n_out ? (int)pos0 + 2 : -1;
                idx[j].w = pos0 + 3 < a.n_out ? (int)pos0 + 3 : -1;
              }
            }
          }
          mbar_wait(a_empty + 8 * ra.slot, ra.phase ^ 1u, 24);
          if (lane == 0) {
            if (nmine) mbar_arrive_expect_tx(full, (uint32_t)(nmine * ROWS * a.wa)); else mbar_arrive(full);
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int b = part + j * a.nwa;
            if (b < a.nab && !(a.pk > 1 && b >= nslots)) {
              const int colx = (a.pk > 1) ? 0 : mt * 128 + b * 64;
              tma_gather4(a_s + b * (ROWS * a.wa) + lane * 4 * a.wa, &tm_x, full, colx, idx[j].x, idx[j].y, idx[j].z, idx[j].w);
            }
          }
          __syncwarp();
        } else if (mine) {
          const uint32_t a_s = smem_base + ra.slot * kSlotA;
          const uint32_t full = a_full + 8 * ra.slot;
          const int k0 = k_first(q);
          const int nslots = min(a.pk, a.kvol - k0);            // packed offsets that exist
          const int64_t pos0 = g * ROWS + 4 * (lane & 15);
          // index quads: lane handles items it = lane, lane + 32, ...; for pk == 1 every block uses offset k0
          int4 idx[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int it = lane + 32 * u;
            const int blk = it >> 4;
            if (it < items) {
              const int k = (a.pk > 1) ? k0 + blk : k0;
              if (a.pk > 1 && blk >= nslots) continue;
              if (a.nbr) {
                idx[u] = ld_nc_int4(a.nbr + (int64_t)k * a.n_pitch + pos0);
              } else {
                idx[u].x = pos0 < a.n_out ? (int)pos0 : -1;
                idx[u].y = pos0 + 1 < a.n_out ? (int)pos0 + 1 : -1;
                idx[u].z = pos0 + 2 < a.n_out ? (int)pos0 + 2 : -1;
                idx[u].w = pos0 + 3 < a.n_out ? (int)pos0 + 3 : -1;
              }
            }
          }
          mbar_wait(a_empty + 8 * ra.slot, ra.phase ^ 1u, 24);
          if (g == g_begin && p == 0) B2M_TRACE(12);
          const int nblk = (a.pk > 1) ? nslots : a.nab;
          if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)(nblk * ROWS * a.wa));
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int it = lane + 32 * u;
            const int blk = it >> 4;
            if (it < items && !(a.pk > 1 && blk >= nslots)) {
              const int colx = (a.pk > 1) ? 0 : mt * 128 + blk * 64;
              tma_gather4(a_s + blk * (ROWS * a.wa) + (it & 15) * 4 * a.wa, &tm_x, full, colx, idx[u].x, idx[u].y, idx[u].z, idx[u].w);
            }
          }
          __syncwarp();
          if (g == g_begin && p == 0) B2M_TRACE(13);
        }
      }
      ra = ra0.at(n_g);
      turn += n_g;
      while (turn >= np) turn -= np;
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) B2M_TRACE(2);
  if (warp == kWgEpi + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// dw[i] = sum over row splits of part[s][i] in a FIXED order (deterministic). SL lanes per float4: lane j adds the splits
// j, j + SL, ... (ascending), then the lanes' sums are added in ascending j. SL = 8 for small dW with many splits (a
// 32 x 32 x 27 kernel split 140 ways is a chain of 140 dependent-latency loads per thread with SL = 1).
template <int SL>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float4* __restrict__ part, int nsplits, int64_t n4, float4* __restrict__ dw) {
  constexpr int E = 256 / SL;
  __shared__ float4 sh[SL > 1 ? SL : 1][SL > 1 ? E : 1];
  const int e = threadIdx.x % E, j = threadIdx.x / E;
  const int64_t i = (int64_t)blockIdx.x * E + e;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
#pragma unroll 4
    for (int s2 = j; s2 < nsplits; s2 += SL) {
      const float4 v = __ldg(part + (int64_t)s2 * n4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (SL > 1) {
    sh[j][e] = acc;
    __syncthreads();
    if (j == 0 && i < n4) {
#pragma unroll
      for (int jj = 1; jj < SL; ++jj) {
        const float4 v = sh[jj][e];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      dw[i] = acc;
    }
  } else if (i < n4) {
    dw[i] = acc;
  }
}

static int pow2_cols(int need) {
  int c = 32;
  while (c < need) c <<= 1;
  return c;
}

// ---- tensor maps (driver entry point resolved through the runtime, no link-time libcuda dependency) ----
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
static bool resolve_encode() {
  if (g_encode_tiled) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return true;
}
// [rows, cols] bf16 row-major tensor, box = {box_cols, 1} for the row-gather mode, swizzle span = box bytes
static bool make_row_map(CUtensorMap* tm, const void* base, int64_t rows, int cols, int box_cols) {
  if (!resolve_encode()) return false;
  if (rows < 1) rows = 1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 1};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : (box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  return g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace b2m

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
using namespace b2m;

// Per-device cache (indexed by the CUDA device ordinal): SM count and whether the kernels' dynamic shared memory
// opt-in has been made on that device (cudaFuncSetAttribute is per device). Filled on first use by whichever thread
// gets there first; the values are idempotent, so a race only repeats the same calls.
constexpr int kMaxDevices = 64;
struct DeviceCache { int sms; bool fwd_attr, wg_attr; };
static DeviceCache g_dev[kMaxDevices];
static DeviceCache* device_cache() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  DeviceCache* d = &g_dev[dev];
  if (d->sms == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    d->sms = n > 0 ? n : 148;
  }
  return d;
}
// Explicit process-wide tuning options (b2m_set_option), replacing round 1's getenv calls inside the C-ABI.
//   B2M_OPT_MAX_CTAS: upper bound on the CTAs the persistent convolution kernels launch (0 = one per SM). Data-parallel
//   training lowers it while a gradient all-reduce runs on a side stream, so that the NCCL kernels find free SMs
//   instead of forcing a second wave of conv CTAs.
//   B2M_OPT_CHUNKS_PER_STAGE: force 1 or 2 64-wide chunks per pipeline stage of the forward kernel (0 = automatic).
//   B2M_OPT_SPLIT_OFFSETS: 0 = automatic offset splitting on levels with few row tiles, 1 = never split (tests).
//   B2M_OPT_GATHER_MODE: 0 = cp.async row gathers for reductions of >= 48 channels (the MMA issuer fences the proxies),
//   1 = TMA gather4 (round 1), 2 = cp.async with the proxy fence on the producer side (forward kernel only),
//   3 = cp.async in the wgrad kernel too (its default stays TMA gather4).
//   B2M_OPT_ISSUER: 1 = general MMA issue loop of the forward kernel everywhere (0 = lean loop where it applies).
//   B2M_OPT_WGRAD_ROWS: 64 = wgrad pipeline stages of 64 reduction rows always (0 = 128 rows on large levels).
static int g_opt_max_ctas = 0, g_opt_cps = 0, g_opt_nosplit = 0, g_opt_gather = 0, g_opt_wgrows = 0, g_opt_issuer = 0;
static int g_opt_wggroup = 0, g_opt_wgbslots = 0;
static int num_sms() {
  const int sms = device_cache()->sms;
  const int cap = g_opt_max_ctas;
  return (cap > 0 && cap < sms) ? cap : sms;
}
extern "C" int b2m_set_option(int32_t option, int64_t value) {
  switch (option) {
    case B2M_OPT_MAX_CTAS: g_opt_max_ctas = value > 0 ? (int)value : 0; return B2M_OK;
    case B2M_OPT_CHUNKS_PER_STAGE: g_opt_cps = (value == 1 || value == 2) ? (int)value : 0; return B2M_OK;
    case B2M_OPT_SPLIT_OFFSETS: g_opt_nosplit = value ? 1 : 0; return B2M_OK;
    case B2M_OPT_GATHER_MODE: g_opt_gather = (value >= 1 && value <= 3) ? (int)value : 0; return B2M_OK;
    case B2M_OPT_ISSUER: g_opt_issuer = value ? 1 : 0; return B2M_OK;
    case B2M_OPT_WGRAD_GROUP: g_opt_wggroup = (value == 2) ? 2 : 0; return B2M_OK;
    case B2M_OPT_WGRAD_BSLOTS: g_opt_wgbslots = (int)value; return B2M_OK;
    case B2M_OPT_WGRAD_ROWS: g_opt_wgrows = (value == 64) ? 64 : 0; return B2M_OK;
    default: return B2M_ERR_INVALID_ARGUMENT;
  }
}

#ifdef B2M_DEBUG_WAIT
// debug builds only: mapped host buffer that timed-out barrier waits report into ([0] = count, entries from [8])
extern "C" unsigned int* b2m_debug_wait_buffer(void) {
  static unsigned int* host = nullptr;
  if (!host) {
    if (cudaHostAlloc(&host, 8192 * 4, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    for (int i = 0; i < 8192; ++i) host[i] = 0;
    unsigned int* dev = nullptr;
    cudaHostGetDevicePointer(&dev, host, 0);
    cudaMemcpyToSymbol(g_b2m_dbg, &dev, sizeof(dev));
  }
  return host;
}
#endif

extern "C" size_t b2m_packed_weight_bytes(int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode) {
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  const int kpack = conv_kpack(c_red);
  const int nkg = (kvol + kpack - 1) / kpack;
  return (size_t)nkg * c_n * (conv_nfull(c_red) * 128 + conv_rem(c_red) * 64);
}

extern "C" int b2m_pack_weights(const float* kernel, int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode,
                                uint16_t* packed, b2m_stream_t stream) {
  if (!kernel || !packed || kvol <= 0 || c_in <= 0 || c_out <= 0 || mode < 0 || mode > 2) return B2M_ERR_INVALID_ARGUMENT;
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  if (c_n % 8 != 0 || (c_red % 16 != 0 && c_red != 8)) return B2M_ERR_UNSUPPORTED_SHAPE;
  const int64_t total = (int64_t)(b2m_packed_weight_bytes(kvol, c_in, c_out, mode) / 16);
  pack_weights_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(kernel, kvol, c_in, c_out, mode, packed, total);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_pack_weights_batched(const float* const* kernels, uint16_t* const* packed, const int32_t* meta,
                                        const int64_t* group_prefix, int32_t n_jobs, int64_t total_groups,
                                        b2m_stream_t stream) {
  if (n_jobs == 0) return B2M_OK;
  if (!kernels || !packed || !meta || !group_prefix || n_jobs < 0 || total_groups < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (total_groups == 0) return B2M_OK;
  pack_weights_batched_kernel<<<cdiv(total_groups, 256), 256, 0, (cudaStream_t)stream>>>(kernels, packed, meta, group_prefix, n_jobs);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

// Offset slices for a convolution with few row tiles (the deep levels): the serial chain of kvol x chunks pipeline
// stages per tile is what a deep-level launch costs (~1 us per stage with so few CTAs in flight), so the offsets are
// dealt to up to `sms / (row tiles x column tiles)` CTAs per tile; the slices' fp32 partial tiles are summed in a
// fixed order by conv_finalize_kernel (deterministic, unlike atomics). Returns 1 when the launch should not split.
static int conv_offset_slices(int64_t n_out, int32_t c_n, int32_t kvol, int32_t c_red, int* ntiles_n_out) {
  int ntiles_n = 1;
  while (c_n / ntiles_n > 256 || (c_n % ntiles_n) != 0 || ((c_n / ntiles_n) % 16) != 0) {
    ++ntiles_n;
    if (ntiles_n > 8) { *ntiles_n_out = 0; return 1; }
  }
  const int64_t row_tiles = (n_out + kTileM - 1) / kTileM;
  const int sms = num_sms();
  const int kpack = conv_kpack(c_red);
  const int nkg = (kvol + kpack - 1) / kpack;
  int slices = 1;
  if (!g_opt_nosplit && nkg > 1 && row_tiles * ntiles_n * 2 <= sms) {
    // 256-wide outputs: two 128-wide column tiles (a 64 KB weight slot per stage would leave two slots)
    if (c_n / ntiles_n > 128 && (c_n / ntiles_n) % 32 == 0) ntiles_n *= 2;
    int want = (int)(sms / (row_tiles * ntiles_n));
    if (want > nkg) want = nkg;
    if (want < 1) want = 1;
    const int per = (nkg + want - 1) / want;
    slices = (nkg + per - 1) / per;
  } else {
    // Few row tiles but nothing to split over (1x1 convolutions): spread the output columns over more CTAs.
    while (row_tiles * ntiles_n * 2 <= sms && (c_n / ntiles_n) % 32 == 0 && c_n / ntiles_n >= 64) ntiles_n *= 2;
  }
  *ntiles_n_out = ntiles_n;
  return slices;
}

extern "C" size_t b2m_conv_forward_workspace_bytes(int64_t n_out, int32_t c_red, int32_t kvol, int32_t c_n) {
  if (n_out <= 0 || c_n <= 0 || kvol <= 0 || c_red <= 0) return 0;
  int ntiles_n = 0;
  const int slices = conv_offset_slices(n_out, c_n, kvol, c_red, &ntiles_n);
  return slices > 1 ? (size_t)slices * (size_t)n_out * (size_t)c_n * 4 : 0;
}

extern "C" int b2m_conv_forward(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, const int32_t* order,
                                const uint32_t* group_mask, int32_t kvol, int64_t n_out, const uint16_t* packed_w,
                                int32_t c_n, uint16_t* y, double* colsum, b2m_stream_t stream) {
  return b2m_conv_forward_ex(x, n_in, c_red, nbr, order, group_mask, kvol, n_out, packed_w, c_n, y, colsum, nullptr, nullptr,
                             nullptr, 0, nullptr, 0, nullptr, 0, stream);
}

extern "C" int b2m_conv_forward_ex(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, const int32_t* order,
                                   const uint32_t* group_mask, int32_t kvol, int64_t n_out, const uint16_t* packed_w,
                                   int32_t c_n, uint16_t* y, double* colsum, const float* scale, const float* shift,
                                   const uint16_t* residual, int32_t relu, float* y32, int32_t c_store, void* workspace,
                                   size_t workspace_bytes, b2m_stream_t stream) {
  return b2m_conv_dgrad_bn_reduce(x, n_in, c_red, nbr, order, group_mask, kvol, n_out, packed_w, c_n, y, colsum, scale, shift,
                                  residual, relu, y32, c_store, workspace, workspace_bytes, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, stream);
}

extern "C" int b2m_conv_dgrad_bn_reduce(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr,
                                        const int32_t* order, const uint32_t* group_mask, int32_t kvol, int64_t n_out,
                                        const uint16_t* packed_w, int32_t c_n, uint16_t* y, double* colsum, const float* scale,
                                        const float* shift, const uint16_t* residual, int32_t relu, float* y32,
                                        int32_t c_store, void* workspace, size_t workspace_bytes, const uint16_t* bn_x,
                                        const uint8_t* bn_relu_mask, const float* bn_mean, const float* bn_invstd,
                                        double* bn_red, b2m_stream_t stream) {
  if (bn_red && (!bn_x || !bn_mean || !bn_invstd || colsum || scale || shift || relu || y32 || !y))
    return B2M_ERR_INVALID_ARGUMENT;
  if (bn_red && (reinterpret_cast<uintptr_t>(bn_x) & 15) != 0) return B2M_ERR_INVALID_ARGUMENT;
  if (bn_red && conv_kpack(c_red) > 2) return B2M_ERR_UNSUPPORTED_SHAPE;      // no fused instantiation for c_red < 32
  if (!bn_red) bn_x = nullptr;
  double* stat_out = bn_red ? bn_red : colsum;
  if (!x || !packed_w || (!y && !y32) || kvol <= 0 || n_out < 0 || n_in < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!nbr && kvol != 1) return B2M_ERR_INVALID_ARGUMENT;
  if (nbr && !group_mask) return B2M_ERR_INVALID_ARGUMENT;
  if (y32 && (c_store <= 0 || c_store > c_n)) return B2M_ERR_INVALID_ARGUMENT;
  if (c_red <= 0 || (c_red % 16 != 0 && c_red != 8) || c_n <= 0 || c_n % 16 != 0 || c_n > 512 || kvol > 128) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n_out == 0) return B2M_OK;
  if (n_out >= ((int64_t)1 << 31) - 256 || n_in >= ((int64_t)1 << 31) - 256) return B2M_ERR_UNSUPPORTED_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return B2M_ERR_INVALID_ARGUMENT;
  if (residual && (reinterpret_cast<uintptr_t>(residual) & 15) != 0) return B2M_ERR_INVALID_ARGUMENT;
  int ntiles_n = 0;
  int slices = conv_offset_slices(n_out, c_n, kvol, c_red, &ntiles_n);
  if (ntiles_n == 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (slices > 1 && (!workspace || workspace_bytes < (size_t)slices * (size_t)n_out * (size_t)c_n * 4)) {
    // no (or too small a) workspace: run unsplit, like round 1 (columns spread over more CTAs instead)
    slices = 1;
    const int64_t row_tiles = (n_out + kTileM - 1) / kTileM;
    while (row_tiles * ntiles_n * 2 <= num_sms() && (c_n / ntiles_n) % 32 == 0 && c_n / ntiles_n >= 64) ntiles_n *= 2;
  }
  FwdArgs a;
  a.x = x; a.nbr = nbr; a.order = nbr ? order : nullptr; a.gmask = nbr ? group_mask : nullptr;
  a.w = reinterpret_cast<const uint8_t*>(packed_w); a.y = y;
  a.colsum = stat_out; a.n_out = n_out; a.n_pitch = b2m_map_pitch(n_out); a.c_red = c_red; a.kvol = kvol; a.c_n = c_n;
  a.ntile = c_n / ntiles_n;
  a.kpack = conv_kpack(c_red);
  a.nkg = (kvol + a.kpack - 1) / a.kpack;
  a.ksplit = slices;
  a.kg_per = (a.nkg + slices - 1) / slices;
  a.part = slices > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
  // split launches leave the epilogue (and the statistics) to conv_finalize_kernel
  a.ep_scale = slices > 1 ? nullptr : scale; a.ep_shift = slices > 1 ? nullptr : shift;
  a.ep_res = slices > 1 ? nullptr : residual; a.ep_relu = slices > 1 ? 0 : (relu ? 1 : 0);
  a.y32 = slices > 1 ? nullptr : y32; a.c_store = c_store;
  if (slices > 1) a.colsum = nullptr;
  // the fused BatchNorm-backward reduction runs wherever the epilogue runs: here, or in conv_finalize_kernel when split
  a.bnr_x = slices > 1 ? nullptr : bn_x; a.bnr_mask = bn_relu_mask; a.bnr_mean = bn_mean; a.bnr_invstd = bn_invstd;
  a.nfull = conv_nfull(c_red);
  a.rem = conv_rem(c_red);
  a.wa = (a.kpack > 1 && a.kpack < 8) ? c_red * 2 : 128;
  a.kg_bytes = c_n * (a.nfull * 128 + a.rem * 64);
  a.mwords = (kvol + 31) / 32;
  a.colstride = (a.ntile + 31) / 32 * 32;
  a.n_tiles = (int)((n_out + kTileM - 1) / kTileM);
  const int sms = num_sms();
#ifndef B2M_T2_MIN_ROUNDS
#define B2M_T2_MIN_ROUNDS 4
#endif
  a.T = (a.ntile <= 128 && a.n_tiles >= B2M_T2_MIN_ROUNDS * sms) ? 2 : 1;
  a.n_work = (a.n_tiles + a.T - 1) / a.T;
  // KPACK == 1: an A stage holds one or two 64-wide chunks of one offset (each gathered by its own warp) and a weight
  // slot the matching slices. The MMA issuer pays ~1000 cycles (0.5 us) of dependent barrier / fence / commit latency
  // per stage whatever the stage holds (tools/small_conv_trace.py), so a 96- or 128-wide reduction is better off with
  // one stage per offset than two, a 256-wide one with two than four - unless the coarser slots leave too few in
  // flight. The 32-wide remainder chunk takes 8 KB, not a whole 16 KB slot.
  //
  // Ring depths. A slot of either ring is reused only after a full turnaround (MMA completion + commit -> producer ->
  // copy lands -> consumer), measured at ~2.9 us for a gathered A stage (1.2 us of it is the producer warp issuing its
  // 32 gather4s) and ~2.0 us for a weight slot. Throughput = slots in flight / turnaround, so for each candidate
  // stage size shared memory is split to minimise max(2.9 / A slots per ring, 2.0 / B slots), and the stage size with
  // the lower time per offset = stages * max(that, 0.5 us issuer floor) wins (ties: the larger stage).
  const int stage_bytes = kEpiWarps * 32 * kStagePitch * 4;
  const int csum_bytes = 2 * a.ntile * 8;
  const int ep_bytes = 2 * a.ntile * 4;
  {
    const int nch = a.nfull + a.rem;
    const int budget = 227 * 1024 - 1024 - 256 - stage_bytes - csum_bytes - ep_bytes - 16 * 32;
    const int force = g_opt_cps;
    float best_cost = 1e30f;
    a.a_slots = 0;
    for (int cps = 1; cps <= 2; ++cps) {
      if (cps > 1 && (a.kpack != 1 || nch < 2)) break;
      if ((force == 1 || force == 2) && a.kpack == 1 && nch >= 2 && cps != force) continue;
      int a_slot = 0, b_slot = 0;
      for (int ch = 0; ch < cps; ++ch) {            // the first stage of an offset is the largest
        const bool full = (a.kpack > 1) || ch < a.nfull;
        a_slot += full ? kASlotBytes : kASlotBytes / 2;
        b_slot += a.ntile * (full ? 128 : 64);
      }
      int best_sa = 0, best_sb = 0;
      float best = 1e30f;
      for (int sb = 2; sb <= 12; ++sb) {
        int sa = (budget - sb * b_slot) / a_slot;
        if (sa > 12) sa = 12;
        sa -= sa % a.T;
        if (sa < 2 * a.T) break;
        const float ta = 2.9f / (float)(sa / a.T), tb = 2.0f / (float)sb;
        const float t = ta > tb ? ta : tb;
        if (t < best - 1e-6f) { best = t; best_sa = sa; best_sb = sb; }
      }
      if (best_sa == 0) continue;
      const int ncg = (nch + cps - 1) / cps;
      const float cost = (float)ncg * (best > 0.5f ? best : 0.5f);
      if (cost <= best_cost + 1e-6f) {
        best_cost = cost;
        a.cps = cps; a.ncg = ncg; a.a_slot_bytes = a_slot; a.b_bytes = b_slot; a.a_slots = best_sa; a.b_slots = best_sb;
      }
    }
    if (a.a_slots == 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  }
  a.off_b = a.a_slots * a.a_slot_bytes;
  a.off_stage = a.off_b + a.b_slots * a.b_bytes;
  a.off_csum = (a.off_stage + stage_bytes + 15) / 16 * 16;
  a.off_ep = (a.off_csum + csum_bytes + 15) / 16 * 16;
  a.off_bars = (a.off_ep + ep_bytes + 15) / 16 * 16;
  a.tmem_cols = pow2_cols(2 * a.T * a.colstride);
  if (a.tmem_cols > 512) return B2M_ERR_UNSUPPORTED_SHAPE;
  a.ldgsts = (a.kpack == 1 && g_opt_gather != 1) ? (g_opt_gather == 2 ? 2 : 1) : 0;     // modes 0 and 3: cp.async, issuer-side fence
  if (a.kpack == 2 && g_opt_gather != 1) a.ldgsts = 1;                                  // 32-channel rows: cp.async too
  // (measured on k27 256->256 over 1.22 M rows: 2.66 ms lean vs 2.45 ms general - with 256-wide tiles the tensor pipe, not
  // the issue loop, paces the kernel, and the straight-line stage block gives the MMAs less slack; lean for tiles <= 128)
  a.lean_off = (g_opt_issuer || a.ntile > 128) ? 1 : 0;
  a.ablate = 0;
#ifdef B2M_DEBUG_BUILD
  if (const char* e = getenv("B2M_ABLATE")) a.ablate = atoi(e);
#endif
  const int smem_bytes = a.off_bars + 16 * a.a_slots + 16 * a.b_slots + 64 + 1024;
  if (smem_bytes > 227 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tm_main, tm_rem;
  if (!make_row_map(&tm_main, x, n_in, c_red, a.kpack > 1 ? c_red : 64)) return B2M_ERR_CUDA_LAUNCH;
  if (a.rem) { if (!make_row_map(&tm_rem, x, n_in, c_red, 32)) return B2M_ERR_CUDA_LAUNCH; }
  else tm_rem = tm_main;
  DeviceCache* dc = device_cache();
  if (!dc->fwd_attr) {
    if (cudaFuncSetAttribute(conv_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_fwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
    dc->fwd_attr = true;
  }
  int gx = sms / ntiles_n;                                  // keep the total CTA count near one per SM
  if (gx < 1) gx = 1;
  if (gx > a.n_work) gx = a.n_work;
  if (slices > 1) {
    gx = sms / (ntiles_n * slices);
    if (gx < 1) gx = 1;
    if (gx > a.n_work) gx = a.n_work;
  }
  dim3 grid((unsigned)gx, (unsigned)ntiles_n, (unsigned)slices);
  const cudaStream_t st = (cudaStream_t)stream;
  if (a.bnr_x != nullptr && a.kpack == 1) conv_fwd_kernel<1, true><<<grid, kFwdThreads, smem_bytes, st>>>(tm_main, tm_rem, a);
  else if (a.bnr_x != nullptr && a.kpack == 2) conv_fwd_kernel<2, true><<<grid, kFwdThreads, smem_bytes, st>>>(tm_main, tm_rem, a);
  else switch (a.kpack) {
    case 1: conv_fwd_kernel<1><<<grid, kFwdThreads, smem_bytes, st>>>(tm_main, tm_rem, a); break;
    case 2: conv_fwd_kernel<2><<<grid, kFwdThreads, smem_bytes, st>>>(tm_main, tm_rem, a); break;
    case 4: conv_fwd_kernel<4><<<grid, kFwdThreads, smem_bytes, st>>>(tm_main, tm_rem, a); break;
    default: conv_fwd_kernel<8><<<grid, kFwdThreads, smem_bytes, st>>>(tm_main, tm_rem, a); break;
  }
  B2M_CHECK_LAUNCH();
  if (slices > 1) {
    const int rpp = kFinThreads / (c_n / 8);
    int64_t blocks = (n_out + rpp - 1) / rpp;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const size_t sh = stat_out ? (size_t)rpp * c_n * 2 * sizeof(float) : 0;
    conv_finalize_kernel<<<(unsigned)blocks, kFinThreads, sh, st>>>(a.part, slices, n_out, c_n, scale, shift, residual,
                                                                  relu ? 1 : 0, y, y32, c_store, stat_out, bn_x, bn_relu_mask,
                                                                  bn_mean, bn_invstd);
    B2M_CHECK_LAUNCH();
  }
  return B2M_OK;
}

// plan_only: fill *ws_need with the partial-sum workspace the launch would use (0 = none) and return
static int wgrad_run(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                     const int32_t* nbr, const int32_t* order, const uint32_t* group_mask, int32_t kvol,
                     int64_t n_out, float* dw, void* workspace, size_t workspace_bytes, bool plan_only, size_t* ws_need,
                     b2m_stream_t stream) {
  if (ws_need) *ws_need = 0;
  if (plan_only) {
    if (kvol <= 0 || n_out <= 0 || c_in <= 0 || c_out <= 0) return B2M_OK;
  } else if (!x || !dy || !dw || kvol <= 0 || n_out < 0 || n_in < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!plan_only && !nbr && kvol != 1) return B2M_ERR_INVALID_ARGUMENT;
  if (!plan_only && nbr && !group_mask) return B2M_ERR_INVALID_ARGUMENT;
  if (c_in <= 0 || c_in % 8 != 0 || c_out <= 0 || c_out % 16 != 0 || c_out > 256 || kvol > 128) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n_out == 0) {
    if (cudaMemsetAsync(dw, 0, (size_t)kvol * c_in * c_out * 4, (cudaStream_t)stream) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
    return B2M_OK;
  }
  if (n_out >= ((int64_t)1 << 31) - 256 || n_in >= ((int64_t)1 << 31) - 256) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (!plan_only && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) != 0) return B2M_ERR_INVALID_ARGUMENT;
  WgArgs a;
  a.x = x; a.dy = dy; a.nbr = nbr; a.order = nbr ? order : nullptr; a.gmask = nbr ? group_mask : nullptr; a.dw = dw;
  a.n_out = n_out; a.n_pitch = b2m_map_pitch(n_out); a.c_in = c_in; a.c_out = c_out; a.kvol = kvol; a.mwords = (kvol + 31) / 32;
  a.cpad = c_in <= 8 ? 8 : (c_in <= 16 ? 16 : (c_in <= 32 ? 32 : (c_in <= 64 ? 64 : 128)));
  a.pk = 128 / a.cpad;
  a.wa = (a.cpad >= 64 || a.cpad == 8) ? 128 : a.cpad * 2;
  a.nab = 128 * 2 / a.wa;
  a.wb = (c_out == 16 || c_out == 32) ? c_out * 2 : 128;
  a.nbb = (c_out * 2 + a.wb - 1) / a.wb;
  // wgrad keeps the TMA gather4 row gathers by default: measured on k27 96->96 over 1.22 M rows 1.06-1.09 ms in every
  // build, against 1.12-1.46 ms for the cp.async variants (B2M_OPT_GATHER_MODE = 3 selects them)
  // Exception: 32-channel operands. A gather4 then moves only 4 x 64 bytes per instruction and the kernel is bound by the
  // TMA request rate: k27 32->32 over 291 k rows 0.213 ms with TMA gathers, 0.082 ms with cp.async.
  a.lda = ((g_opt_gather == 3 && a.cpad >= 64) || (g_opt_gather != 1 && c_in == 32)) ? 1 : 0;   // SW128 blocks / SW64 for 32 channels
  a.ldb = ((g_opt_gather == 3 && a.wb >= 64) || (g_opt_gather != 1 && a.wb == 64)) ? 1 : 0;
  // (measured on k27 96->96 over 1.22 M rows: two-warp producer groups 1.28 ms, one warp per stage 1.12 ms)
  a.nwa = (g_opt_wggroup == 2 && a.lda && a.nab >= 2) ? 2 : 1;
  a.nwb = (g_opt_wggroup == 2 && a.ldb && a.nbb >= 2) ? 2 : 1;
  const int mtiles = (c_in + 127) / 128;
  a.colstride = (c_out + 31) / 32 * 32;
  int G = 512 / a.colstride;
  if (G > 16) G = 16;
  const int kslots = (kvol + a.pk - 1) / a.pk;          // accumulators needed for all offsets
  if (G > kslots) G = kslots;
  a.G = G;
  const int columns = (kslots + G - 1) / G;
  a.tmem_cols = pow2_cols(G * a.colstride);
  // Reduction rows per pipeline stage. The MMA issuer pays ~1000 cycles of dependent barrier / fence / commit latency per
  // stage whatever it holds, and a 64-row stage is only 4 MMAs (~270 cycles of tensor time): with cp.async gathers on
  // both operands and enough rows, a stage is 128 rows (8 MMAs) of the union of the two 64-row groups' offsets.
  // (measured over 1.22 M rows, k27: 96->96 0.91 ms with 128-row stages / 1.03 ms with 64; 64->64, where a stage packs two
  // offsets of which often only one occurs, 0.66 / 0.57 ms: 128 rows only for one offset per stage or cp.async operands)
  const int rows = (n_out >= 64 * 1024 && g_opt_wgrows != 64 && (a.pk == 1 || (a.lda && a.ldb))) ? 128 : 64;
  if (rows == 128) {            // TMA operands: the blocks of a 128-row stage are dealt to two gather warps
    if (!a.lda && a.nab >= 2) a.nwa = 2;
    if (!a.ldb && a.nbb >= 2) a.nwb = 2;
  }
  const int64_t total_groups = (n_out + rows - 1) / rows;
  const int sms = num_sms();
  // Row splits. Plenty of row groups: one CTA per SM, as many row splits as fit (the kernel is throughput bound). Few row
  // groups (the deep levels, where dW is as large as the activations): every extra row split adds a full set of fp32
  // atomics over dW (measured: 140 CTAs x 256 KB of red.global.add for a 7 MB dW = the whole 50 us of a launch), while
  // a CTA's serial chain is groups x accumulators stages of ~0.6 us. Pick the split count that minimises
  // chain + atomics (1 us per MB of atomics, a third of that for the plain stores of the unsplit case).
  int64_t max_splits = sms / ((int64_t)columns * mtiles);
  if (max_splits < 1) max_splits = 1;
  if (max_splits > total_groups) max_splits = total_groups;
  // With a workspace the row splits write partial sums with plain stores and wgrad_reduce_kernel adds them in a fixed
  // order (deterministic; measured on the full-resolution layers: the 24 row splits' fp32 atomics on the same 1 MB of
  // dW, issued by all CTAs in the same order at the same time, were 27 % of the launch): a split then costs a store and
  // a read of dW at ~3 MB/us each instead of contended atomics.
  const bool can_part = plan_only || workspace != nullptr;
  const double dw_mb = (double)kvol * c_in * c_out * 4.0 / 1e6;
  int64_t splits = max_splits;
  if (total_groups * columns * mtiles < 8 * (int64_t)sms) {
    double best = 1e30;
    for (int64_t sp = 1; sp <= max_splits; ++sp) {
      const int64_t gpc = (total_groups + sp - 1) / sp;
      const double chain = 3.0 + 0.6 * (double)gpc * G;
      const double atom = (sp == 1) ? dw_mb / 3.0 : (can_part ? 3.0 + (double)sp * dw_mb * 2.0 / 3.0 : (double)sp * dw_mb);
      if (chain + atom < best - 1e-9) { best = chain + atom; splits = sp; }
    }
  }
  a.groups_per_cta = (int)((total_groups + splits - 1) / splits);
  splits = (total_groups + a.groups_per_cta - 1) / a.groups_per_cta;
  a.store = (splits == 1) ? 1 : 0;
  const size_t dw_bytes = (size_t)kvol * c_in * c_out * 4;
  const size_t need = (splits > 1) ? (size_t)splits * dw_bytes : 0;
  if (ws_need) *ws_need = need;
  if (plan_only) return B2M_OK;
  a.part = (splits > 1 && workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0)
               ? reinterpret_cast<float*>(workspace) : nullptr;
  a.ncols = columns;
  const int a_slot_bytes = 128 * rows * 2;
  a.b_bytes = a.nbb * rows * a.wb;
  if (a.b_bytes < 1024) a.b_bytes = 1024;
  a.b_bytes = (a.b_bytes + 1023) / 1024 * 1024;
  a.b_slots = a.b_bytes >= 65536 ? 2 : (a.b_bytes >= 32768 ? 3 : 4);
  if (g_opt_wgbslots >= 2 && g_opt_wgbslots <= 8 && g_opt_wgbslots * a.b_bytes <= 128 * 1024) a.b_slots = g_opt_wgbslots;
  a.a_slots = (227 * 1024 - 2048 - 256 - a.b_slots * a.b_bytes) / a_slot_bytes;   // (1 KB of static shared memory)
  if (a.a_slots > 10) a.a_slots = 10;
  if (a.a_slots < 2) return B2M_ERR_UNSUPPORTED_SHAPE;
  a.off_b = a.a_slots * a_slot_bytes;
  a.off_bars = a.off_b + a.b_slots * a.b_bytes;
  const int smem_bytes = a.off_bars + 16 * a.a_slots + 16 * a.b_slots + 64 + 1024;
  if (smem_bytes > 226 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tm_x, tm_dy;
  if (!make_row_map(&tm_x, x, n_in, c_in, a.cpad == 8 ? 8 : a.wa / 2)) return B2M_ERR_CUDA_LAUNCH;   // unused when cpad == 8
  if (!make_row_map(&tm_dy, dy, n_out, c_out, a.wb / 2)) return B2M_ERR_CUDA_LAUNCH;
  DeviceCache* dc = device_cache();
  if (!dc->wg_attr) {
    // (the kernel also has ~0.6 KB of static shared memory: the two together must stay within the 227 KB per block)
    if (cudaFuncSetAttribute(conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
    dc->wg_attr = true;
  }
  if (!a.store && !a.part && cudaMemsetAsync(dw, 0, dw_bytes, (cudaStream_t)stream) != cudaSuccess)
    return B2M_ERR_CUDA_LAUNCH;
  dim3 grid((unsigned)columns, (unsigned)splits, (unsigned)mtiles);
  if (rows == 128) conv_wgrad_kernel<128><<<grid, kWgThreads, smem_bytes, (cudaStream_t)stream>>>(tm_x, tm_dy, a);
  else conv_wgrad_kernel<64><<<grid, kWgThreads, smem_bytes, (cudaStream_t)stream>>>(tm_x, tm_dy, a);
  B2M_CHECK_LAUNCH();
  if (a.part) {
    const int64_t n4 = (int64_t)(dw_bytes / 16);
    if (splits >= 16 && n4 <= (int64_t)256 * 2 * sms)
      wgrad_reduce_kernel<8><<<cdiv(n4, 32), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(a.part),
                                                                          (int)splits, n4, reinterpret_cast<float4*>(dw));
    else
      wgrad_reduce_kernel<1><<<cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(a.part),
                                                                           (int)splits, n4, reinterpret_cast<float4*>(dw));
    B2M_CHECK_LAUNCH();
  }
  return B2M_OK;
}

extern "C" int b2m_conv_wgrad(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                              const int32_t* nbr, const int32_t* order, const uint32_t* group_mask, int32_t kvol,
                              int64_t n_out, float* dw, b2m_stream_t stream) {
  return wgrad_run(x, n_in, c_in, dy, c_out, nbr, order, group_mask, kvol, n_out, dw, nullptr, 0, false, nullptr, stream);
}

extern "C" size_t b2m_conv_wgrad_workspace_bytes(int64_t n_out, int32_t c_in, int32_t c_out, int32_t kvol) {
  size_t need = 0;
  if (wgrad_run(nullptr, 0, c_in, nullptr, c_out, nullptr, nullptr, nullptr, kvol, n_out, nullptr, nullptr, 0, true, &need,
                nullptr) != B2M_OK)
    return 0;
  return need;
}

extern "C" int b2m_conv_wgrad_ex(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                                 const int32_t* nbr, const int32_t* order, const uint32_t* group_mask, int32_t kvol,
                                 int64_t n_out, float* dw, void* workspace, size_t workspace_bytes, b2m_stream_t stream) {
  if ((reinterpret_cast<uintptr_t>(dw) & 15) != 0) workspace = nullptr;      // the reduce kernel writes float4s
  return wgrad_run(x, n_in, c_in, dy, c_out, nbr, order, group_mask, kvol, n_out, dw, workspace, workspace_bytes, false, nullptr,
                   stream);
}
