// Sparse convolution on sm_100a: gather -> tcgen05.mma (UMMA) implicit GEMM with TMEM accumulators.
//
// Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward+backward
// (reference call sites: /root/reference/models/detection_net.py:235-337, models/resnet.py:70-83).
//
//  conv_fwd_kernel   output-stationary: one CTA owns 128 output rows x (<=256) output columns.
//                    For every kernel offset k that has at least one neighbour in the tile, and every
//                    64-wide slice of the reduction dim, 128 producer threads gather the neighbour rows
//                    with cp.async (16 B, zero-fill for missing neighbours) into a 128B-swizzled K-major
//                    A tile, one thread bulk-copies (TMA engine) the pre-swizzled weight slice B, and one
//                    thread issues tcgen05.mma into a TMEM accumulator. No atomics, one store per output.
//                    The same kernel computes dgrad (weights packed transposed / mirrored).
//  conv_wgrad_kernel dW[k] = X_gathered^T * dY : M = c_in, N = c_out, reduction over output rows;
//                    both operands are MN-major (rows are gathered along K), split over row ranges,
//                    fp32 atomics into dW.
#include "common.cuh"

namespace b2m {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
// packed[k][chunk][n][64]: row n of B (N index), 64 reduction elements of slice `chunk`, the eight 16-byte
// groups of a row XOR-swizzled with (n & 7) (the UMMA SWIZZLE_128B image of a K-major tile).
__global__ void pack_weights_kernel(const float* __restrict__ w, int kvol, int c_in, int c_out, int mode,
                                    uint16_t* __restrict__ packed) {
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  const int nchunks = (c_red + 63) / 64;
  const int64_t total = (int64_t)kvol * nchunks * c_n * 8;  // one thread per 16-byte group
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int n = (int)(gid % c_n);
  int64_t rest = gid / c_n;
  const int g = (int)(rest % 8);
  rest /= 8;
  const int chunk = (int)(rest % nchunks);
  const int k = (int)(rest / nchunks);
  const int ksrc = (mode == 1) ? (kvol - 1 - k) : k;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int r = chunk * 64 + g * 8 + e;
    float f = 0.f;
    if (r < c_red) {
      f = (mode == 0) ? w[((int64_t)ksrc * c_in + r) * c_out + n] : w[((int64_t)ksrc * c_in + n) * c_out + r];
    }
    v[e] = __float2bfloat16_rn(f);
  }
  const int64_t row_base = (((int64_t)k * nchunks + chunk) * c_n + n) * 64;
  const int pg = g ^ (n & 7);
  *reinterpret_cast<uint4*>(packed + row_base + pg * 8) = *reinterpret_cast<const uint4*>(v);
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 192;  // warps 0-3: gather producers + epilogue, warp 4: MMA issuer, warp 5: B loader
constexpr int kTileM = 128;
constexpr int kStagePitch = 33;   // fp32 staging pitch of the epilogue (conflict-free)

struct FwdSmemLayout {
  int stages;
  int a_bytes;       // 128 rows * 128 B
  int b_bytes;       // ntile rows * 128 B
  int off_idx;       // int32 [kvol][128]
  int off_flags;     // uint8 [4][kvol_pad]
  int off_stage;     // float [128][33]
  int off_bars;      // uint64 full[S], empty[S], accum ; uint32 tmem ptr
  int total;
};

static FwdSmemLayout fwd_smem_layout(int kvol, int ntile, int stages) {
  FwdSmemLayout L;
  L.stages = stages;
  L.a_bytes = kTileM * 128;
  L.b_bytes = ntile * 128;
  int off = stages * (L.a_bytes + L.b_bytes);
  L.off_idx = off;      off += kvol * kTileM * 4;
  L.off_flags = off;    off += 4 * ((kvol + 15) / 16 * 16);
  L.off_stage = off;    off += kTileM * kStagePitch * 4;
  off = (off + 15) / 16 * 16;
  L.off_bars = off;     off += (2 * stages + 1) * 8 + 16;
  L.total = off + 1024;  // slack for manual 1024-byte alignment
  return L;
}

__global__ void __launch_bounds__(kFwdThreads)
conv_fwd_kernel(const uint16_t* __restrict__ x, int c_red, const int32_t* __restrict__ nbr, int kvol,
                int64_t n_out, const uint16_t* __restrict__ packed_w, int c_n, int ntile,
                uint16_t* __restrict__ y, double* __restrict__ colsum, FwdSmemLayout L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int S = L.stages;
  const int stage_bytes = L.a_bytes + L.b_bytes;
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem + L.off_idx);
  uint8_t* flags_s = smem + L.off_flags;
  const int kvol_pad = (kvol + 15) / 16 * 16;
  float* stage_s = reinterpret_cast<float*>(smem + L.off_stage);
  const uint32_t bars = smem_base + L.off_bars;
  const uint32_t full_bar0 = bars;
  const uint32_t empty_bar0 = bars + 8 * S;
  const uint32_t accum_bar = bars + 16 * S;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + L.off_bars + (2 * S + 1) * 8);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTileM;
  const int n0 = blockIdx.y * ntile;
  const int nchunks = (c_red + 63) >> 6;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < ntile) tmem_cols <<= 1;

  // ---- phase A: barriers, TMEM, neighbour indices of this tile ----
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar0 + 8 * s, kTileM + 1);  // 128 cp.async arrivals + 1 expect_tx arrival
      mbar_init(empty_bar0 + 8 * s, 1);          // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_ptr_s), tmem_cols);
    tmem_relinquish();
  }
  if (warp < 4) {
    const int64_t r = row0 + tid;
    for (int k = 0; k < kvol; ++k) {
      int32_t v = -1;
      if (r < n_out) v = nbr ? __ldg(nbr + (int64_t)k * n_out + r) : (int32_t)r;
      idx_s[k * kTileM + tid] = v;
      const unsigned any = __ballot_sync(0xffffffffu, v >= 0);
      if (lane == 0) flags_s[warp * kvol_pad + k] = any ? 1 : 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  auto offset_active = [&](int k) -> bool {
    return (flags_s[k] | flags_s[kvol_pad + k] | flags_s[2 * kvol_pad + k] | flags_s[3 * kvol_pad + k]) != 0;
  };

  if (warp < 4) {
    // ================= gather producers =================
    const int sub = tid & 7;     // 16-byte group inside the 128-byte row slice
    const int rbase = tid >> 3;  // 0..15
    int it = 0;
    for (int k = 0; k < kvol; ++k) {
      if (!offset_active(k)) continue;
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(empty_bar0 + 8 * s, ph ^ 1u);
        const int kc = min(64, c_red - c * 64);
        const uint32_t a_s = smem_base + s * stage_bytes;
        if (sub * 8 < kc) {
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const int row = p * 16 + rbase;
            const int32_t idx = idx_s[k * kTileM + row];
            const uint16_t* src = x + (idx >= 0 ? ((int64_t)idx * c_red + c * 64 + sub * 8) : 0);
            cp_async16(a_s + row * 128 + ((sub ^ (row & 7)) << 4), src, idx >= 0 ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive_noinc(full_bar0 + 8 * s);
      }
    }
    // ================= epilogue =================
    const bool any_work = it > 0;
    if (any_work) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const int nchunk32 = (ntile + 31) / 32;
    for (int cc = 0; cc < nchunk32; ++cc) {
      const int cw = min(32, ntile - cc * 32);  // 32 or 16
      uint32_t v[32];
      if (any_work) {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32);
        if (cw == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cw) stage_s[tid * kStagePitch + j] = __uint_as_float(v[j]);
      named_bar_sync(1, 128);
      // per-column partial sums over this warp's 32 rows -> global fp64 accumulators
      if (colsum != nullptr && lane < cw) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const float f = stage_s[(warp * 32 + r) * kStagePitch + lane];
          s1 += f;
          s2 = fmaf(f, f, s2);
        }
        const int col = n0 + cc * 32 + lane;
        atomicAdd(colsum + col, (double)s1);
        atomicAdd(colsum + c_n + col, (double)s2);
      }
      // bf16 store: 8 columns (16 B) per thread per pass
      const int groups = cw / 8;  // 4 or 2
      for (int e = tid; e < kTileM * groups; e += 128) {
        const int row = e / groups;
        const int g = e % groups;
        const int64_t r = row0 + row;
        if (r < n_out) {
          const float* sp = stage_s + row * kStagePitch + g * 8;
          __align__(16) __nv_bfloat162 o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = __floats2bfloat162_rn(sp[2 * q], sp[2 * q + 1]);
          *reinterpret_cast<uint4*>(y + r * c_n + n0 + cc * 32 + g * 8) = *reinterpret_cast<const uint4*>(o);
        }
      }
      named_bar_sync(1, 128);
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, ntile, 0, 0);
      int it = 0;
      for (int k = 0; k < kvol; ++k) {
        if (!offset_active(k)) continue;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(full_bar0 + 8 * s, ph);
          tc_fence_after();
          const int kc = min(64, c_red - c * 64);
          const uint32_t a_s = smem_base + s * stage_bytes;
          const uint32_t b_s = a_s + L.a_bytes;
          for (int ks = 0; ks < kc / 16; ++ks) {
            const uint64_t adesc = umma_desc_sw128(a_s + ks * 32, 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(b_s + ks * 32, 16, 1024);
            umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar0 + 8 * s);
        }
      }
      if (it > 0) umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    // ================= B loader (TMA engine bulk copies of pre-swizzled weight slices) =================
    if (lane == 0) {
      int it = 0;
      const uint32_t bytes = (uint32_t)ntile * 128u;
      for (int k = 0; k < kvol; ++k) {
        if (!offset_active(k)) continue;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(empty_bar0 + 8 * s, ph ^ 1u);
          const uint32_t b_s = smem_base + s * stage_bytes + L.a_bytes;
          const uint16_t* src = packed_w + (((int64_t)k * nchunks + c) * c_n + n0) * 64;
          mbar_arrive_expect_tx(full_bar0 + 8 * s, bytes);
          bulk_g2s(b_s, src, bytes, full_bar0 + 8 * s);
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad kernel
// ------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 160;     // warps 0-3 producers + epilogue, warp 4 MMA issuer
constexpr int kWgRows = 64;         // output rows (reduction elements) per stage
constexpr int kWgMaxGroups = 4096;  // 64-row groups per CTA

struct WgSmemLayout {
  int stages;
  int a_bytes;    // 2 channel blocks * 64 rows * 128 B
  int b_bytes;    // nblk * 64 rows * 128 B
  int off_flags;  // uint8 [kWgMaxGroups]
  int off_bars;
  int total;
};

static WgSmemLayout wg_smem_layout(int ntile, int stages) {
  WgSmemLayout L;
  L.stages = stages;
  L.a_bytes = 2 * kWgRows * 128;
  L.b_bytes = ((ntile + 63) / 64) * kWgRows * 128;
  int off = stages * (L.a_bytes + L.b_bytes);
  L.off_flags = off;  off += kWgMaxGroups;
  L.off_bars = off;   off += (2 * stages + 1) * 8 + 16;
  L.total = off + 1024;
  return L;
}

__global__ void __launch_bounds__(kWgThreads)
conv_wgrad_kernel(const uint16_t* __restrict__ x, int c_in, const uint16_t* __restrict__ dy, int c_out,
                  const int32_t* __restrict__ nbr, int kvol, int64_t n_out, int groups_per_cta,
                  float* __restrict__ dw, WgSmemLayout L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int S = L.stages;
  const int stage_bytes = L.a_bytes + L.b_bytes;
  uint8_t* gflag_s = smem + L.off_flags;
  const uint32_t bars = smem_base + L.off_bars;
  const uint32_t full_bar0 = bars;
  const uint32_t empty_bar0 = bars + 8 * S;
  const uint32_t accum_bar = bars + 16 * S;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + L.off_bars + (2 * S + 1) * 8);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int k = blockIdx.x;
  const int m0 = blockIdx.z * 128;  // first input channel of this M tile
  const int ntile = c_out;
  const int64_t total_groups = (n_out + kWgRows - 1) / kWgRows;
  const int64_t g_begin = (int64_t)blockIdx.y * groups_per_cta;
  const int64_t g_end = min(total_groups, g_begin + groups_per_cta);
  const int ngroups = (int)max((int64_t)0, g_end - g_begin);
  const int32_t* nbr_k = nbr ? nbr + (int64_t)k * n_out : nullptr;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < ntile) tmem_cols <<= 1;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar0 + 8 * s, 128);
      mbar_init(empty_bar0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_ptr_s), tmem_cols);
    tmem_relinquish();
  }
  // which 64-row groups of this CTA's range contain at least one pair for offset k
  if (warp < 4) {
    for (int g = warp; g < ngroups; g += 4) {
      const int64_t r = (g_begin + g) * kWgRows + lane;
      bool v0 = false, v1 = false;
      if (nbr_k) {
        if (r < n_out) v0 = __ldg(nbr_k + r) >= 0;
        if (r + 32 < n_out) v1 = __ldg(nbr_k + r + 32) >= 0;
      } else {
        v0 = r < n_out;
        v1 = r + 32 < n_out;
      }
      const unsigned any = __ballot_sync(0xffffffffu, v0 || v1);
      if (lane == 0) gflag_s[g] = any ? 1 : 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp < 4) {
    // ================= producers: gather X rows (A) and copy dY rows (B), MN-major swizzled =================
    const int bgroups = ntile / 8;  // 16-byte groups per dY row
    int it = 0;
    for (int g = 0; g < ngroups; ++g) {
      if (!gflag_s[g]) continue;
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(empty_bar0 + 8 * s, ph ^ 1u);
      const uint32_t a_s = smem_base + s * stage_bytes;
      const uint32_t b_s = a_s + L.a_bytes;
      const int64_t rbase = (g_begin + g) * kWgRows;
      // A: 64 rows x 16 groups (2 blocks of 64 channels)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int q = i * 128 + tid;
        const int row = q >> 4;
        const int j16 = q & 15;
        const int64_t r = rbase + row;
        int32_t idx = -1;
        if (r < n_out) idx = nbr_k ? __ldg(nbr_k + r) : (int32_t)r;
        const int ch = m0 + j16 * 8;
        const bool ok = idx >= 0 && ch < c_in;
        const uint16_t* src = x + (ok ? ((int64_t)idx * c_in + ch) : 0);
        cp_async16(a_s + (j16 >> 3) * (kWgRows * 128) + row * 128 + (((j16 & 7) ^ (row & 7)) << 4), src, ok ? 16u : 0u);
      }
      // B: 64 rows x bgroups groups
      for (int q = tid; q < kWgRows * bgroups; q += 128) {
        const int row = q / bgroups;
        const int jj = q % bgroups;
        const int64_t r = rbase + row;
        const bool ok = r < n_out;
        const uint16_t* src = dy + (ok ? (r * c_out + jj * 8) : 0);
        cp_async16(b_s + (jj >> 3) * (kWgRows * 128) + row * 128 + (((jj & 7) ^ (row & 7)) << 4), src, ok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(full_bar0 + 8 * s);
      ++it;
    }
    // ================= epilogue: TMEM -> fp32 atomics into dw[k] =================
    if (it > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const int ci = m0 + tid;  // TMEM lane == M row == input channel
      const int nchunk32 = (ntile + 31) / 32;
      for (int cc = 0; cc < nchunk32; ++cc) {
        const int cw = min(32, ntile - cc * 32);
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32);
        if (cw == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (ci < c_in) {
          float* dst = dw + ((int64_t)k * c_in + ci) * c_out + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < cw) {
              float4 f = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              atomicAdd(reinterpret_cast<float4*>(dst + j), f);
            }
          }
        }
      }
    }
  } else {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, ntile, 1, 1);
      int it = 0;
      for (int g = 0; g < ngroups; ++g) {
        if (!gflag_s[g]) continue;
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(full_bar0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t a_s = smem_base + s * stage_bytes;
        const uint32_t b_s = a_s + L.a_bytes;
#pragma unroll
        for (int ks = 0; ks < kWgRows / 16; ++ks) {
          // MN-major SW128: LBO = stride between 64-element MN blocks, SBO = stride between 8-row K groups
          const uint64_t adesc = umma_desc_sw128(a_s + ks * 2048, kWgRows * 128, 1024);
          const uint64_t bdesc = umma_desc_sw128(b_s + ks * 2048, kWgRows * 128, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar0 + 8 * s);
        ++it;
      }
      if (it > 0) umma_commit(accum_bar);
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace b2m

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
using namespace b2m;

extern "C" size_t b2m_packed_weight_bytes(int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode) {
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  const int nchunks = (c_red + 63) / 64;
  return (size_t)kvol * nchunks * c_n * 64 * sizeof(uint16_t);
}

extern "C" int b2m_pack_weights(const float* kernel, int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode,
                                uint16_t* packed, b2m_stream_t stream) {
  if (!kernel || !packed || kvol <= 0 || c_in <= 0 || c_out <= 0 || mode < 0 || mode > 2) return B2M_ERR_INVALID_ARGUMENT;
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  if (c_n % 8 != 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  const int nchunks = (c_red + 63) / 64;
  const int64_t total = (int64_t)kvol * nchunks * c_n * 8;
  pack_weights_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(kernel, kvol, c_in, c_out, mode, packed);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_conv_forward(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, int32_t kvol,
                                int64_t n_out, const uint16_t* packed_w, int32_t c_n, uint16_t* y, double* colsum,
                                b2m_stream_t stream) {
  (void)n_in;
  if (!x || !packed_w || !y || kvol <= 0 || n_out < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!nbr && kvol != 1) return B2M_ERR_INVALID_ARGUMENT;
  if (c_red <= 0 || c_red % 16 != 0 || c_n <= 0 || c_n % 16 != 0 || c_n > 512 || kvol > 125) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n_out == 0) return B2M_OK;
  int ntiles_n = 1;
  while (c_n / ntiles_n > 256 || (c_n % ntiles_n) != 0 || ((c_n / ntiles_n) % 16) != 0) {
    ++ntiles_n;
    if (ntiles_n > 8) return B2M_ERR_UNSUPPORTED_SHAPE;
  }
  const int ntile = c_n / ntiles_n;
  // stages: as many as fit in ~100 KB (two CTAs per SM), at least 2, at most 6
  int stages = 6;
  FwdSmemLayout L = fwd_smem_layout(kvol, ntile, stages);
  while (stages > 2 && L.total > 110 * 1024) { --stages; L = fwd_smem_layout(kvol, ntile, stages); }
  if (L.total > 227 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
    attr_set = true;
  }
  dim3 grid((unsigned)cdiv(n_out, kTileM), (unsigned)ntiles_n);
  conv_fwd_kernel<<<grid, kFwdThreads, L.total, (cudaStream_t)stream>>>(x, c_red, nbr, kvol, n_out, packed_w, c_n,
                                                                        ntile, y, colsum, L);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_conv_wgrad(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                              const int32_t* nbr, int32_t kvol, int64_t n_out, float* dw, b2m_stream_t stream) {
  (void)n_in;
  if (!x || !dy || !dw || kvol <= 0 || n_out < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!nbr && kvol != 1) return B2M_ERR_INVALID_ARGUMENT;
  if (c_in <= 0 || c_in % 8 != 0 || c_out <= 0 || c_out % 16 != 0 || c_out > 256) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n_out == 0) return B2M_OK;
  const int64_t total_groups = (n_out + kWgRows - 1) / kWgRows;
  const int mtiles = (c_in + 127) / 128;
  // aim for ~4 waves of 148 CTAs, bounded by the per-CTA group-flag capacity
  int64_t splits = (4 * 148 + (int64_t)kvol * mtiles - 1) / ((int64_t)kvol * mtiles);
  if (splits < 1) splits = 1;
  if (splits > total_groups) splits = total_groups;
  int64_t gpc = (total_groups + splits - 1) / splits;
  if (gpc > kWgMaxGroups) gpc = kWgMaxGroups;
  splits = (total_groups + gpc - 1) / gpc;
  if (splits > 65535) return B2M_ERR_UNSUPPORTED_SHAPE;
  int stages = 4;
  WgSmemLayout L = wg_smem_layout(c_out, stages);
  while (stages > 2 && L.total > 110 * 1024) { --stages; L = wg_smem_layout(c_out, stages); }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
    attr_set = true;
  }
  dim3 grid((unsigned)kvol, (unsigned)splits, (unsigned)mtiles);
  conv_wgrad_kernel<<<grid, kWgThreads, L.total, (cudaStream_t)stream>>>(x, c_in, dy, c_out, nbr, kvol, n_out,
                                                                        (int)gpc, dw, L);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
