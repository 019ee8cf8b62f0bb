// Sparse convolution on sm_100a: gather -> tcgen05.mma (UMMA) implicit GEMM with TMEM accumulators.
//
// Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward+backward
// (reference call sites: /root/reference/models/detection_net.py:235-337, models/resnet.py:70-83).
//
// Both kernels consume a SORTED kernel map (b2m_kernel_map_sort): output rows are visited in `order`
// (rows grouped by their neighbour-occupancy bit mask inside blocks of rows), nbr is already permuted to
// that order and group_mask[g] says which offsets occur in each 64-row group, so whole (tile, offset)
// blocks without pairs are never touched.
//
//  conv_fwd_kernel   persistent, warp-specialised, output-stationary implicit GEMM.
//      work item  = T (1 or 2) tiles of 128 output rows x <=256 output columns, accumulators in TMEM,
//                   double-buffered across work items so the epilogue overlaps the next main loop.
//      producers  = 8 warps, one A stage each: gather 128 neighbour rows x 64 reduction elements with
//                   cp.async (16 B) into a 128B-swizzled K-major tile; missing neighbours are zero-stored.
//                   For c_red in {16, 32} a stage packs 4 / 2 kernel offsets along K (no K padding waste).
//      B loader   = one thread: 1-D bulk copies (TMA engine) of pre-swizzled weight slices; a slice is
//                   shared by the T tiles of the work item.
//      MMA issuer = one thread: tcgen05.mma.cta_group::1.kind::f16, M=128, N=ntile, K=16.
//      epilogue   = 4 warps: tcgen05.ld -> per-warp fp32 staging -> BatchNorm column statistics (fp64 atomics)
//                   -> bf16 row stores (scattered through `order`).
//      The same kernel computes dgrad (weights packed transposed / mirrored).
//  conv_wgrad_kernel dW[k] = X_gathered^T * dY.  M = input channels (several offsets packed into the 128 M
//      rows when c_in <= 64), N = c_out, reduction over output rows in 64-row stages, both operands
//      MN-major. One CTA owns up to G accumulators (G*N <= 512 TMEM columns) = G offset groups and a range
//      of row groups; a dY stage is shared by the G gathers. fp32 vector atomics into dW at the end.
#include "common.cuh"

namespace b2m {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
// Reduction index of B: per-offset layout  R = (k, c) with c padded to a multiple of 64 per offset, or the
// "flat" layout (c_red in {16,32}) where 64/c_red offsets share one 64-wide slice. packed[slice][n][64], the
// eight 16-byte groups of a row XOR-swizzled with (n & 7) (UMMA SWIZZLE_128B image of a K-major tile).
__host__ __device__ inline int conv_kpack(int c_red) { return (c_red == 16 || c_red == 32) ? 64 / c_red : 1; }

__global__ void pack_weights_kernel(const float* __restrict__ w, int kvol, int c_in, int c_out, int mode,
                                    uint16_t* __restrict__ packed) {
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  const int kpack = conv_kpack(c_red);
  const int nchunks = (kpack > 1) ? 1 : (c_red + 63) / 64;
  const int nkg = (kvol + kpack - 1) / kpack;
  const int64_t total = (int64_t)nkg * nchunks * c_n * 8;  // one thread per 16-byte group
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int n = (int)(gid % c_n);
  int64_t rest = gid / c_n;
  const int g = (int)(rest % 8);
  rest /= 8;
  const int chunk = (int)(rest % nchunks);
  const int kg = (int)(rest / nchunks);
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int p = g * 8 + e;  // position inside the 64-wide slice
    int k, r;
    if (kpack > 1) { k = kg * kpack + p / c_red; r = p % c_red; } else { k = kg; r = chunk * 64 + p; }
    float f = 0.f;
    if (k < kvol && r < c_red) {
      const int ksrc = (mode == 1) ? (kvol - 1 - k) : k;
      f = (mode == 0) ? w[((int64_t)ksrc * c_in + r) * c_out + n] : w[((int64_t)ksrc * c_in + n) * c_out + r];
    }
    v[e] = __float2bfloat16_rn(f);
  }
  const int64_t row_base = (((int64_t)kg * nchunks + chunk) * c_n + n) * 64;
  const int pg = g ^ (n & 7);
  *reinterpret_cast<uint4*>(packed + row_base + pg * 8) = *reinterpret_cast<const uint4*>(v);
}

struct MaskBits { uint32_t w[4]; };

__device__ __forceinline__ bool mask_any(const MaskBits& m, int k0, int cnt, int kvol) {
  bool any = false;
  for (int j = 0; j < cnt; ++j) {
    const int k = k0 + j;
    if (k < kvol) any = any || ((m.w[k >> 5] >> (k & 31)) & 1u);
  }
  return any;
}
__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 4;
constexpr int kFwdProd = 8;                                   // producer warps == A ring slots
constexpr int kFwdThreads = (kEpiWarps + 2 + kFwdProd) * 32;  // 448
constexpr int kTileM = 128;
constexpr int kStagePitch = 33;

struct FwdArgs {
  const uint16_t* x; const int32_t* nbr; const int32_t* order; const uint32_t* gmask; const uint16_t* w;
  uint16_t* y; double* colsum;
  int64_t n_out;
  int c_red, kvol, c_n, ntile, T, kpack, nkg, nchunks, mwords, colstride, n_tiles, n_work, b_slots, b_bytes;
  int off_b, off_stage, off_bars, tmem_cols;
};

__device__ __forceinline__ void fwd_tile_mask(const FwdArgs& a, int tile, MaskBits& m) {
#pragma unroll
  for (int w = 0; w < 4; ++w) m.w[w] = 0;
  if (tile < a.n_tiles) {
    if (a.gmask == nullptr) { m.w[0] = 1u; return; }  // identity map (kvol == 1)
    const int64_t ngroups = (a.n_out + 63) / 64;
    const int64_t g0 = 2 * (int64_t)tile;
    for (int w = 0; w < a.mwords; ++w) {
      uint32_t v = __ldg(a.gmask + g0 * a.mwords + w);
      if (g0 + 1 < ngroups) v |= __ldg(a.gmask + (g0 + 1) * a.mwords + w);
      m.w[w] = v;
    }
  }
}

__device__ __forceinline__ void fwd_gather_stage(const FwdArgs& a, uint32_t a_s, int tile, int kg, int c, int lane) {
  const int sub = lane & 7, rb = lane >> 3;  // this lane copies 16-byte group `sub` of rows 4*it + rb
  int k, ch, kc;
  if (a.kpack > 1) {
    const int cpr = a.c_red >> 3;  // 16-byte groups per offset
    k = kg * a.kpack + sub / cpr;
    ch = (sub % cpr) * 8;
    kc = 64;
  } else {
    k = kg;
    ch = c * 64 + sub * 8;
    kc = min(64, a.c_red - c * 64);
  }
  if (sub * 8 >= kc) return;
  const bool lane_on = k < a.kvol;
  const int32_t* nb = (a.nbr && lane_on) ? a.nbr + (int64_t)k * a.n_out : nullptr;
  const int64_t row0 = (int64_t)tile * kTileM;
#pragma unroll 1
  for (int it0 = 0; it0 < 32; it0 += 8) {
    int32_t idx[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t o = row0 + 4 * (it0 + u) + rb;
      int32_t v = -1;
      if (lane_on && o < a.n_out) v = nb ? __ldg(nb + o) : (int32_t)o;
      idx[u] = v;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = 4 * (it0 + u) + rb;
      const uint32_t dst = a_s + row * 128 + ((sub ^ (row & 7)) << 4);
      if (idx[u] >= 0) cp_async16(dst, a.x + (int64_t)idx[u] * a.c_red + ch, 16u);
      else st_shared_zero16(dst);
    }
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1) conv_fwd_kernel(const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = kFwdProd, SB = a.b_slots;
  const uint32_t bars = smem_base + a.off_bars;
  const uint32_t a_full = bars, a_empty = bars + 8 * P;
  const uint32_t b_full = bars + 16 * P, b_empty = b_full + 8 * SB;
  const uint32_t acc_full = b_empty + 8 * SB, acc_empty = acc_full + 16;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + a.off_bars + 16 * P + 16 * SB + 32);
  const int n0 = blockIdx.y * a.ntile;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < P; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + 8 * s, 1); mbar_init(acc_empty + 8 * s, kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 5) { tmem_alloc(smem_u32(tmem_ptr_s), (uint32_t)a.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp < kEpiWarps) {
    // ================= epilogue warps: TMEM -> staging -> statistics + stores =================
    float* stage_s = reinterpret_cast<float*>(smem + a.off_stage) + warp * 32 * kStagePitch;
    int wi = 0;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
      const int par = wi & 1;
      mbar_wait(acc_full + 8 * par, (uint32_t)(wi >> 1) & 1u);
      tc_fence_after();
      for (int t = 0; t < a.T; ++t) {
        const int tile = w * a.T + t;
        if (tile >= a.n_tiles) break;
        MaskBits m;
        fwd_tile_mask(a, tile, m);
        const bool has_acc = (m.w[0] | m.w[1] | m.w[2] | m.w[3]) != 0;
        const int64_t pos = (int64_t)tile * kTileM + warp * 32 + lane;  // this lane's row (position in `order`)
        int32_t orow = -1;
        if (pos < a.n_out) orow = a.order ? __ldg(a.order + pos) : (int32_t)pos;
        const int nchunk32 = (a.ntile + 31) / 32;
        for (int cc = 0; cc < nchunk32; ++cc) {
          const int cw = min(32, a.ntile - cc * 32);
          uint32_t v[32];
          if (has_acc) {
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                   (uint32_t)((par * a.T + t) * a.colstride + cc * 32);
            if (cw == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < cw) stage_s[lane * kStagePitch + j] = __uint_as_float(v[j]);
          __syncwarp();
          if (a.colsum != nullptr && lane < cw) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const float f = stage_s[r * kStagePitch + lane];
              s1 += f;
              s2 = fmaf(f, f, s2);
            }
            const int col = n0 + cc * 32 + lane;
            atomicAdd(a.colsum + col, (double)s1);
            atomicAdd(a.colsum + a.c_n + col, (double)s2);
          }
          const int groups = cw / 8;
          for (int e = lane; e < 32 * groups; e += 32) {
            const int row = e / groups, g = e % groups;
            const int32_t r = __shfl_sync(0xffffffffu, orow, row);
            if (r >= 0) {
              const float* sp = stage_s + row * kStagePitch + g * 8;
              __align__(16) __nv_bfloat162 o[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) o[q] = __floats2bfloat162_rn(sp[2 * q], sp[2 * q + 1]);
              *reinterpret_cast<uint4*>(a.y + (int64_t)r * a.c_n + n0 + cc * 32 + g * 8) = *reinterpret_cast<const uint4*>(o);
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * par);
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, a.ntile, 0, 0);
      int ai = 0, bi = 0, wi = 0;
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
        const int par = wi & 1;
        mbar_wait(acc_empty + 8 * par, ((uint32_t)(wi >> 1) & 1u) ^ 1u);
        tc_fence_after();
        MaskBits m[2];
        bool started[2] = {false, false};
        for (int t = 0; t < a.T; ++t) fwd_tile_mask(a, w * a.T + t, m[t]);
        for (int kg = 0; kg < a.nkg; ++kg) {
          bool act[2];
          act[0] = mask_any(m[0], kg * a.kpack, a.kpack, a.kvol);
          act[1] = (a.T > 1) && mask_any(m[1], kg * a.kpack, a.kpack, a.kvol);
          if (!(act[0] || act[1])) continue;
          for (int c = 0; c < a.nchunks; ++c, ++bi) {
            const int bs = bi % SB;
            mbar_wait(b_full + 8 * bs, (uint32_t)(bi / SB) & 1u);
            const uint32_t b_s = smem_base + a.off_b + bs * a.b_bytes;
            const int kc = (a.kpack > 1) ? 64 : min(64, a.c_red - c * 64);
            for (int t = 0; t < a.T; ++t) {
              if (!act[t]) continue;
              const int as = ai % P;
              mbar_wait(a_full + 8 * as, (uint32_t)(ai / P) & 1u);
              tc_fence_after();
              const uint32_t a_s = smem_base + as * (kTileM * 128);
              const uint32_t d = tmem_base + (uint32_t)((par * a.T + t) * a.colstride);
              for (int ks = 0; ks < kc / 16; ++ks) {
                umma_bf16(d, umma_desc_sw128(a_s + ks * 32, 16, 1024), umma_desc_sw128(b_s + ks * 32, 16, 1024), idesc,
                          (started[t] || ks > 0) ? 1u : 0u);
              }
              started[t] = true;
              umma_commit(a_empty + 8 * as);
              ++ai;
            }
            umma_commit(b_empty + 8 * bs);
          }
        }
        umma_commit(acc_full + 8 * par);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ================= B loader: bulk copies of pre-swizzled weight slices =================
    if (lane == 0) {
      int bi = 0;
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        MaskBits m[2];
        for (int t = 0; t < a.T; ++t) fwd_tile_mask(a, w * a.T + t, m[t]);
        for (int kg = 0; kg < a.nkg; ++kg) {
          const bool act = mask_any(m[0], kg * a.kpack, a.kpack, a.kvol) ||
                           ((a.T > 1) && mask_any(m[1], kg * a.kpack, a.kpack, a.kvol));
          if (!act) continue;
          for (int c = 0; c < a.nchunks; ++c, ++bi) {
            const int bs = bi % SB;
            mbar_wait(b_empty + 8 * bs, ((uint32_t)(bi / SB) & 1u) ^ 1u);
            const uint32_t b_s = smem_base + a.off_b + bs * a.b_bytes;
            const uint16_t* src = a.w + (((int64_t)kg * a.nchunks + c) * a.c_n + n0) * 64;
            mbar_arrive_expect_tx(b_full + 8 * bs, (uint32_t)a.b_bytes);
            bulk_g2s(b_s, src, (uint32_t)a.b_bytes, b_full + 8 * bs);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= producer warps: one A stage each =================
    const int p = warp - (kEpiWarps + 2);
    int ai = 0;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      MaskBits m[2];
      for (int t = 0; t < a.T; ++t) fwd_tile_mask(a, w * a.T + t, m[t]);
      for (int kg = 0; kg < a.nkg; ++kg) {
        bool act[2];
        act[0] = mask_any(m[0], kg * a.kpack, a.kpack, a.kvol);
        act[1] = (a.T > 1) && mask_any(m[1], kg * a.kpack, a.kpack, a.kvol);
        if (!(act[0] || act[1])) continue;
        for (int c = 0; c < a.nchunks; ++c) {
          for (int t = 0; t < a.T; ++t) {
            if (!act[t]) continue;
            if (ai % P == p) {
              mbar_wait(a_empty + 8 * p, ((uint32_t)(ai / P) & 1u) ^ 1u);
              fwd_gather_stage(a, smem_base + p * (kTileM * 128), w * a.T + t, kg, c, lane);
              cp_async_wait_all();
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(a_full + 8 * p);
            }
            ++ai;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad kernel
// ------------------------------------------------------------------------------------------------
constexpr int kWgBProd = 4;                                    // warps 0-3: dY producers, then epilogue
constexpr int kWgAProd = 6;                                    // A ring slots == gather warps
constexpr int kWgThreads = (kWgBProd + 1 + kWgAProd) * 32;     // 352
constexpr int kWgRows = 64;

struct WgArgs {
  const uint16_t* x; const uint16_t* dy; const int32_t* nbr; const int32_t* order; const uint32_t* gmask; float* dw;
  int64_t n_out;
  int c_in, c_out, kvol, mwords, cpad, pk, G, colstride, groups_per_cta, b_bytes, off_b, off_bars, tmem_cols;
};

__device__ __forceinline__ void wg_group_mask(const WgArgs& a, int64_t g, MaskBits& m) {
#pragma unroll
  for (int w = 0; w < 4; ++w) m.w[w] = 0;
  if (a.gmask == nullptr) { m.w[0] = 1u; return; }
  for (int w = 0; w < a.mwords; ++w) m.w[w] = __ldg(a.gmask + g * a.mwords + w);
}

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PA = kWgAProd, SB = kWgBProd;
  const uint32_t bars = smem_base + a.off_bars;
  const uint32_t a_full = bars, a_empty = bars + 8 * PA;
  const uint32_t b_full = bars + 16 * PA, b_empty = b_full + 8 * SB;
  const uint32_t accum_bar = b_empty + 8 * SB, flags_bar = accum_bar + 8;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + a.off_bars + 16 * PA + 16 * SB + 16);
  uint8_t* used_s = smem + a.off_bars + 16 * PA + 16 * SB + 32;  // [16]

  const int col = blockIdx.x;             // which G offset groups
  const int mt = blockIdx.z;              // M tile (input-channel block of 128) when c_in > 128
  const int64_t total_groups = (a.n_out + kWgRows - 1) / kWgRows;
  const int64_t g_begin = (int64_t)blockIdx.y * a.groups_per_cta;
  const int64_t g_end = min(total_groups, g_begin + a.groups_per_cta);
  const int kbase = col * a.G * a.pk;     // first kernel offset of accumulator 0

  if (warp == kWgBProd && lane == 0) {
    for (int s = 0; s < PA; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    mbar_init(flags_bar, 1);
    mbar_fence_init();
  }
  if (warp == kWgBProd + 1) { tmem_alloc(smem_u32(tmem_ptr_s), (uint32_t)a.tmem_cols); tmem_relinquish(); }
  if (tid < 16) used_s[tid] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  auto group_any = [&](const MaskBits& m) -> bool {
    return mask_any(m, kbase, min(a.G * a.pk, 128), a.kvol);
  };

  if (warp < kWgBProd) {
    // ================= dY producers (B operand, MN-major): warp b owns B slot b =================
    const int bgroups = a.c_out / 8;
    int bi = 0;
    for (int64_t g = g_begin; g < g_end; ++g) {
      MaskBits m;
      wg_group_mask(a, g, m);
      if (!group_any(m)) continue;
      if (bi % SB == warp) {
        mbar_wait(b_empty + 8 * warp, ((uint32_t)(bi / SB) & 1u) ^ 1u);
        const uint32_t b_s = smem_base + a.off_b + warp * a.b_bytes;
        for (int i = lane; i < kWgRows * bgroups; i += 32) {
          const int row = i / bgroups, jj = i % bgroups;
          const int64_t pos = g * kWgRows + row;
          const uint32_t dst = b_s + (jj >> 3) * (kWgRows * 128) + row * 128 + (((jj & 7) ^ (row & 7)) << 4);
          if (pos < a.n_out) {
            const int64_t o = a.order ? (int64_t)__ldg(a.order + pos) : pos;
            cp_async16(dst, a.dy + o * a.c_out + jj * 8, 16u);
          } else {
            st_shared_zero16(dst);
          }
        }
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_full + 8 * warp);
      }
      ++bi;
    }
    // ================= epilogue: TMEM -> fp32 vector atomics into dw =================
    mbar_wait(accum_bar, 0);
    mbar_wait(flags_bar, 0);
    tc_fence_after();
    const int mrow = warp * 32 + lane;              // TMEM lane == M row
    const int slot = mrow / a.cpad;                 // which packed offset
    const int ci = mrow % a.cpad + mt * 128;
    const int nchunk32 = (a.c_out + 31) / 32;
    for (int q = 0; q < a.G; ++q) {
      if (!used_s[q]) continue;
      const int k = kbase + q * a.pk + slot;
      const bool ok = slot < a.pk && k < a.kvol && ci < a.c_in;
      for (int cc = 0; cc < nchunk32; ++cc) {
        const int cw = min(32, a.c_out - cc * 32);
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * a.colstride + cc * 32);
        if (cw == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (ok) {
          float* dst = a.dw + ((int64_t)k * a.c_in + ci) * a.c_out + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < cw) {
              atomicAdd(reinterpret_cast<float4*>(dst + j),
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                    __uint_as_float(v[j + 3])));
            }
          }
        }
      }
    }
  } else if (warp == kWgBProd) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, a.c_out, 1, 1);
      bool used[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) used[q] = false;
      int ai = 0, bi = 0;
      for (int64_t g = g_begin; g < g_end; ++g) {
        MaskBits m;
        wg_group_mask(a, g, m);
        if (!group_any(m)) continue;
        const int bs = bi % SB;
        mbar_wait(b_full + 8 * bs, (uint32_t)(bi / SB) & 1u);
        const uint32_t b_s = smem_base + a.off_b + bs * a.b_bytes;
#pragma unroll 1
        for (int q = 0; q < a.G; ++q) {
          if (!mask_any(m, kbase + q * a.pk, a.pk, a.kvol)) continue;
          const int as = ai % PA;
          mbar_wait(a_full + 8 * as, (uint32_t)(ai / PA) & 1u);
          tc_fence_after();
          const uint32_t a_s = smem_base + as * (2 * kWgRows * 128);
          const uint32_t d = tmem_base + (uint32_t)(q * a.colstride);
#pragma unroll
          for (int ks = 0; ks < kWgRows / 16; ++ks) {
            // MN-major SW128: LBO = stride between 64-element MN blocks, SBO = stride between 8-row K groups
            umma_bf16(d, umma_desc_sw128(a_s + ks * 2048, kWgRows * 128, 1024),
                      umma_desc_sw128(b_s + ks * 2048, kWgRows * 128, 1024), idesc, (used[q] || ks > 0) ? 1u : 0u);
          }
          used[q] = true;
          umma_commit(a_empty + 8 * as);
          ++ai;
        }
        umma_commit(b_empty + 8 * bs);
        ++bi;
      }
      umma_commit(accum_bar);
#pragma unroll
      for (int q = 0; q < 16; ++q) used_s[q] = used[q] ? 1 : 0;
      __threadfence_block();
      mbar_arrive(flags_bar);
    }
    __syncwarp();
  } else {
    // ================= gather warps (A operand = X rows, MN-major): warp p owns A slot p =================
    const int p = warp - (kWgBProd + 1);
    const int j16 = lane & 15, rsel = lane >> 4;         // 16-byte group j16 of rows 2*it + rsel
    const int m0 = j16 * 8;                               // first M row covered by this lane's group
    const int slot = m0 / a.cpad;
    const int ci = m0 % a.cpad + mt * 128;
    int ai = 0;
    for (int64_t g = g_begin; g < g_end; ++g) {
      MaskBits m;
      wg_group_mask(a, g, m);
      if (!group_any(m)) continue;
      for (int q = 0; q < a.G; ++q) {
        if (!mask_any(m, kbase + q * a.pk, a.pk, a.kvol)) continue;
        if (ai % PA == p) {
          mbar_wait(a_empty + 8 * p, ((uint32_t)(ai / PA) & 1u) ^ 1u);
          const uint32_t a_s = smem_base + p * (2 * kWgRows * 128);
          const int k = kbase + q * a.pk + slot;
          const bool lane_on = slot < a.pk && k < a.kvol && ci < a.c_in;
          const int32_t* nb = (a.nbr && lane_on) ? a.nbr + (int64_t)k * a.n_out : nullptr;
#pragma unroll 1
          for (int it0 = 0; it0 < 32; it0 += 8) {
            int32_t idx[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int64_t pos = g * kWgRows + 2 * (it0 + u) + rsel;
              int32_t v = -1;
              if (lane_on && pos < a.n_out) v = nb ? __ldg(nb + pos) : (int32_t)pos;
              idx[u] = v;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int row = 2 * (it0 + u) + rsel;
              const uint32_t dst = a_s + (j16 >> 3) * (kWgRows * 128) + row * 128 + (((j16 & 7) ^ (row & 7)) << 4);
              if (idx[u] >= 0) cp_async16(dst, a.x + (int64_t)idx[u] * a.c_in + ci, 16u);
              else st_shared_zero16(dst);
            }
          }
          cp_async_wait_all();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full + 8 * p);
        }
        ++ai;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWgBProd + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

static int pow2_cols(int need) {
  int c = 32;
  while (c < need) c <<= 1;
  return c;
}

}  // namespace b2m

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
using namespace b2m;

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

extern "C" size_t b2m_packed_weight_bytes(int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode) {
  const int c_red = (mode == 0) ? c_in : c_out;
  const int c_n = (mode == 0) ? c_out : c_in;
  const int kpack = conv_kpack(c_red);
  const int nchunks = (kpack > 1) ? 1 : (c_red + 63) / 64;
  const int nkg = (kvol + kpack - 1) / kpack;
  return (size_t)nkg * nchunks * c_n * 64 * sizeof(uint16_t);
}

extern "C" int b2m_pack_weights(const float* kernel, int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode,
                                uint16_t* packed, b2m_stream_t stream) {
  if (!kernel || !packed || kvol <= 0 || c_in <= 0 || c_out <= 0 || mode < 0 || mode > 2) return B2M_ERR_INVALID_ARGUMENT;
  const int c_n = (mode == 0) ? c_out : c_in;
  if (c_n % 8 != 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  const int64_t total = (int64_t)(b2m_packed_weight_bytes(kvol, c_in, c_out, mode) / 16);
  pack_weights_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(kernel, kvol, c_in, c_out, mode, packed);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_conv_forward(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, const int32_t* order,
                                const uint32_t* group_mask, int32_t kvol, int64_t n_out, const uint16_t* packed_w,
                                int32_t c_n, uint16_t* y, double* colsum, b2m_stream_t stream) {
  (void)n_in;
  if (!x || !packed_w || !y || kvol <= 0 || n_out < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!nbr && kvol != 1) return B2M_ERR_INVALID_ARGUMENT;
  if (nbr && !group_mask) return B2M_ERR_INVALID_ARGUMENT;
  if (c_red <= 0 || c_red % 16 != 0 || c_n <= 0 || c_n % 16 != 0 || c_n > 512 || kvol > 128) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n_out == 0) return B2M_OK;
  if (n_out >= ((int64_t)1 << 31) - 256) return B2M_ERR_UNSUPPORTED_SHAPE;
  int ntiles_n = 1;
  while (c_n / ntiles_n > 256 || (c_n % ntiles_n) != 0 || ((c_n / ntiles_n) % 16) != 0) {
    ++ntiles_n;
    if (ntiles_n > 8) return B2M_ERR_UNSUPPORTED_SHAPE;
  }
  FwdArgs a;
  a.x = x; a.nbr = nbr; a.order = nbr ? order : nullptr; a.gmask = nbr ? group_mask : nullptr; a.w = packed_w; a.y = y;
  a.colsum = colsum; a.n_out = n_out; a.c_red = c_red; a.kvol = kvol; a.c_n = c_n;
  a.ntile = c_n / ntiles_n;
  a.kpack = conv_kpack(c_red);
  a.nkg = (kvol + a.kpack - 1) / a.kpack;
  a.nchunks = (a.kpack > 1) ? 1 : (c_red + 63) / 64;
  a.mwords = (kvol + 31) / 32;
  a.colstride = (a.ntile + 31) / 32 * 32;
  a.n_tiles = (int)((n_out + kTileM - 1) / kTileM);
  const int sms = num_sms();
  a.T = (a.ntile <= 128 && a.n_tiles >= 4 * sms) ? 2 : 1;
  a.n_work = (a.n_tiles + a.T - 1) / a.T;
  a.b_bytes = a.ntile * 128;
  a.b_slots = (a.b_bytes >= 32768) ? 2 : 3;
  a.off_b = kFwdProd * kTileM * 128;
  a.off_stage = a.off_b + a.b_slots * a.b_bytes;
  a.off_bars = a.off_stage + kEpiWarps * 32 * kStagePitch * 4;
  a.off_bars = (a.off_bars + 15) / 16 * 16;
  a.tmem_cols = pow2_cols(2 * a.T * a.colstride);
  if (a.tmem_cols > 512) return B2M_ERR_UNSUPPORTED_SHAPE;
  const int smem_bytes = a.off_bars + 16 * kFwdProd + 16 * a.b_slots + 64 + 1024;
  if (smem_bytes > 227 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
    attr_set = true;
  }
  int gx = a.n_work < sms ? a.n_work : sms;
  if (ntiles_n > 1) gx = (gx + ntiles_n - 1) / ntiles_n;   // keep the total CTA count near one per SM
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)ntiles_n);
  conv_fwd_kernel<<<grid, kFwdThreads, smem_bytes, (cudaStream_t)stream>>>(a);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_conv_wgrad(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                              const int32_t* nbr, const int32_t* order, const uint32_t* group_mask, int32_t kvol,
                              int64_t n_out, float* dw, b2m_stream_t stream) {
  (void)n_in;
  if (!x || !dy || !dw || kvol <= 0 || n_out < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!nbr && kvol != 1) return B2M_ERR_INVALID_ARGUMENT;
  if (nbr && !group_mask) return B2M_ERR_INVALID_ARGUMENT;
  if (c_in <= 0 || c_in % 8 != 0 || c_out <= 0 || c_out % 16 != 0 || c_out > 256 || kvol > 128) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n_out == 0) return B2M_OK;
  WgArgs a;
  a.x = x; a.dy = dy; a.nbr = nbr; a.order = nbr ? order : nullptr; a.gmask = nbr ? group_mask : nullptr; a.dw = dw;
  a.n_out = n_out; a.c_in = c_in; a.c_out = c_out; a.kvol = kvol; a.mwords = (kvol + 31) / 32;
  a.cpad = c_in <= 16 ? 16 : (c_in <= 32 ? 32 : (c_in <= 64 ? 64 : 128));
  a.pk = 128 / a.cpad;
  const int mtiles = (c_in + 127) / 128;
  a.colstride = (c_out + 31) / 32 * 32;
  int G = 512 / a.colstride;
  if (G > 16) G = 16;
  const int kslots = (kvol + a.pk - 1) / a.pk;          // accumulators needed for all offsets
  if (G > kslots) G = kslots;
  a.G = G;
  const int columns = (kslots + G - 1) / G;
  a.tmem_cols = pow2_cols(G * a.colstride);
  const int64_t total_groups = (n_out + kWgRows - 1) / kWgRows;
  const int sms = num_sms();
  int64_t splits = sms / ((int64_t)columns * mtiles);
  if (splits < 1) splits = 1;
  if (splits > total_groups) splits = total_groups;
  a.groups_per_cta = (int)((total_groups + splits - 1) / splits);
  splits = (total_groups + a.groups_per_cta - 1) / a.groups_per_cta;
  a.b_bytes = ((c_out + 63) / 64) * kWgRows * 128;
  a.off_b = kWgAProd * 2 * kWgRows * 128;
  a.off_bars = a.off_b + kWgBProd * a.b_bytes;
  const int smem_bytes = a.off_bars + 16 * kWgAProd + 16 * kWgBProd + 64 + 1024;
  if (smem_bytes > 227 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
    attr_set = true;
  }
  dim3 grid((unsigned)columns, (unsigned)splits, (unsigned)mtiles);
  conv_wgrad_kernel<<<grid, kWgThreads, smem_bytes, (cudaStream_t)stream>>>(a);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
