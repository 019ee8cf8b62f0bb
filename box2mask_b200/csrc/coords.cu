// Coordinate hash, strided coordinate maps and kernel-map builders.
//
// Replaces the coordinate manager of MinkowskiEngine 0.5.4 as reached from
// /root/reference/models/model.py:43 (ME.SparseTensor), models/detection_net.py:42-133 (stride-2 and
// transposed convolutions) and models/resnet.py:61-65 (3x3x3 convolutions).
// All of this is integer work bounded by HBM/L2 latency: one thread per (row, offset) probe,
// coalesced int4 coordinate reads, coalesced int32 table writes.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace b2m {

__device__ __forceinline__ int floor_to_multiple(int v, int s) {
  int q = (v >= 0) ? (v / s) : -((-v + s - 1) / s);
  return q * s;
}

__global__ void hash_insert_kernel(const int4* __restrict__ coords, int64_t n, unsigned long long* keys,
                                   int32_t* vals, uint64_t mask, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(coords + i);
  if (!coord_in_range(c.x, c.y, c.z, c.w)) { atomicAdd(status + 1, 1); return; }
  const unsigned long long key = pack_key(c.x, c.y, c.z, c.w);
  uint64_t slot = hash64(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(keys + slot, (unsigned long long)B2M_KEY_EMPTY, key);
    if (prev == B2M_KEY_EMPTY || prev == key) {
      atomicMin(vals + slot, (int32_t)i);
      if (prev == key) atomicAdd(status, 1);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ int32_t hash_lookup(const unsigned long long* __restrict__ keys,
                                               const int32_t* __restrict__ vals, uint64_t mask,
                                               unsigned long long key) {
  uint64_t slot = hash64(key) & mask;
  while (true) {
    const unsigned long long kk = __ldg(keys + slot);
    if (kk == key) return __ldg(vals + slot);
    if (kk == B2M_KEY_EMPTY) return -1;
    slot = (slot + 1) & mask;
  }
}

__global__ void hash_query_kernel(const int4* __restrict__ q, int64_t m, const unsigned long long* __restrict__ keys,
                                  const int32_t* __restrict__ vals, uint64_t mask, int32_t* __restrict__ rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = __ldg(q + i);
  int32_t r = -1;
  if (coord_in_range(c.x, c.y, c.z, c.w)) r = hash_lookup(keys, vals, mask, pack_key(c.x, c.y, c.z, c.w));
  rows[i] = r;
}

// nbr[k][row] for an odd kernel, offsets (ix - K/2, iy - K/2, iz - K/2) * tensor_stride, x fastest.
__global__ void kmap_submanifold_kernel(const int4* __restrict__ coords, int64_t n, int ts, int ksize,
                                        const unsigned long long* __restrict__ keys,
                                        const int32_t* __restrict__ vals, uint64_t mask,
                                        int32_t* __restrict__ nbr, int64_t pitch) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kvol = ksize * ksize * ksize;
  if (gid >= pitch * kvol) return;
  const int64_t row = gid % pitch;
  const int k = (int)(gid / pitch);
  if (row >= n) { nbr[gid] = -1; return; }  // padding up to the pitch
  const int h = ksize / 2;
  const int dx = (k % ksize - h) * ts;
  const int dy = ((k / ksize) % ksize - h) * ts;
  const int dz = (k / (ksize * ksize) - h) * ts;
  int32_t r;
  if (dx == 0 && dy == 0 && dz == 0) {
    r = (int32_t)row;  // the centre offset always maps a row to itself (coordinates are unique)
  } else {
    const int4 c = __ldg(coords + row);
    const int x = c.y + dx, y = c.z + dy, z = c.w + dz;
    r = -1;
    if (coord_in_range(c.x, x, y, z)) r = hash_lookup(keys, vals, mask, pack_key(c.x, x, y, z));
  }
  nbr[gid] = r;
}

__global__ void downsample_keys_kernel(const int4* __restrict__ coords, int64_t n, int stride,
                                       unsigned long long* __restrict__ keys, int32_t* __restrict__ idx,
                                       int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(coords + i);
  const int x = floor_to_multiple(c.y, stride), y = floor_to_multiple(c.z, stride), z = floor_to_multiple(c.w, stride);
  if (!coord_in_range(c.x, x, y, z)) atomicAdd(status, 1);
  keys[i] = pack_key(c.x, x, y, z);
  idx[i] = (int32_t)i;
}

__global__ void head_flags_kernel(const unsigned long long* __restrict__ sorted_keys, int64_t n, int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1 : 0;
}

__global__ void downsample_emit_kernel(const unsigned long long* __restrict__ sorted_keys,
                                       const int32_t* __restrict__ sorted_idx, const int32_t* __restrict__ rank_incl,
                                       int64_t n, int4* __restrict__ out_coords, int32_t* __restrict__ parent_row,
                                       int32_t* __restrict__ n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t row = rank_incl[i] - 1;
  parent_row[sorted_idx[i]] = row;
  if (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) {
    int b, x, y, z;
    unpack_key(sorted_keys[i], b, x, y, z);
    out_coords[row] = make_int4(b, x, y, z);
  }
  if (i == n - 1) *n_out = row + 1;
}

__global__ void kmap_stride2_kernel(const int4* __restrict__ fine, int64_t n_fine, const int32_t* __restrict__ parent,
                                    int64_t coarse_pitch, int fs, int32_t* __restrict__ nbr_down,
                                    int32_t* __restrict__ nbr_up, int64_t fine_pitch) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= fine_pitch) return;
  if (f >= n_fine) {
    if (nbr_up) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) nbr_up[(int64_t)kk * fine_pitch + f] = -1;
    }
    return;
  }
  const int4 c = __ldg(fine + f);
  const int cs = 2 * fs;
  const int ox = (c.y - floor_to_multiple(c.y, cs)) / fs;
  const int oy = (c.z - floor_to_multiple(c.z, cs)) / fs;
  const int oz = (c.w - floor_to_multiple(c.w, cs)) / fs;
  const int k = ox + 2 * oy + 4 * oz;
  const int32_t p = parent[f];
  if (nbr_down) nbr_down[(int64_t)k * coarse_pitch + p] = (int32_t)f;
  if (nbr_up) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) nbr_up[(int64_t)kk * fine_pitch + f] = (kk == k) ? p : -1;
  }
}

// Kernel map of a level from the maps of the next COARSER level (no hashing): the coordinate o + d * ts (|d| <= 1 for a
// 3^3 kernel, <= 2 for 5^3) lies in the parent cell of o shifted by dP in {-1, 0, 1}^3, so
//   nbr[k][o] = nbr_down[child slot of the target][ nbr3_coarse[kP][ parent[o] ] ]
// two dependent reads of small dense tables (the coarse 3^3 map and the 8-slot child table of the stride-2 map) that
// neighbouring rows share, instead of a random probe of a hash table per (row, offset). With s = child slot of o inside
// its parent (0/1 per axis): dP = floor((s + d) / 2), target slot = (s + d) - 2 dP. Exact: a coordinate exists iff its
// parent cell exists and has that child. group_mask (optional, zeroed by the caller) receives the per-64-row-group
// offset bits (for maps that are not re-ordered afterwards, i.e. the 125-offset map).
__global__ void kmap_from_coarse_kernel(const int4* __restrict__ coords, int64_t n, int ts, int ksize,
                                        const int32_t* __restrict__ parent, const int32_t* __restrict__ nbr3_coarse,
                                        const int32_t* __restrict__ nbr_down, int64_t coarse_pitch,
                                        int32_t* __restrict__ nbr, int64_t pitch, uint32_t* __restrict__ group_mask, int words) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kvol = ksize * ksize * ksize;
  if (gid >= pitch * kvol) return;             // pitch is a multiple of 128: whole warps leave together
  const int64_t row = gid % pitch;
  const int k = (int)(gid / pitch);
  int32_t r = -1;
  if (row < n) {
    const int h = ksize / 2;
    const int dx = k % ksize - h, dy = (k / ksize) % ksize - h, dz = k / (ksize * ksize) - h;
    if (dx == 0 && dy == 0 && dz == 0) {
      r = (int32_t)row;
    } else {
      const int4 c = __ldg(coords + row);
      const int cs = 2 * ts;
      const int px = (c.y - floor_to_multiple(c.y, cs)) / ts + dx;     // position inside the parent cell, in [-2, 3]
      const int py = (c.z - floor_to_multiple(c.z, cs)) / ts + dy;
      const int pz = (c.w - floor_to_multiple(c.w, cs)) / ts + dz;
      const int qx = (px + 2) / 2 - 1, qy = (py + 2) / 2 - 1, qz = (pz + 2) / 2 - 1;   // floor(p / 2) in {-1, 0, 1}
      const int kp = (qx + 1) + 3 * (qy + 1) + 9 * (qz + 1);
      const int slot = (px - 2 * qx) + 2 * (py - 2 * qy) + 4 * (pz - 2 * qz);
      const int32_t pc = __ldg(nbr3_coarse + (int64_t)kp * coarse_pitch + __ldg(parent + row));
      if (pc >= 0) r = __ldg(nbr_down + (int64_t)slot * coarse_pitch + pc);
    }
  }
  nbr[gid] = r;
  if (group_mask != nullptr) {
    const unsigned any = __ballot_sync(0xffffffffu, r >= 0);
    if ((threadIdx.x & 31) == 0 && any && row < n) atomicOr(group_mask + (row >> 6) * words + (k >> 5), 1u << (k & 31));
  }
}

// ---- the same table, one thread per ROW (used when the caller provides a workspace for the transposed child table) ----
// The kernel above spends two dependent 4-byte reads of L2-resident tables per ENTRY (250 per row for a 5^3 map): 306 M
// sector requests for the 125 x 1.22 M table of a ScanNet-shape batch, which is what its 1.6 ms were. Per row there are
// only 27 coarse cells (8 for a 3^3 map) and the 8 children of a cell are one 32-byte sector of the TRANSPOSED child table
// child8[coarse row][slot]: 27 + 27 reads per row. A thread collects its K^3 entries in shared memory ([k][thread], bank =
// thread) and the block then writes the table rows k = 0 .. K^3 - 1 fully coalesced.
__global__ void child_table_kernel(const int32_t* __restrict__ nbr_down, int64_t coarse_pitch, int32_t* __restrict__ child8) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= coarse_pitch) return;
  int4 lo, hi;
  lo.x = __ldg(nbr_down + 0 * coarse_pitch + c); lo.y = __ldg(nbr_down + 1 * coarse_pitch + c);
  lo.z = __ldg(nbr_down + 2 * coarse_pitch + c); lo.w = __ldg(nbr_down + 3 * coarse_pitch + c);
  hi.x = __ldg(nbr_down + 4 * coarse_pitch + c); hi.y = __ldg(nbr_down + 5 * coarse_pitch + c);
  hi.z = __ldg(nbr_down + 6 * coarse_pitch + c); hi.w = __ldg(nbr_down + 7 * coarse_pitch + c);
  reinterpret_cast<int4*>(child8)[2 * c] = lo;
  reinterpret_cast<int4*>(child8)[2 * c + 1] = hi;
}

constexpr int kFromCoarseThreads = 128;
template <int KS>
__global__ void __launch_bounds__(kFromCoarseThreads)
kmap_from_coarse_rows_kernel(const int4* __restrict__ coords, int64_t n, int ts, const int32_t* __restrict__ parent,
                             const int32_t* __restrict__ nbr3_coarse, int64_t coarse_pitch, const int4* __restrict__ child8,
                             int32_t* __restrict__ nbr, int64_t pitch, uint32_t* __restrict__ group_mask, int words) {
  constexpr int KV = KS * KS * KS, H = KS / 2;
  extern __shared__ int32_t sh_tab[];                       // [KV][kFromCoarseThreads]
  const int t = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * kFromCoarseThreads + t;
  const bool valid = row < n;
  for (int k = 0; k < KV; ++k) sh_tab[k * kFromCoarseThreads + t] = -1;
  if (valid) {
    const int4 c = __ldg(coords + row);
    const int cs = 2 * ts;
    const int sx = (c.y - floor_to_multiple(c.y, cs)) / ts, sy = (c.z - floor_to_multiple(c.z, cs)) / ts,
              sz = (c.w - floor_to_multiple(c.w, cs)) / ts;       // child slot of this row inside its parent, 0/1 per axis
    const int32_t p = __ldg(parent + row);
#pragma unroll 1
    for (int cell = 0; cell < 27; ++cell) {
      const int qx = cell % 3 - 1, qy = (cell / 3) % 3 - 1, qz = cell / 9 - 1;
      // fine offsets that land in this coarse cell: d = 2 q + b - s with b in {0, 1}, |d| <= H, per axis
      const int dx0 = 2 * qx - sx, dy0 = 2 * qy - sy, dz0 = 2 * qz - sz;
      if (dx0 + 1 < -H || dx0 > H || dy0 + 1 < -H || dy0 > H || dz0 + 1 < -H || dz0 > H) continue;
      const int32_t pc = (cell == 13) ? p : __ldg(nbr3_coarse + (int64_t)cell * coarse_pitch + p);
      if (pc < 0) continue;
      const int4 lo = __ldg(child8 + 2 * (int64_t)pc), hi = __ldg(child8 + 2 * (int64_t)pc + 1);
      const int32_t ch[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int dx = dx0 + (b & 1), dy = dy0 + ((b >> 1) & 1), dz = dz0 + (b >> 2);
        if (dx < -H || dx > H || dy < -H || dy > H || dz < -H || dz > H) continue;
        sh_tab[((dx + H) + KS * (dy + H) + KS * KS * (dz + H)) * kFromCoarseThreads + t] = ch[b];
      }
    }
  }
  __syncthreads();                                            // (only orders this thread's own writes for the compiler)
  uint32_t m[4] = {0u, 0u, 0u, 0u};
  if (row < pitch) {
#pragma unroll 5
    for (int k = 0; k < KV; ++k) {
      const int32_t r = sh_tab[k * kFromCoarseThreads + t];
      nbr[(int64_t)k * pitch + row] = r;
      if (r >= 0) m[k >> 5] |= 1u << (k & 31);
    }
  }
  if (group_mask != nullptr) {
    // 64-row group = 2 warps of this block: OR the lanes' masks, one atomicOr per (warp, word)
    for (int wd = 0; wd < words; ++wd) {
      const uint32_t any = __reduce_or_sync(0xffffffffu, m[wd]);
      if ((t & 31) == 0 && any && row < n) atomicOr(group_mask + (row >> 6) * words + wd, any);
    }
  }
}

__global__ void kmap_count_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int64_t pitch, int32_t* __restrict__ counts) {
  const int k = blockIdx.y;
  int local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x)
    local += nbr[(int64_t)k * pitch + i] >= 0;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(counts + k, local);
}

// ---- kernel-map sorting: rows ordered by (block of rows, occupancy bit mask) so that the rows of a 128-row
// ---- MMA tile share their set of present offsets (whole (tile, offset) blocks can then be skipped) ----
__global__ void kmap_rowmask_kernel(const int32_t* __restrict__ nbr, int kvol, int64_t n_out, int64_t pitch, int block_rows,
                                    unsigned long long* __restrict__ keys, int32_t* __restrict__ idx) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  unsigned int m = 0;
  for (int k = 0; k < kvol; ++k) m |= (unsigned int)(__ldg(nbr + (int64_t)k * pitch + o) >= 0) << k;
  keys[o] = ((unsigned long long)(o / block_rows) << 32) | m;
  idx[o] = (int32_t)o;
}
// the same with the block number packed right above the kvol mask bits in a 32-bit key (block bits + kvol <= 32):
// half the key bytes and one radix pass fewer than the 64-bit form
__global__ void kmap_rowmask32_kernel(const int32_t* __restrict__ nbr, int kvol, int64_t n_out, int64_t pitch, int block_rows,
                                      unsigned int* __restrict__ keys, int32_t* __restrict__ idx) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  unsigned int m = 0;
  for (int k = 0; k < kvol; ++k) m |= (unsigned int)(__ldg(nbr + (int64_t)k * pitch + o) >= 0) << k;
  keys[o] = (kvol < 32 ? ((unsigned int)(o / block_rows) << kvol) : 0u) | m;
  idx[o] = (int32_t)o;
}

__global__ void kmap_permute_kernel(const int32_t* __restrict__ nbr, int kvol, int64_t n_out, int64_t pitch,
                                    int32_t* __restrict__ order, int32_t* __restrict__ nbr_sorted) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= pitch * kvol) return;
  const int64_t j = gid % pitch;
  const int64_t k = gid / pitch;
  if (j >= n_out) {  // padding up to the pitch: no row, no neighbour
    nbr_sorted[gid] = -1;
    if (k == 0) order[j] = -1;
    return;
  }
  nbr_sorted[gid] = __ldg(nbr + k * pitch + order[j]);
}

// group_mask[g][w]: bit (k % 32) of word (k / 32) set iff some row of the 64-row group g has offset k
__global__ void kmap_groupmask_kernel(const int32_t* __restrict__ nbr_sorted, int kvol, int64_t n_out, int64_t pitch, int words,
                                      uint32_t* __restrict__ group_mask) {
  const int64_t g = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  __shared__ uint32_t sm[4];
  if (threadIdx.x < 4) sm[threadIdx.x] = 0;
  __syncthreads();
  for (int k = warp; k < kvol; k += nwarps) {
    const int64_t r = g * 64 + lane;
    bool v = false;
    if (r < n_out) v = __ldg(nbr_sorted + (int64_t)k * pitch + r) >= 0;
    if (r + 32 < n_out) v = v || (__ldg(nbr_sorted + (int64_t)k * pitch + r + 32) >= 0);
    const unsigned any = __ballot_sync(0xffffffffu, v);
    if (lane == 0 && any) atomicOr(&sm[k >> 5], 1u << (k & 31));
  }
  __syncthreads();
  if (threadIdx.x < words) group_mask[g * words + threadIdx.x] = sm[threadIdx.x];
}

__global__ void iota_kernel(int32_t* __restrict__ out, int64_t n, int64_t pitch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < pitch) out[i] = (i < n) ? (int32_t)i : -1;
}

struct DownsampleWs {
  size_t off_keys_in, off_keys_out, off_idx_in, off_idx_out, off_flags, off_rank, off_status, off_cub, cub_bytes, total;
};
static DownsampleWs downsample_ws(int64_t n) {
  DownsampleWs w;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t off = 0;
  w.off_keys_in = off;  off += al(8 * (size_t)n);
  w.off_keys_out = off; off += al(8 * (size_t)n);
  w.off_idx_in = off;   off += al(4 * (size_t)n);
  w.off_idx_out = off;  off += al(4 * (size_t)n);
  w.off_flags = off;    off += al(4 * (size_t)n);
  w.off_rank = off;     off += al(4 * (size_t)n);
  w.off_status = off;   off += 256;
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  w.cub_bytes = al(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
  w.off_cub = off;      off += w.cub_bytes;
  w.total = off;
  return w;
}

}  // namespace b2m

using namespace b2m;

extern "C" int64_t b2m_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

extern "C" int b2m_hash_build(const int32_t* coords, int64_t n, uint64_t* table_keys, int32_t* table_vals,
                              int64_t capacity, int32_t* status, b2m_stream_t stream) {
  if ((!coords && n > 0) || !table_keys || !table_vals || !status || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (capacity < 2 * n || (capacity & (capacity - 1)) != 0) return B2M_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(table_keys, 0xFF, (size_t)capacity * 8, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (cudaMemsetAsync(table_vals, 0x7F, (size_t)capacity * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (cudaMemsetAsync(status, 0, 8, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (n == 0) return B2M_OK;
  hash_insert_kernel<<<cdiv(n, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n,
                                                  reinterpret_cast<unsigned long long*>(table_keys), table_vals,
                                                  (uint64_t)(capacity - 1), status);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_hash_query(const int32_t* query_coords, int64_t m, const uint64_t* table_keys,
                              const int32_t* table_vals, int64_t capacity, int32_t* rows, b2m_stream_t stream) {
  if (m == 0) return B2M_OK;
  if (!query_coords || !table_keys || !table_vals || !rows || m < 0) return B2M_ERR_INVALID_ARGUMENT;
  hash_query_kernel<<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(query_coords), m, reinterpret_cast<const unsigned long long*>(table_keys),
      table_vals, (uint64_t)(capacity - 1), rows);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" size_t b2m_downsample_workspace_bytes(int64_t n) { return downsample_ws(n < 1 ? 1 : n).total; }

extern "C" int b2m_downsample_coords(const int32_t* coords, int64_t n, int32_t new_stride, int32_t* out_coords,
                                     int32_t* parent_row, int32_t* n_out, void* workspace, size_t workspace_bytes,
                                     b2m_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 && n_out && new_stride > 0) {
    if (cudaMemsetAsync(n_out, 0, 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
    return B2M_OK;
  }
  if (!coords || !out_coords || !parent_row || !n_out || !workspace || n < 0 || new_stride <= 0)
    return B2M_ERR_INVALID_ARGUMENT;
  if (n >= (int64_t)1 << 31) return B2M_ERR_UNSUPPORTED_SHAPE;
  const DownsampleWs w = downsample_ws(n);
  if (workspace_bytes < w.total) return B2M_ERR_WORKSPACE_TOO_SMALL;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  auto* keys_in = reinterpret_cast<unsigned long long*>(ws + w.off_keys_in);
  auto* keys_out = reinterpret_cast<unsigned long long*>(ws + w.off_keys_out);
  auto* idx_in = reinterpret_cast<int32_t*>(ws + w.off_idx_in);
  auto* idx_out = reinterpret_cast<int32_t*>(ws + w.off_idx_out);
  auto* flags = reinterpret_cast<int32_t*>(ws + w.off_flags);
  auto* rank = reinterpret_cast<int32_t*>(ws + w.off_rank);
  auto* status = reinterpret_cast<int32_t*>(ws + w.off_status);
  if (cudaMemsetAsync(status, 0, 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  const int blocks = cdiv(n, 256);
  downsample_keys_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, new_stride, keys_in, idx_in, status);
  B2M_CHECK_LAUNCH();
  size_t cub_bytes = w.cub_bytes;
  if (cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, 64, st) != cudaSuccess)
    return B2M_ERR_CUDA_LAUNCH;
  head_flags_kernel<<<blocks, 256, 0, st>>>(keys_out, n, flags);
  B2M_CHECK_LAUNCH();
  cub_bytes = w.cub_bytes;
  if (cub::DeviceScan::InclusiveSum(ws + w.off_cub, cub_bytes, flags, rank, (int)n, st) != cudaSuccess)
    return B2M_ERR_CUDA_LAUNCH;
  downsample_emit_kernel<<<blocks, 256, 0, st>>>(keys_out, idx_out, rank, n, reinterpret_cast<int4*>(out_coords), parent_row, n_out);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int64_t b2m_map_pitch(int64_t n) { return (n + 127) / 128 * 128; }

extern "C" int b2m_kernel_map_submanifold(const int32_t* coords, int64_t n, int32_t tensor_stride, int32_t kernel_size,
                                          const uint64_t* table_keys, const int32_t* table_vals, int64_t capacity,
                                          int32_t* nbr, b2m_stream_t stream) {
  if (kernel_size != 1 && kernel_size != 3 && kernel_size != 5) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  if (!coords || !table_keys || !table_vals || !nbr || n < 0 || tensor_stride <= 0) return B2M_ERR_INVALID_ARGUMENT;
  const int64_t pitch = b2m_map_pitch(n);
  const int64_t total = pitch * kernel_size * kernel_size * kernel_size;
  kmap_submanifold_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(coords), n, tensor_stride, kernel_size,
      reinterpret_cast<const unsigned long long*>(table_keys), table_vals, (uint64_t)(capacity - 1), nbr, pitch);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" size_t b2m_kernel_map_from_coarse_workspace_bytes(int64_t n_coarse) {
  return n_coarse > 0 ? (size_t)b2m_map_pitch(n_coarse) * 32 : 0;
}

extern "C" int b2m_kernel_map_from_coarse(const int32_t* coords, int64_t n, int32_t tensor_stride, int32_t kernel_size,
                                          const int32_t* parent_row, const int32_t* nbr3_coarse, const int32_t* nbr_down,
                                          int64_t n_coarse, int32_t* nbr, uint32_t* group_mask, void* workspace,
                                          size_t workspace_bytes, b2m_stream_t stream) {
  if (kernel_size != 3 && kernel_size != 5) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  if (!coords || !parent_row || !nbr3_coarse || !nbr_down || !nbr || n < 0 || n_coarse <= 0 || tensor_stride <= 0)
    return B2M_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int kvol = kernel_size * kernel_size * kernel_size;
  const int words = (kvol + 31) / 32;
  const int64_t pitch = b2m_map_pitch(n), coarse_pitch = b2m_map_pitch(n_coarse);
  if (group_mask)
    if (cudaMemsetAsync(group_mask, 0, (size_t)((n + 63) / 64) * words * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (workspace && workspace_bytes >= (size_t)coarse_pitch * 32 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
    // one thread per row over the transposed child table (27 + 27 table reads per row instead of 2 per entry)
    int32_t* child8 = reinterpret_cast<int32_t*>(workspace);
    child_table_kernel<<<cdiv(coarse_pitch, 256), 256, 0, st>>>(nbr_down, coarse_pitch, child8);
    B2M_CHECK_LAUNCH();
    const unsigned blocks = (unsigned)(pitch / kFromCoarseThreads);
    const size_t sh = (size_t)kvol * kFromCoarseThreads * 4;
    if (kernel_size == 5) {
      static bool attr5[64];
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < 64 && !attr5[dev]) {
        if (cudaFuncSetAttribute(kmap_from_coarse_rows_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh) != cudaSuccess)
          return B2M_ERR_CUDA_LAUNCH;
        attr5[dev] = true;
      }
      kmap_from_coarse_rows_kernel<5><<<blocks, kFromCoarseThreads, sh, st>>>(
          reinterpret_cast<const int4*>(coords), n, tensor_stride, parent_row, nbr3_coarse, coarse_pitch,
          reinterpret_cast<const int4*>(child8), nbr, pitch, group_mask, words);
    } else {
      kmap_from_coarse_rows_kernel<3><<<blocks, kFromCoarseThreads, sh, st>>>(
          reinterpret_cast<const int4*>(coords), n, tensor_stride, parent_row, nbr3_coarse, coarse_pitch,
          reinterpret_cast<const int4*>(child8), nbr, pitch, group_mask, words);
    }
    B2M_CHECK_LAUNCH();
    return B2M_OK;
  }
  kmap_from_coarse_kernel<<<cdiv(pitch * kvol, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, tensor_stride,
                                                                  kernel_size, parent_row, nbr3_coarse, nbr_down,
                                                                  coarse_pitch, nbr, pitch, group_mask, words);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_kernel_map_stride2(const int32_t* fine_coords, int64_t n_fine, const int32_t* parent_row,
                                      int64_t n_coarse, int32_t fine_stride, int32_t* nbr_down, int32_t* nbr_up,
                                      b2m_stream_t stream) {
  if (!fine_coords || !parent_row || n_fine < 0 || n_coarse < 0 || fine_stride <= 0) return B2M_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t coarse_pitch = b2m_map_pitch(n_coarse), fine_pitch = b2m_map_pitch(n_fine);
  if (nbr_down && n_coarse > 0)
    if (cudaMemsetAsync(nbr_down, 0xFF, (size_t)8 * coarse_pitch * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (n_fine == 0) return B2M_OK;
  kmap_stride2_kernel<<<cdiv(fine_pitch, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(fine_coords), n_fine, parent_row,
                                                            coarse_pitch, fine_stride, nbr_down, nbr_up, fine_pitch);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_kernel_map_count(const int32_t* nbr, int32_t kvol, int64_t n_out, int32_t* counts, b2m_stream_t stream) {
  if (!nbr || !counts || kvol <= 0 || n_out < 0) return B2M_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, (size_t)kvol * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (n_out == 0) return B2M_OK;
  int bx = cdiv(n_out, 256 * 8);
  if (bx > 1024) bx = 1024;
  kmap_count_kernel<<<dim3(bx, kvol), 256, 0, st>>>(nbr, n_out, b2m_map_pitch(n_out), counts);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" size_t b2m_kernel_map_sort_workspace_bytes(int64_t n_out) { return downsample_ws(n_out < 1 ? 1 : n_out).total; }

extern "C" int b2m_kernel_map_sort(const int32_t* nbr, int32_t kvol, int64_t n_out, int32_t block_rows, int32_t* order,
                                   int32_t* nbr_sorted, uint32_t* group_mask, void* workspace, size_t workspace_bytes,
                                   b2m_stream_t stream) {
  if (n_out == 0) return B2M_OK;
  if (!nbr || !order || !nbr_sorted || !group_mask || kvol <= 0 || kvol > 128 || n_out < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (n_out >= (int64_t)1 << 31) return B2M_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (kvol + 31) / 32;
  const int64_t pitch = b2m_map_pitch(n_out);
  const int blocks = cdiv(pitch, 256);
  const bool do_sort = block_rows > 0 && kvol <= 32;
  if (do_sort) {
    if (!workspace) return B2M_ERR_INVALID_ARGUMENT;
    const DownsampleWs w = downsample_ws(n_out);
    if (workspace_bytes < w.total) return B2M_ERR_WORKSPACE_TOO_SMALL;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    auto* keys_in = reinterpret_cast<unsigned long long*>(ws + w.off_keys_in);
    auto* keys_out = reinterpret_cast<unsigned long long*>(ws + w.off_keys_out);
    auto* idx_in = reinterpret_cast<int32_t*>(ws + w.off_idx_in);
    int64_t nblk = (n_out + block_rows - 1) / block_rows;
    int blk_bits = 0;
    while (((int64_t)1 << blk_bits) < nblk) ++blk_bits;
    size_t cub_bytes = w.cub_bytes;
    if (blk_bits + kvol <= 32) {
      // (block, mask) fits 32 bits: radix-sort only the significant bits of a 32-bit key
      auto* k32_in = reinterpret_cast<unsigned int*>(keys_in);
      auto* k32_out = reinterpret_cast<unsigned int*>(keys_out);
      kmap_rowmask32_kernel<<<blocks, 256, 0, st>>>(nbr, kvol, n_out, pitch, block_rows, k32_in, idx_in);
      B2M_CHECK_LAUNCH();
      if (cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, k32_in, k32_out, idx_in, order, (int)n_out, 0,
                                          blk_bits + kvol, st) != cudaSuccess)
        return B2M_ERR_CUDA_LAUNCH;
    } else {
      kmap_rowmask_kernel<<<blocks, 256, 0, st>>>(nbr, kvol, n_out, pitch, block_rows, keys_in, idx_in);
      B2M_CHECK_LAUNCH();
      if (cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, keys_in, keys_out, idx_in, order, (int)n_out, 0,
                                          32 + blk_bits, st) != cudaSuccess)
        return B2M_ERR_CUDA_LAUNCH;
    }
    kmap_permute_kernel<<<cdiv(pitch * kvol, 256), 256, 0, st>>>(nbr, kvol, n_out, pitch, order, nbr_sorted);
    B2M_CHECK_LAUNCH();
  } else {
    iota_kernel<<<blocks, 256, 0, st>>>(order, n_out, pitch);
    B2M_CHECK_LAUNCH();
    if (nbr_sorted != nbr)
      if (cudaMemcpyAsync(nbr_sorted, nbr, (size_t)kvol * pitch * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  }
  kmap_groupmask_kernel<<<(unsigned)((n_out + 63) / 64), 256, 0, st>>>(nbr_sorted, kvol, n_out, pitch, words, group_mask);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_version(void) { return 3; }

extern "C" const char* b2m_error_string(int code) {
  switch (code) {
    case B2M_OK: return "ok";
    case B2M_ERR_INVALID_ARGUMENT: return "invalid argument";
    case B2M_ERR_CUDA_LAUNCH: return "CUDA launch or runtime error";
    case B2M_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case B2M_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
    case B2M_ERR_COORD_RANGE: return "coordinate outside the packable range";
    default: return "unknown error";
  }
}
