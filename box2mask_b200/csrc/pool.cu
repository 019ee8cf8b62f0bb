// Superpoint pooling: segmented mean / max over voxel rows keyed by a dense superpoint id.
//
// Replaces the re-keyed ME.SparseTensor + MinkowskiGlobalAvgPooling / MinkowskiGlobalMaxPooling of
// /root/reference/models/detection_net.py:345-352 (ids from utils/util.py:123-130) without building a
// second coordinate hash. HBM-bound: each voxel row is read once with 16-byte loads; the [S, C] fp32
// accumulator (a few MB) lives in L2 and takes vectorised float4 reductions.
#include "common.cuh"

namespace b2m {

// Segmented reduction of one 64-row tile per block iteration: the tile is staged in shared memory with coalesced
// 16-byte loads, then thread (column, row quarter) walks its 16 rows in order and keeps a running sum while the
// superpoint id stays the same; only a change of id (or the end of the quarter) costs a global fp32 reduction.
// Superpoints are spatial patches and rows arrive in coordinate order, so consecutive rows mostly share an id:
// the atomics drop from one per element to one per (run, column), and a warp's reductions still cover 128
// contiguous bytes of one accumulator row.
constexpr int kSegRows = 64;
constexpr int kSegQuarter = 16;
__global__ void __launch_bounds__(512)
segsum_kernel(const uint16_t* __restrict__ f, const int64_t* __restrict__ ids, int64_t n, int c,
              int64_t s, float* __restrict__ out, float* __restrict__ counts) {
  extern __shared__ uint8_t seg_smem[];
  uint16_t* tile = reinterpret_cast<uint16_t*>(seg_smem);                                   // [64][c] bf16
  int64_t* tid_s = reinterpret_cast<int64_t*>(seg_smem + (size_t)kSegRows * c * 2);         // [64]
  const int G = c / 8;
  const int64_t n_tiles = (n + kSegRows - 1) / kSegRows;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t r0 = t * kSegRows;
    const int rows = (int)min((int64_t)kSegRows, n - r0);
    for (int v = threadIdx.x; v < rows * G; v += blockDim.x)
      reinterpret_cast<uint4*>(tile)[v] = __ldg(reinterpret_cast<const uint4*>(f + r0 * c) + v);
    for (int r = threadIdx.x; r < kSegRows; r += blockDim.x) tid_s[r] = (r < rows) ? __ldg(ids + r0 + r) : -1;
    __syncthreads();
    for (int w = threadIdx.x; w < 4 * c; w += blockDim.x) {
      const int col = w % c, q = w / c;
      const int rb = q * kSegQuarter, re = min(rows, rb + kSegQuarter);
      float acc = 0.f;
      int run = 0;
      int64_t cur = -1;
      for (int r = rb; r < re; ++r) {
        const int64_t id = tid_s[r];
        if (id != cur) {
          if (run > 0 && cur >= 0 && cur < s) {
            atomicAdd(out + cur * c + col, acc);
            if (col == 0) atomicAdd(counts + cur, (float)run);
          }
          cur = id; acc = 0.f; run = 0;
        }
        acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(tile)[r * c + col]);
        ++run;
      }
      if (run > 0 && cur >= 0 && cur < s) {
        atomicAdd(out + cur * c + col, acc);
        if (col == 0) atomicAdd(counts + cur, (float)run);
      }
    }
    __syncthreads();
  }
}

__global__ void segdiv_kernel(float* __restrict__ out, const float* __restrict__ counts, int64_t s, int c) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= s * c) return;
  const float cnt = counts[gid / c];
  if (cnt > 0.f) out[gid] = out[gid] / cnt;
}

__global__ void segmean_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ ids,
                                   const float* __restrict__ counts, int64_t n, int c, int64_t s,
                                   uint16_t* __restrict__ df) {
  const int G = c / 8;
  const int64_t total = n * G;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = v / G;
    const int g = (int)(v % G);
    const int64_t id = __ldg(ids + r);
    if (id < 0 || id >= s) {      // rows the forward pass skipped (ignore ids) receive no gradient
      reinterpret_cast<uint4*>(df)[v] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const float inv = 1.f / __ldg(counts + id);
    const float4 a = __ldg(reinterpret_cast<const float4*>(dout + id * c + g * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dout + id * c + g * 8 + 4));
    uint4 u;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
    p[0] = __floats2bfloat162_rn(a.x * inv, a.y * inv);
    p[1] = __floats2bfloat162_rn(a.z * inv, a.w * inv);
    p[2] = __floats2bfloat162_rn(b.x * inv, b.y * inv);
    p[3] = __floats2bfloat162_rn(b.z * inv, b.w * inv);
    reinterpret_cast<uint4*>(df)[v] = u;
  }
}

// order-preserving float <-> int key so that integer atomicMax implements float max
__device__ __forceinline__ int float_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

__global__ void segmax_kernel(const uint16_t* __restrict__ f, const int64_t* __restrict__ ids, int64_t n, int c,
                              int64_t s, int* __restrict__ keys) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * c) return;
  const int64_t r = gid / c;
  const int64_t id = __ldg(ids + r);
  if (id < 0 || id >= s) return;
  const float v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(f)[gid]);
  atomicMax(keys + id * c + (gid % c), float_key(v));
}
__global__ void segargmax_kernel(const uint16_t* __restrict__ f, const int64_t* __restrict__ ids, int64_t n, int c,
                                 int64_t s, const int* __restrict__ keys, int32_t* __restrict__ argmax) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * c) return;
  const int64_t r = gid / c;
  const int64_t id = __ldg(ids + r);
  if (id < 0 || id >= s) return;
  const float v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(f)[gid]);
  if (float_key(v) == keys[id * c + (gid % c)]) atomicMin(argmax + id * c + (gid % c), (int32_t)r);
}
__global__ void segmax_finish_kernel(float* __restrict__ out, int64_t total) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  out[gid] = key_float(reinterpret_cast<const int*>(out)[gid]);
}

// backward of the segment max: df[argmax[s, j], j] = dout[s, j], everything else zero (df zeroed by the caller side)
__global__ void segmax_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ argmax, int64_t total, int c,
                                  int64_t n, uint16_t* __restrict__ df) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int32_t r = argmax[gid];
  if (r < 0 || r >= n) return;                       // empty segment
  reinterpret_cast<__nv_bfloat16*>(df)[(int64_t)r * c + (gid % c)] = __float2bfloat16_rn(dout[gid]);
}

}  // namespace b2m

using namespace b2m;

static int pool_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int b2m_segment_mean_forward(const uint16_t* f, const int64_t* ids, int64_t n, int32_t c, int64_t s, float* out,
                                        float* counts, b2m_stream_t stream) {
  if (!f || !ids || !out || !counts || n < 0 || s < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (c <= 0 || c % 8 != 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (s == 0) return B2M_OK;
  if (cudaMemsetAsync(out, 0, (size_t)s * c * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (cudaMemsetAsync(counts, 0, (size_t)s * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (n == 0) return B2M_OK;
  {
    const int64_t n_tiles = (n + kSegRows - 1) / kSegRows;
    const int grid = (int)(n_tiles < 148 * 4 ? n_tiles : 148 * 4);
    const size_t sh = (size_t)kSegRows * c * 2 + kSegRows * 8;
    if (sh > 48 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
    segsum_kernel<<<grid, 512, sh, st>>>(f, ids, n, c, s, out, counts);
  }
  B2M_CHECK_LAUNCH();
  segdiv_kernel<<<cdiv(s * c, 256), 256, 0, st>>>(out, counts, s, c);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_segment_mean_backward(const float* dout, const int64_t* ids, const float* counts, int64_t n, int32_t c,
                                         int64_t s, uint16_t* df, b2m_stream_t stream) {
  if (!dout || !ids || !counts || !df || n < 0 || s < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (c <= 0 || c % 8 != 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  segmean_bwd_kernel<<<pool_grid(n * (c / 8)), 256, 0, (cudaStream_t)stream>>>(dout, ids, counts, n, c, s, df);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_segment_max_forward(const uint16_t* f, const int64_t* ids, int64_t n, int32_t c, int64_t s, float* out,
                                       int32_t* argmax, b2m_stream_t stream) {
  if (!f || !ids || !out || !argmax || n < 0 || s < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (c <= 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (s == 0) return B2M_OK;
  // keys start at the key of -inf (0x807FFFFF ^ ... ) : use the minimum int so that any value wins
  if (cudaMemsetAsync(out, 0x80, (size_t)s * c * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (cudaMemsetAsync(argmax, 0x7F, (size_t)s * c * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (n == 0) return B2M_OK;
  segmax_kernel<<<cdiv(n * c, 256), 256, 0, st>>>(f, ids, n, c, s, reinterpret_cast<int*>(out));
  B2M_CHECK_LAUNCH();
  segargmax_kernel<<<cdiv(n * c, 256), 256, 0, st>>>(f, ids, n, c, s, reinterpret_cast<const int*>(out), argmax);
  B2M_CHECK_LAUNCH();
  segmax_finish_kernel<<<cdiv(s * c, 256), 256, 0, st>>>(out, s * c);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

/* df bf16[n, c]: gradient of b2m_segment_max_forward (a row receives dout[s, j] iff it attained the maximum of (s, j)) */
extern "C" int b2m_segment_max_backward(const float* dout, const int32_t* argmax, int64_t s, int32_t c, int64_t n,
                                        uint16_t* df, b2m_stream_t stream) {
  if (!dout || !argmax || !df || n < 0 || s < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (c <= 0) return B2M_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return B2M_OK;
  if (cudaMemsetAsync(df, 0, (size_t)n * c * 2, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (s == 0) return B2M_OK;
  segmax_bwd_kernel<<<cdiv(s * c, 256), 256, 0, st>>>(dout, argmax, s * c, c, n, df);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
