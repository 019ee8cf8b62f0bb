// Box-vote decoding: axis-aligned 3-D IoU NMS clustering with heat-maps, heat-map -> voxel-mask
// projection, and mask NMS on bit-packed masks.
//
// Bit-exact restatement for the GPU of /root/reference/models/iou_nms.py:26-45 (torch_IOUs),
// :68-105 (NMS_clustering), :109-144 (masks_iou, mask_NMS) and models/detection_net.py:436-446.
// All fp32 arithmetic uses the explicit round-to-nearest intrinsics so that nvcc can not contract
// multiplies and adds into FMAs; the operation order is the one torch's CPU kernels use
// (prod over 3 elements = (a*b)*c; union = ((A_r + A_j) - I) + 1e-6f; IEEE division).
#include "common.cuh"

namespace b2m {

struct Box { float x0, y0, z0, x1, y1, z1; };

__device__ __forceinline__ Box load_box(const float* __restrict__ boxes, int i) {
  const float* p = boxes + (int64_t)i * 7;
  Box b;
  b.x0 = p[1]; b.y0 = p[2]; b.z0 = p[3]; b.x1 = p[4]; b.y1 = p[5]; b.z1 = p[6];
  return b;
}
__device__ __forceinline__ float box_volume(const Box& b) {
  return __fmul_rn(__fmul_rn(__fsub_rn(b.x1, b.x0), __fsub_rn(b.y1, b.y0)), __fsub_rn(b.z1, b.z0));
}
// IoU(r, j) exactly as torch_IOUs(box_r, boxes)[j]
__device__ __forceinline__ float aabb_iou(const Box& r, float vol_r, const Box& j) {
  const float dx = fmaxf(__fsub_rn(fminf(r.x1, j.x1), fmaxf(r.x0, j.x0)), 0.f);
  const float dy = fmaxf(__fsub_rn(fminf(r.y1, j.y1), fmaxf(r.y0, j.y0)), 0.f);
  const float dz = fmaxf(__fsub_rn(fminf(r.z1, j.z1), fmaxf(r.z0, j.z0)), 0.f);
  const float inter = __fmul_rn(__fmul_rn(dx, dy), dz);
  const float uni = __fadd_rn(__fsub_rn(__fadd_rn(vol_r, box_volume(j)), inter), 0.000001f);
  return __fdiv_rn(inter, uni);
}

// order[rank] = box index, descending score, ties by lower index (stable argsort of -score)
__global__ void nms_rank_kernel(const float* __restrict__ boxes, int m, int32_t* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float si = boxes[(int64_t)i * 7];
  int rank = 0;
  for (int j = 0; j < m; ++j) {
    const float sj = __ldg(boxes + (int64_t)j * 7);
    rank += (sj > si) || (sj == si && j < i);
  }
  order[rank] = i;
}

// bit (p, q) = IoU(box[order[p]], box[order[q]]) > th  (p == q forced to 1), in sorted positions
__global__ void nms_mask_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ order, int m, int words,
                                float th, uint32_t* __restrict__ mask) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (w >= words) return;
  const Box r = load_box(boxes, order[p]);
  const float vr = box_volume(r);
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    const int q = w * 32 + b;
    if (q >= m) break;
    bool on = (q == p);
    if (!on) on = !(aabb_iou(r, vr, load_box(boxes, order[q])) <= th);  // members = ~(iou <= th), iou_nms.py:99-101
    bits |= (on ? 1u : 0u) << b;
  }
  mask[(int64_t)p * words + w] = bits;
}

// Greedy scan in one block: the remaining set is a bit vector in shared memory.
__global__ void __launch_bounds__(1024)
nms_scan_kernel(const uint32_t* __restrict__ mask, const int32_t* __restrict__ order, int m, int words,
                int32_t* __restrict__ n_clusters, int32_t* __restrict__ reps, int32_t* __restrict__ cluster_of) {
  extern __shared__ uint32_t remaining[];  // [words]
  __shared__ int next_p;
  for (int w = threadIdx.x; w < words; w += blockDim.x) {
    const int lo = w * 32;
    remaining[w] = (m - lo >= 32) ? 0xFFFFFFFFu : ((m > lo) ? ((1u << (m - lo)) - 1u) : 0u);
  }
  if (threadIdx.x == 0) next_p = (m > 0) ? 0 : -1;
  __syncthreads();
  int nc = 0;
  while (true) {
    const int p = next_p;
    if (p < 0) break;
    __syncthreads();  // everyone has read next_p
    if (threadIdx.x == 0) { reps[nc] = order[p]; next_p = 0x7FFFFFFF; }
    __syncthreads();
    int local_next = 0x7FFFFFFF;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
      const uint32_t rem = remaining[w];
      const uint32_t members = rem & mask[(int64_t)p * words + w];
      uint32_t mm = members;
      while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        cluster_of[order[w * 32 + b]] = nc;
      }
      const uint32_t left = rem & ~members;
      remaining[w] = left;
      if (left && local_next == 0x7FFFFFFF) local_next = w * 32 + __ffs(left) - 1;
    }
    if (local_next != 0x7FFFFFFF) atomicMin(&next_p, local_next);
    __syncthreads();
    if (threadIdx.x == 0 && next_p == 0x7FFFFFFF) next_p = -1;
    ++nc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_clusters = nc;
}

// heat[c][j] = IoU(box[reps[c]], box[j]) in ORIGINAL box order, self = 1
__global__ void nms_heat_kernel(const float* __restrict__ boxes, int m, const int32_t* __restrict__ n_clusters,
                                const int32_t* __restrict__ reps, int64_t max_clusters, float* __restrict__ heat) {
  const int c = blockIdx.y;
  if (c >= *n_clusters || c >= max_clusters) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int rep = reps[c];
  const Box r = load_box(boxes, rep);
  float v = aabb_iou(r, box_volume(r), load_box(boxes, j));
  if (j == rep) v = 1.f;
  heat[(int64_t)c * m + j] = v;
}

__global__ void heat_project_kernel(const float* __restrict__ heat, int64_t k, int64_t m_fg,
                                    const int32_t* __restrict__ fg_rank, const int64_t* __restrict__ seg2vox,
                                    int64_t n_vox, int64_t words, float th, uint32_t* __restrict__ masks) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (w >= words || c >= k) return;
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    const int64_t v = w * 32 + b;
    if (v >= n_vox) break;
    const int32_t r = __ldg(fg_rank + __ldg(seg2vox + v));
    if (r >= 0 && __ldg(heat + c * m_fg + r) > th) bits |= 1u << b;
  }
  masks[c * words + w] = bits;
}

// inter[a][b] = |mask_a & mask_b| for a <= b, area[a] = |mask_a|
__global__ void mask_inter_kernel(const uint32_t* __restrict__ masks, int64_t k, int64_t words, int32_t* __restrict__ inter,
                                  int32_t* __restrict__ area) {
  const int a = blockIdx.y, b = blockIdx.x;
  if (b < a) return;
  const uint32_t* ma = masks + (int64_t)a * words;
  const uint32_t* mb = masks + (int64_t)b * words;
  int local = 0;
  for (int64_t w = threadIdx.x; w < words; w += blockDim.x) local += __popc(__ldg(ma + w) & __ldg(mb + w));
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    inter[(int64_t)a * k + b] = t;
    inter[(int64_t)b * k + a] = t;
    if (a == b) area[a] = t;
  }
}

__global__ void mask_scan_kernel(const int32_t* __restrict__ inter, const int32_t* __restrict__ area, int k, float th,
                                 uint8_t* __restrict__ keep, int32_t* __restrict__ n_keep) {
  extern __shared__ uint8_t alive[];  // [k]
  for (int i = threadIdx.x; i < k; i += blockDim.x) alive[i] = 1;
  __syncthreads();
  int nk = 0;
  for (int a = 0; a < k; ++a) {
    const bool is_kept = alive[a] != 0;  // uniform across the block
    __syncthreads();
    if (is_kept) {
      ++nk;
      for (int b = a + 1 + threadIdx.x; b < k; b += blockDim.x) {
        if (!alive[b]) continue;
        const int i = inter[(int64_t)a * k + b];
        const int u = area[a] + area[b] - i;
        const float iou = __fdiv_rn((float)i, (float)u);
        if (!(iou <= th)) alive[b] = 0;  // suppressed = ~(iou <= th), iou_nms.py:137-141
      }
    }
    if (threadIdx.x == 0) keep[a] = is_kept ? 1 : 0;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_keep = nk;
}

__global__ void unpack_masks_kernel(const uint32_t* __restrict__ masks, int64_t k, int64_t words, int64_t n_vox,
                                    uint8_t* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (v >= n_vox || c >= k) return;
  out[c * n_vox + v] = (masks[c * words + (v >> 5)] >> (v & 31)) & 1u;
}

}  // namespace b2m

using namespace b2m;

extern "C" size_t b2m_nms_workspace_bytes(int64_t m) {
  const int64_t words = (m + 31) / 32;
  size_t order = ((size_t)m * 4 + 255) / 256 * 256;
  return order + (size_t)m * words * 4 + 256;
}

extern "C" int b2m_aabb_nms(const float* boxes, int64_t m, float cluster_th, int32_t* n_clusters, int32_t* representatives,
                            int32_t* cluster_of, float* heatmaps, int64_t max_clusters, void* workspace,
                            size_t workspace_bytes, b2m_stream_t stream) {
  if (!boxes || !n_clusters || !representatives || !cluster_of || !workspace || m < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!(cluster_th > 0.f && cluster_th < 1.f)) return B2M_ERR_INVALID_ARGUMENT;  // models/iou_nms.py:71
  if (m > 65535 * 32) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (workspace_bytes < b2m_nms_workspace_bytes(m)) return B2M_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) {
    if (cudaMemsetAsync(n_clusters, 0, 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
    return B2M_OK;
  }
  const int mi = (int)m;
  const int words = (mi + 31) / 32;
  if ((size_t)words * 4 > 200 * 1024) return B2M_ERR_UNSUPPORTED_SHAPE;
  int32_t* order = reinterpret_cast<int32_t*>(workspace);
  uint32_t* mask = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + ((size_t)m * 4 + 255) / 256 * 256);
  nms_rank_kernel<<<cdiv(mi, 128), 128, 0, st>>>(boxes, mi, order);
  B2M_CHECK_LAUNCH();
  nms_mask_kernel<<<dim3(cdiv(words, 64), mi), 64, 0, st>>>(boxes, order, mi, words, cluster_th, mask);
  B2M_CHECK_LAUNCH();
  const size_t sh = (size_t)words * 4;
  if (sh > 40 * 1024) {
    if (cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh) != cudaSuccess)
      return B2M_ERR_CUDA_LAUNCH;
  }
  nms_scan_kernel<<<1, 1024, sh, st>>>(mask, order, mi, words, n_clusters, representatives, cluster_of);
  B2M_CHECK_LAUNCH();
  if (heatmaps && max_clusters > 0) {
    const int64_t rows = max_clusters < m ? max_clusters : m;
    nms_heat_kernel<<<dim3(cdiv(mi, 128), (unsigned)rows), 128, 0, st>>>(boxes, mi, n_clusters, representatives, max_clusters, heatmaps);
    B2M_CHECK_LAUNCH();
  }
  return B2M_OK;
}

extern "C" int b2m_heatmap_project(const float* heat, int64_t k, int64_t m_fg, const int32_t* fg_rank,
                                   const int64_t* seg2vox, int64_t n_vox, float mask_bin_th, uint32_t* masks,
                                   b2m_stream_t stream) {
  if (!heat || !fg_rank || !seg2vox || !masks || k < 0 || n_vox < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (k == 0 || n_vox == 0) return B2M_OK;
  if (k > 65535) return B2M_ERR_UNSUPPORTED_SHAPE;
  const int64_t words = (n_vox + 31) / 32;
  heat_project_kernel<<<dim3(cdiv(words, 128), (unsigned)k), 128, 0, (cudaStream_t)stream>>>(heat, k, m_fg, fg_rank, seg2vox,
                                                                                            n_vox, words, mask_bin_th, masks);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" size_t b2m_mask_nms_workspace_bytes(int64_t k) { return (size_t)(k * k + k) * 4 + 256; }

extern "C" int b2m_mask_nms(const uint32_t* masks, int64_t k, int64_t words, float th, uint8_t* keep, int32_t* n_keep,
                            void* workspace, size_t workspace_bytes, b2m_stream_t stream) {
  if (!masks || !keep || !n_keep || !workspace || k < 0 || words < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < b2m_mask_nms_workspace_bytes(k)) return B2M_ERR_WORKSPACE_TOO_SMALL;
  if (k > 65535) return B2M_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 0) {
    if (cudaMemsetAsync(n_keep, 0, 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
    return B2M_OK;
  }
  int32_t* inter = reinterpret_cast<int32_t*>(workspace);
  int32_t* area = inter + k * k;
  mask_inter_kernel<<<dim3((unsigned)k, (unsigned)k), 128, 0, st>>>(masks, k, words, inter, area);
  B2M_CHECK_LAUNCH();
  mask_scan_kernel<<<1, 256, (size_t)k, st>>>(inter, area, (int)k, th, keep, n_keep);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_unpack_masks(const uint32_t* masks, int64_t k, int64_t words, int64_t n_vox, uint8_t* out,
                                b2m_stream_t stream) {
  if (!masks || !out || k < 0 || n_vox < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (k == 0 || n_vox == 0) return B2M_OK;
  unpack_masks_kernel<<<dim3(cdiv(n_vox, 256), (unsigned)k), 256, 0, (cudaStream_t)stream>>>(masks, k, words, n_vox, out);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
