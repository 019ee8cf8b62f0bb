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

// Greedy scan by ONE warp: the remaining set lives in registers (kScanWords words per lane, lane l owns words
// l, l + 32, ...), so a cluster costs one ballot to find the first remaining box, one coalesced read of its mask row
// and an and-not: no block barriers, no shared memory; the dependent chain per cluster is one L2 read (~0.5 us).
// Outputs the representatives as sorted positions; box indices and memberships follow in nms_assign_kernel.
// Instantiated per words-per-lane count so that the loops over a lane's words are exactly as long as the box count needs
// (M <= 2048: two words, ~20 instructions + one L2 read per cluster).
constexpr int kScanWords = 64;      // most words per lane (template parameter WPL = ceil(words / 32) rounded up): up to 32 * 64 * 32 = 65536 boxes
template <int WPL>
__global__ void __launch_bounds__(32)
nms_scan_kernel(const uint32_t* __restrict__ mask, int m, int words, int32_t* __restrict__ n_clusters,
                int32_t* __restrict__ rep_pos) {
  const int lane = threadIdx.x;
  uint32_t rem[WPL];
#pragma unroll
  for (int i = 0; i < WPL; ++i) {
    const int w = i * 32 + lane;
    const int lo = w * 32;
    rem[i] = (w < words) ? ((m - lo >= 32) ? 0xFFFFFFFFu : ((m > lo) ? ((1u << (m - lo)) - 1u) : 0u)) : 0u;
  }
  // first remaining position at or after nothing: min over lanes of the first set bit
  auto first_remaining = [&]() -> int {
    int best = 0x7FFFFFFF;
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      if (rem[i] && best == 0x7FFFFFFF) best = (i * 32 + lane) * 32 + __ffs(rem[i]) - 1;
    }
    return __reduce_min_sync(0xffffffffu, best);
  };
  int nc = 0;
  int p = first_remaining();
  while (p != 0x7FFFFFFF) {
    if (lane == 0) rep_pos[nc] = p;
    ++nc;
    const uint32_t* row = mask + (int64_t)p * words;
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      const int w = i * 32 + lane;
      if (w < words) rem[i] &= ~__ldg(row + w);            // members = remaining & row (incl. p itself: diagonal bit)
    }
    p = first_remaining();
  }
  if (lane == 0) *n_clusters = nc;
}

// reps[c] = box index of representative c; cluster_of[box] = first cluster (in greedy order) whose representative's
// mask row has the box's bit: exactly the greedy membership, because a box leaves the remaining set at the first such row.
__global__ void nms_assign_kernel(const uint32_t* __restrict__ mask, const int32_t* __restrict__ order, int m, int words,
                                  const int32_t* __restrict__ n_clusters, const int32_t* __restrict__ rep_pos,
                                  int32_t* __restrict__ reps, int32_t* __restrict__ cluster_of) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;    // sorted position of the box
  const int nc = *n_clusters;
  if (q < nc) reps[q] = order[rep_pos[q]];
  if (q >= m) return;
  const int w = q >> 5;
  const uint32_t bit = 1u << (q & 31);
  int c = 0;
  for (; c < nc; ++c)
    if (__ldg(mask + (int64_t)__ldg(rep_pos + c) * words + w) & bit) break;
  cluster_of[order[q]] = c;      // c < nc always: a box that is never suppressed becomes a representative itself
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

// counts[c][l] = number of voxels v in mask c with label[v] == l (the np.bincount of models/detection_net.py:461-466,
// on the bit-packed masks): a block takes one mask and a slice of its words, histograms in shared memory.
__global__ void __launch_bounds__(256)
mask_label_hist_kernel(const uint32_t* __restrict__ masks, int64_t words, int64_t n_vox, const int32_t* __restrict__ label,
                       int n_labels, int32_t* __restrict__ counts) {
  extern __shared__ int hist[];   // [n_labels]
  for (int i = threadIdx.x; i < n_labels; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int64_t c = blockIdx.y;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t bits = __ldg(masks + c * words + w);
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int64_t v = w * 32 + b;
      if (v < n_vox) {
        const int l = __ldg(label + v);
        if (l >= 0 && l < n_labels) atomicAdd(&hist[l], 1);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_labels; i += blockDim.x)
    if (hist[i]) atomicAdd(counts + c * n_labels + i, hist[i]);
}
// counts[seg[v]][label[v]] += 1 (per-segment label histogram: the torch.mode of models/detection_net.py:404-410)
__global__ void segment_label_hist_kernel(const int64_t* __restrict__ seg, const int32_t* __restrict__ label, int64_t n,
                                          int64_t n_seg, int n_labels, int32_t* __restrict__ counts) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int64_t s = __ldg(seg + v);
  const int l = __ldg(label + v);
  if (s >= 0 && s < n_seg && l >= 0 && l < n_labels) atomicAdd(counts + s * n_labels + l, 1);
}
// best[r] = smallest label with the largest count (np.argmax / torch.mode tie rule: lowest value wins)
__global__ void hist_argmax_kernel(const int32_t* __restrict__ counts, int64_t rows, int n_labels, int32_t* __restrict__ best) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int bl = 0, bc = counts[r * n_labels];
  for (int l = 1; l < n_labels; ++l) {
    const int cnt = counts[r * n_labels + l];
    if (cnt > bc) { bc = cnt; bl = l; }
  }
  best[r] = bl;
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
  return order + ((size_t)m * words * 4 + 255) / 256 * 256 + order + 256;   // order | pair mask | representative positions
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
  if (words > 32 * kScanWords) return B2M_ERR_UNSUPPORTED_SHAPE;
  int32_t* order = reinterpret_cast<int32_t*>(workspace);
  uint32_t* mask = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + ((size_t)m * 4 + 255) / 256 * 256);
  nms_rank_kernel<<<cdiv(mi, 128), 128, 0, st>>>(boxes, mi, order);
  B2M_CHECK_LAUNCH();
  nms_mask_kernel<<<dim3(cdiv(words, 64), mi), 64, 0, st>>>(boxes, order, mi, words, cluster_th, mask);
  B2M_CHECK_LAUNCH();
  int32_t* rep_pos = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) + ((size_t)m * 4 + 255) / 256 * 256 +
                                                ((size_t)m * words * 4 + 255) / 256 * 256);
  const int wpl = (words + 31) / 32;
  if (wpl <= 1) nms_scan_kernel<1><<<1, 32, 0, st>>>(mask, mi, words, n_clusters, rep_pos);
  else if (wpl <= 2) nms_scan_kernel<2><<<1, 32, 0, st>>>(mask, mi, words, n_clusters, rep_pos);
  else if (wpl <= 4) nms_scan_kernel<4><<<1, 32, 0, st>>>(mask, mi, words, n_clusters, rep_pos);
  else if (wpl <= 8) nms_scan_kernel<8><<<1, 32, 0, st>>>(mask, mi, words, n_clusters, rep_pos);
  else if (wpl <= 16) nms_scan_kernel<16><<<1, 32, 0, st>>>(mask, mi, words, n_clusters, rep_pos);
  else nms_scan_kernel<kScanWords><<<1, 32, 0, st>>>(mask, mi, words, n_clusters, rep_pos);
  B2M_CHECK_LAUNCH();
  nms_assign_kernel<<<cdiv(mi, 128), 128, 0, st>>>(mask, order, mi, words, n_clusters, rep_pos, representatives, cluster_of);
  B2M_CHECK_LAUNCH();
  if (heatmaps && max_clusters > 0) {
    const int64_t rows = max_clusters < m ? max_clusters : m;
    nms_heat_kernel<<<dim3(cdiv(mi, 128), (unsigned)rows), 128, 0, st>>>(boxes, mi, n_clusters, representatives, max_clusters, heatmaps);
    B2M_CHECK_LAUNCH();
  }
  return B2M_OK;
}

/* Heat-map rows of the first k clusters (k read back by the caller): heat float[k, m]. Split from b2m_aabb_nms so that
 * the caller allocates k rows instead of m. */
extern "C" int b2m_aabb_heatmaps(const float* boxes, int64_t m, const int32_t* n_clusters, const int32_t* representatives,
                                 int64_t k, float* heatmaps, b2m_stream_t stream) {
  if (!boxes || !n_clusters || !representatives || !heatmaps || m < 0 || k < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (m == 0 || k == 0) return B2M_OK;
  if (k > 65535) return B2M_ERR_UNSUPPORTED_SHAPE;
  nms_heat_kernel<<<dim3(cdiv(m, 128), (unsigned)k), 128, 0, (cudaStream_t)stream>>>(boxes, (int)m, n_clusters, representatives, k,
                                                                                    heatmaps);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

/* counts int32[k, n_labels] (zeroed inside) = label histogram of every bit-packed mask; best int32[k] = its arg-max
 * (lowest label on ties) - the per-instance majority vote of models/detection_net.py:461-466. label int32[n_vox]. */
extern "C" int b2m_mask_label_vote(const uint32_t* masks, int64_t k, int64_t words, int64_t n_vox, const int32_t* label,
                                   int32_t n_labels, int32_t* counts, int32_t* best, b2m_stream_t stream) {
  if (!masks || !label || !counts || !best || k < 0 || words < 0 || n_labels <= 0) return B2M_ERR_INVALID_ARGUMENT;
  if (n_labels > 4096 || k > 65535) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (k == 0) return B2M_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, (size_t)k * n_labels * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (words > 0) {
    int gx = cdiv(words, 256 * 8);
    if (gx < 1) gx = 1;
    mask_label_hist_kernel<<<dim3((unsigned)gx, (unsigned)k), 256, (size_t)n_labels * 4, st>>>(masks, words, n_vox, label, n_labels,
                                                                                              counts);
    B2M_CHECK_LAUNCH();
  }
  hist_argmax_kernel<<<cdiv(k, 128), 128, 0, st>>>(counts, k, n_labels, best);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

/* counts int32[n_seg, n_labels] (zeroed inside) = label histogram per segment, best int32[n_seg] = its mode (lowest
 * label on ties, like torch.mode) - the per-segment vote of the S3DIS branch, models/detection_net.py:398-410. */
extern "C" int b2m_segment_label_vote(const int64_t* seg, const int32_t* label, int64_t n, int64_t n_seg, int32_t n_labels,
                                      int32_t* counts, int32_t* best, b2m_stream_t stream) {
  if (!seg || !label || !counts || !best || n < 0 || n_seg < 0 || n_labels <= 0) return B2M_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return B2M_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, (size_t)n_seg * n_labels * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (n > 0) {
    segment_label_hist_kernel<<<cdiv(n, 256), 256, 0, st>>>(seg, label, n, n_seg, n_labels, counts);
    B2M_CHECK_LAUNCH();
  }
  hist_argmax_kernel<<<cdiv(n_seg, 128), 128, 0, st>>>(counts, n_seg, n_labels, best);
  B2M_CHECK_LAUNCH();
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
