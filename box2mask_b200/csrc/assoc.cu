// Weak-supervision label association on the GPU.
//
// Replaces the per-point / per-superpoint Python loops of /root/reference/models/dataloader.py:236-314
// (ScanNet.approx_association; the ARKitScenes and S3DIS variants at :539-621 and :805-927 use the same three steps):
//   1. point-in-box occupancy of every point against every (foreground) instance box: how many boxes contain the point,
//      the lowest such box and the one of least volume (float64 comparisons and volumes, like numpy);
//   2. per-point instance ids (`point_association`), or
//   3. per-superpoint decisions: from the least-covered point of the superpoint (default) or by majority vote, pooled back
//      to the points.
// Everything is integer / comparison work on a few hundred thousand points and <= a few hundred boxes: HBM / latency
// bound, one thread per point, boxes staged in shared memory.
#include "common.cuh"

namespace b2m {

constexpr int kAssocThreads = 256;
constexpr int kBoxChunk = 512;          // boxes staged per pass: 7 doubles each = 28 KB of shared memory

__global__ void __launch_bounds__(kAssocThreads)
point_box_kernel(const double* __restrict__ pos, int64_t n, const double* __restrict__ bmin, const double* __restrict__ bmax,
                 const double* __restrict__ vol, int nb, int32_t* __restrict__ num, int32_t* __restrict__ first,
                 int32_t* __restrict__ smallest) {
  __shared__ double sb[kBoxChunk * 7];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double x = 0, y = 0, z = 0;
  if (i < n) { x = pos[3 * i]; y = pos[3 * i + 1]; z = pos[3 * i + 2]; }
  int cnt = 0, f = -1, sm = -1;
  double best = 0.0;
  for (int b0 = 0; b0 < nb; b0 += kBoxChunk) {
    const int m = min(kBoxChunk, nb - b0);
    __syncthreads();
    for (int t = threadIdx.x; t < m; t += blockDim.x) {
      sb[7 * t + 0] = bmin[3 * (b0 + t)]; sb[7 * t + 1] = bmin[3 * (b0 + t) + 1]; sb[7 * t + 2] = bmin[3 * (b0 + t) + 2];
      sb[7 * t + 3] = bmax[3 * (b0 + t)]; sb[7 * t + 4] = bmax[3 * (b0 + t) + 1]; sb[7 * t + 5] = bmax[3 * (b0 + t) + 2];
      sb[7 * t + 6] = vol[b0 + t];
    }
    __syncthreads();
    if (i < n) {
      for (int t = 0; t < m; ++t) {
        const double* q = sb + 7 * t;
        const bool in = x >= q[0] && y >= q[1] && z >= q[2] && x <= q[3] && y <= q[4] && z <= q[5];
        if (in) {
          if (cnt == 0) f = b0 + t;
          if (cnt == 0 || q[6] < best) { best = q[6]; sm = b0 + t; }     // strict <: ties keep the lowest box index
          ++cnt;
        }
      }
    }
  }
  if (i < n) { num[i] = cnt; first[i] = f; smallest[i] = sm; }
}

// dataloader.py:243-259: -1 no box, the single box's instance, several boxes: -2 or the smallest box's instance
__global__ void point_instance_kernel(const int32_t* __restrict__ num, const int32_t* __restrict__ first,
                                      const int32_t* __restrict__ smallest, const int64_t* __restrict__ ids, int64_t n,
                                      int heuristic, int64_t* __restrict__ inst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = num[i];
  inst[i] = c == 0 ? -1 : (c == 1 ? ids[first[i]] : (heuristic ? ids[smallest[i]] : -2));
}

// default mode, dataloader.py:278-312: per superpoint the (number of boxes, point index) pair of its least-covered point,
// lowest point index on ties = the reference's np.where(...)[0][0] / argmin
__global__ void seg_min_kernel(const int32_t* __restrict__ num, const int32_t* __restrict__ seg_rank, int64_t n,
                               unsigned long long* __restrict__ key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg_rank[i];
  if (s >= 0) atomicMin(key + s, ((unsigned long long)(unsigned)num[i] << 32) | (unsigned long long)(unsigned)i);
}
__global__ void seg_decide_kernel(const unsigned long long* __restrict__ key, const int32_t* __restrict__ first,
                                  const int32_t* __restrict__ smallest, const int64_t* __restrict__ ids, int64_t n_segs,
                                  int heuristic, int64_t* __restrict__ per_seg) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  const unsigned long long k = key[s];
  int64_t v = -2;                                   // superpoint without points: keeps the default
  if (k != ~0ull) {
    const unsigned mn = (unsigned)(k >> 32), p = (unsigned)(k & 0xFFFFFFFFull);
    if (mn == 1) v = ids[first[p]];
    else if (mn == 0) v = -1;
    else if (heuristic) v = ids[smallest[p]];
  }
  per_seg[s] = v;
}

// majority vote, dataloader.py:264-274: histogram of value codes per superpoint, then the most common code, the lowest code
// on ties (codes are ordered like the values: -2, -1, then the instance ids ascending = scipy.stats.mode's tie break)
__global__ void seg_hist_kernel(const int32_t* __restrict__ num, const int32_t* __restrict__ first,
                                const int32_t* __restrict__ smallest, const int32_t* __restrict__ id_rank,
                                const int32_t* __restrict__ seg_rank, int64_t n, int heuristic, int ncodes,
                                int32_t* __restrict__ hist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg_rank[i];
  if (s < 0) return;
  const int c = num[i];
  const int code = c == 0 ? 1 : (c == 1 ? 2 + id_rank[first[i]] : (heuristic ? 2 + id_rank[smallest[i]] : 0));
  atomicAdd(hist + (int64_t)s * ncodes + code, 1);
}
__global__ void seg_mode_kernel(const int32_t* __restrict__ hist, const int64_t* __restrict__ sorted_ids, int64_t n_segs,
                                int ncodes, int64_t* __restrict__ per_seg) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  int best = -1, bc = 0;
  for (int c = 0; c < ncodes; ++c) {
    const int h = hist[s * ncodes + c];
    if (h > bc) { bc = h; best = c; }
  }
  per_seg[s] = best < 0 ? -2 : (best == 0 ? -2 : (best == 1 ? -1 : sorted_ids[best - 2]));
}

__global__ void pool_to_points_kernel(const int64_t* __restrict__ per_seg, const int32_t* __restrict__ seg_rank, int64_t n,
                                      int64_t* __restrict__ per_point) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg_rank[i];
  per_point[i] = s >= 0 ? per_seg[s] : -2;          // points of superpoints lost in voxelisation keep "unknown"
}

}  // namespace b2m

using namespace b2m;

extern "C" int b2m_point_box_occupancy(const double* positions, int64_t n, const double* box_min, const double* box_max,
                                       const double* volume, int32_t n_boxes, int32_t* num, int32_t* first,
                                       int32_t* smallest, b2m_stream_t stream) {
  if (n == 0) return B2M_OK;
  if (!positions || !num || !first || !smallest || n < 0 || n_boxes < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (n_boxes > 0 && (!box_min || !box_max || !volume)) return B2M_ERR_INVALID_ARGUMENT;
  point_box_kernel<<<cdiv(n, kAssocThreads), kAssocThreads, 0, (cudaStream_t)stream>>>(positions, n, box_min, box_max, volume,
                                                                                     n_boxes, num, first, smallest);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_point_instances(const int32_t* num, const int32_t* first, const int32_t* smallest,
                                   const int64_t* instance_ids, int64_t n, int32_t smallest_heuristic, int64_t* inst,
                                   b2m_stream_t stream) {
  if (n == 0) return B2M_OK;
  if (!num || !first || !smallest || !inst || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  point_instance_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(num, first, smallest, instance_ids, n,
                                                                      smallest_heuristic ? 1 : 0, inst);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" size_t b2m_segment_association_workspace_bytes(int64_t n_segs, int32_t n_boxes, int32_t majority_vote) {
  if (n_segs <= 0) return 16;
  return majority_vote ? (size_t)n_segs * (size_t)(n_boxes + 2) * 4 : (size_t)n_segs * 8;
}

extern "C" int b2m_segment_association(const int32_t* num, const int32_t* first, const int32_t* smallest,
                                       const int32_t* seg_rank, int64_t n, int64_t n_segs, const int64_t* instance_ids,
                                       const int64_t* sorted_ids, const int32_t* id_rank, int32_t n_boxes,
                                       int32_t majority_vote, int32_t smallest_heuristic, int64_t* per_seg,
                                       int64_t* per_point, void* workspace, size_t workspace_bytes, b2m_stream_t stream) {
  if (!num || !first || !smallest || !seg_rank || !per_seg || !per_point || !workspace || n < 0 || n_segs < 0 || n_boxes < 0)
    return B2M_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < b2m_segment_association_workspace_bytes(n_segs, n_boxes, majority_vote)) return B2M_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  const int heur = smallest_heuristic ? 1 : 0;
  if (n_segs > 0) {
    if (majority_vote) {
      if (n_boxes > 0 && (!sorted_ids || !id_rank)) return B2M_ERR_INVALID_ARGUMENT;
      const int ncodes = n_boxes + 2;
      int32_t* hist = reinterpret_cast<int32_t*>(workspace);
      if (cudaMemsetAsync(hist, 0, (size_t)n_segs * ncodes * 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
      if (n > 0) {
        seg_hist_kernel<<<cdiv(n, 256), 256, 0, st>>>(num, first, smallest, id_rank, seg_rank, n, heur, ncodes, hist);
        B2M_CHECK_LAUNCH();
      }
      seg_mode_kernel<<<cdiv(n_segs, 256), 256, 0, st>>>(hist, sorted_ids, n_segs, ncodes, per_seg);
      B2M_CHECK_LAUNCH();
    } else {
      auto* key = reinterpret_cast<unsigned long long*>(workspace);
      if (cudaMemsetAsync(key, 0xFF, (size_t)n_segs * 8, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
      if (n > 0) {
        seg_min_kernel<<<cdiv(n, 256), 256, 0, st>>>(num, seg_rank, n, key);
        B2M_CHECK_LAUNCH();
      }
      seg_decide_kernel<<<cdiv(n_segs, 256), 256, 0, st>>>(key, first, smallest, instance_ids, n_segs, heur, per_seg);
      B2M_CHECK_LAUNCH();
    }
  }
  if (n > 0) {
    pool_to_points_kernel<<<cdiv(n, 256), 256, 0, st>>>(per_seg, seg_rank, n, per_point);
    B2M_CHECK_LAUNCH();
  }
  return B2M_OK;
}
