// Voxelisation of a point cloud on the GPU: the step right before the hot path.
//
// Replaces the per-scene CPU numpy / scikit-learn code of /root/reference/models/dataloader.py:61-77:
//   input_coords = (positions - min(0, min(positions))) / voxel_size           (float64)
//   vox_coords, vox2point = np.unique(np.round(input_coords), axis=0, return_inverse=True)
//   point2vox = 1-nearest scene point of every voxel centre (ball tree over input_coords)
// b2m_voxel_coords produces the rounded integer coordinate of every point; the sorted unique voxel set and the
// point -> voxel map then come from b2m_downsample_coords with stride 1 (radix sort + head flags + scan), and
// b2m_nearest_point finds the nearest point of every voxel centre EXACTLY: the voxel's own points are at most
// sqrt(3)/2 from its centre, and any point that close rounds into the 3x3x3 neighbourhood, so the minimum over the
// points of the 27 neighbouring voxels (the k = 3 kernel map of the voxel set) is the global minimum.
#include "common.cuh"

namespace b2m {

// coords[i] = (0, rint((p - shift) / voxel)) with shift = min(0, *min_pos): the reference's operation order in fp64;
// rint is round-half-to-even like np.round. status[0] counts points outside the packable coordinate range.
__global__ void voxel_coords_kernel(const double* __restrict__ pos, int64_t n, const double* __restrict__ min_pos,
                                    double voxel, int4* __restrict__ coords, int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double shift = fmin(0.0, *min_pos);
  const double x = rint((pos[3 * i] - shift) / voxel), y = rint((pos[3 * i + 1] - shift) / voxel),
               z = rint((pos[3 * i + 2] - shift) / voxel);
  const bool ok = x >= -32768.0 && x <= 32767.0 && y >= -32768.0 && y <= 32767.0 && z >= -32768.0 && z <= 32767.0;
  if (!ok) atomicAdd(status, 1);
  coords[i] = ok ? make_int4(0, (int)x, (int)y, (int)z) : make_int4(0, 0, 0, 0);
}

// One thread per voxel: scan the points of the 27 neighbouring voxels (CSR: point_order[start[v] .. start[v+1])),
// squared distance to the voxel centre in fp64, ties -> the lower point index.
__global__ void nearest_point_kernel(const double* __restrict__ pos, const double* __restrict__ min_pos, double voxel,
                                     const int4* __restrict__ vox_coords, int64_t n_vox, const int32_t* __restrict__ nbr,
                                     int64_t n_pitch, const int64_t* __restrict__ start,
                                     const int64_t* __restrict__ point_order, int64_t* __restrict__ nearest) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  const double shift = fmin(0.0, *min_pos);
  const int4 c = vox_coords[v];
  const double cx = (double)c.y, cy = (double)c.z, cz = (double)c.w;
  double best = 1e300;
  int64_t best_i = -1;
  for (int k = 0; k < 27; ++k) {
    const int32_t u = __ldg(nbr + (int64_t)k * n_pitch + v);
    if (u < 0) continue;
    const int64_t e = start[u + 1];
    for (int64_t j = start[u]; j < e; ++j) {
      const int64_t i = point_order[j];
      const double dx = (pos[3 * i] - shift) / voxel - cx, dy = (pos[3 * i + 1] - shift) / voxel - cy,
                   dz = (pos[3 * i + 2] - shift) / voxel - cz;
      const double d = dx * dx + dy * dy + dz * dz;
      if (d < best || (d == best && i < best_i)) { best = d; best_i = i; }
    }
  }
  nearest[v] = best_i;
}

}  // namespace b2m

using namespace b2m;

extern "C" int b2m_voxel_coords(const double* positions, int64_t n, const double* min_position, double voxel_size,
                                int32_t* coords, int32_t* status, b2m_stream_t stream) {
  if (n == 0) return B2M_OK;
  if (!positions || !min_position || !coords || !status || n < 0 || !(voxel_size > 0.0)) return B2M_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(status, 0, 4, st) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  voxel_coords_kernel<<<cdiv(n, 256), 256, 0, st>>>(positions, n, min_position, voxel_size, reinterpret_cast<int4*>(coords), status);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_nearest_point(const double* positions, const double* min_position, double voxel_size,
                                 const int32_t* vox_coords, int64_t n_vox, const int32_t* nbr, const int64_t* start,
                                 const int64_t* point_order, int64_t* nearest, b2m_stream_t stream) {
  if (n_vox == 0) return B2M_OK;
  if (!positions || !min_position || !vox_coords || !nbr || !start || !point_order || !nearest || n_vox < 0 ||
      !(voxel_size > 0.0))
    return B2M_ERR_INVALID_ARGUMENT;
  nearest_point_kernel<<<cdiv(n_vox, 128), 128, 0, (cudaStream_t)stream>>>(
      positions, min_position, voxel_size, reinterpret_cast<const int4*>(vox_coords), n_vox, nbr, b2m_map_pitch(n_vox),
      start, point_order, nearest);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
