// Probe of the sm_100a TMA row-gather / row-scatter modes (cp.async.bulk.tensor.2d ... tile::gather4 / tile::scatter4)
// used by the sparse-convolution kernels: which box shape the tensor map needs, what lands where in shared memory
// (128B swizzle), how many bytes complete_tx counts, and what happens for row indices outside the tensor
// (negative / >= rows) and for column boxes that overhang the channel count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tma_probe tools/tma_probe.cu && gpurun_out/tma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void gather_probe(const __grid_constant__ CUtensorMap tm, int col, int r0, int r1, int r2, int r3, uint32_t tx,
                             uint16_t* out, int* status) {
  __shared__ __align__(1024) uint16_t buf[4096];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 0xFFFF;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  int ok = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(tx) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
    for (int spin = 0; spin < (1 << 18) && !ok; ++spin) {
      uint32_t p;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(p) : "r"(b) : "memory");
      ok = (int)p;
    }
    status[0] = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = buf[i];
}

__global__ void scatter_probe(const __grid_constant__ CUtensorMap tm, int col, int r0, int r1, int r2, int r3) {
  __shared__ __align__(1024) uint16_t buf[4096];
  // dense (unswizzled view): element e of row j = 1000*(j+1) + e, written at the 128B-swizzled position
  for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x) {
    const int j = i / 64, e = i % 64;
    const int chunk = e / 8, within = e % 8;
    buf[j * 64 + ((chunk ^ (j & 7)) * 8) + within] = (uint16_t)(1000 * (j + 1) + e);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(buf)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int make_map(CUtensorMap* tm, void* base, int rows, int cols, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

static void describe(const std::vector<uint16_t>& h, int cols) {
  // print, for each 128-byte line of the first 8 lines, what each 16-byte chunk holds (row:col of its first element)
  for (int line = 0; line < 8; ++line) {
    printf("   line %d:", line);
    for (int ch = 0; ch < 8; ++ch) {
      const uint16_t v = h[line * 64 + ch * 8];
      if (v == 0xFFFF) printf("  ----- ");
      else if (v == 0) printf("  zero  ");
      else printf(" %3d:%-3d", v / cols, v % cols);
    }
    printf("\n");
  }
}

int main() {
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  const int rows = 600, cols = 96;
  std::vector<uint16_t> hx((size_t)rows * cols);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) hx[(size_t)r * cols + c] = (uint16_t)(r * cols + c);
  hx[0] = 1;  // keep (0,0) distinguishable from a zero fill
  uint16_t *dx, *dout, *dy;
  int* dstatus;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMalloc(&dy, hx.size() * 2));
  CK(cudaMalloc(&dout, 4096 * 2));
  CK(cudaMalloc(&dstatus, 16));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  std::vector<uint16_t> h(4096);

  struct Case { const char* name; int box_cols, box_rows, col, r[4]; uint32_t tx; CUtensorMapSwizzle sw; };
  const Case cases[] = {
      {"box{64,1} sw128 rows 5,17,100,599 tx512", 64, 1, 0, {5, 17, 100, 599}, 512, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box{64,1} sw128 rows 5,-1,100,600 (oob rows) tx512", 64, 1, 0, {5, -1, 100, 600}, 512, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box{64,1} sw128 col 64 (overhang, c=96) tx512", 64, 1, 64, {5, 17, 100, 599}, 512, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box{32,1} sw64 col 64 tx256", 32, 1, 64, {5, 17, 100, 599}, 256, CU_TENSOR_MAP_SWIZZLE_64B},
      {"box{64,1} sw128 all rows -1 tx512", 64, 1, 0, {-1, -1, -1, -1}, 512, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box{64,1} nosw rows 5,17,100,599 tx512", 64, 1, 0, {5, 17, 100, 599}, 512, CU_TENSOR_MAP_SWIZZLE_NONE},
  };
  for (const Case& c : cases) {
    CUtensorMap tm;
    const int er = make_map(&tm, dx, rows, cols, c.box_cols, c.box_rows, c.sw);
    printf("== %s: encode rc=%d\n", c.name, er);
    if (er != 0) continue;
    CK(cudaMemset(dstatus, 0, 16));
    gather_probe<<<1, 128>>>(tm, c.col, c.r[0], c.r[1], c.r[2], c.r[3], c.tx, dout, dstatus);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 2; }
    int st = 0;
    CK(cudaMemcpy(&st, dstatus, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h.data(), dout, 4096 * 2, cudaMemcpyDeviceToHost));
    printf("   barrier completed: %d\n", st);
    describe(h, cols);
  }
  // scatter4
  {
    CUtensorMap tm;
    const int er = make_map(&tm, dy, rows, cols, 64, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    printf("== scatter4 box{64,1} sw128 rows 7,-1,9,600 col 0: encode rc=%d\n", er);
    CK(cudaMemset(dy, 0, hx.size() * 2));
    scatter_probe<<<1, 128>>>(tm, 0, 7, -1, 9, 600);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<uint16_t> hy(hx.size());
    CK(cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost));
    for (int r = 0; r < rows; ++r) {
      int nz = 0;
      for (int cc = 0; cc < cols; ++cc) nz += hy[(size_t)r * cols + cc] != 0;
      if (nz) printf("   y row %d: %d non-zero, first %d, [63]=%d [64]=%d\n", r, nz, hy[(size_t)r * cols], hy[(size_t)r * cols + 63], hy[(size_t)r * cols + 64]);
    }
    // overhanging column box on the store side
    CK(cudaMemset(dy, 0, hx.size() * 2));
    scatter_probe<<<1, 128>>>(tm, 64, 3, 4, -1, 8);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost));
    printf("== scatter4 col 64 (overhang) rows 3,4,-1,8\n");
    for (int r = 0; r < rows; ++r) {
      int nz = 0;
      for (int cc = 0; cc < cols; ++cc) nz += hy[(size_t)r * cols + cc] != 0;
      if (nz) printf("   y row %d: %d non-zero, [63]=%d [64]=%d [95]=%d\n", r, nz, hy[(size_t)r * cols + 63], hy[(size_t)r * cols + 64], hy[(size_t)r * cols + 95]);
    }
  }
  printf("probe done\n");
  return 0;
}
