// Micro-benchmark: throughput of the TMA row gather (tile::gather4) per SM as a function of the number of issuing
// warps, the box width and the table size (L2-resident vs HBM), next to an LDGSTS (cp.async 16 B) gather of the
// same rows. One persistent CTA per SM; each warp owns a private ring of shared-memory slots so nothing but the
// copy engines is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_gather_bench tools/tma_gather_bench.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++spins > (1u << 24)) __trap();
  }
}

constexpr int kSlots = 2;   // ring depth per warp; a slot = 128 rows x box bytes

// mode 0: every lane issues one gather4 (32 x 4 rows = 128 rows per stage)
// mode 1: lane 0 issues all 32 gather4 of the stage (indices broadcast by shuffle)
__global__ void __launch_bounds__(512, 1)
gather_bench(const __grid_constant__ CUtensorMap tm, const int32_t* __restrict__ idx, int64_t n_idx, int iters, int box_bytes,
             int mode, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[16 * kSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const uint32_t slot_bytes = 128 * box_bytes;
  if (lane == 0)
    for (int s = 0; s < kSlots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp * kSlots + s])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  const int64_t base = ((int64_t)blockIdx.x * nwarps + warp) * iters * 128;
  for (int it = 0; it < iters; ++it) {
    const int s = it % kSlots;
    const uint32_t bar = smem_u32(&bars[warp * kSlots + s]);
    const uint32_t dst = smem_u32(smem) + (warp * kSlots + s) * slot_bytes;
    const int4 r = *reinterpret_cast<const int4*>(idx + (base + (int64_t)it * 128 + 4 * lane) % n_idx);
    if (it >= kSlots) mbar_wait(bar, ((it / kSlots) - 1) & 1);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slot_bytes) : "memory");
    __syncwarp();
    if (mode == 0) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(dst + lane * 4 * box_bytes), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar), "r"(0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
    } else {
      for (int l = 0; l < 32; ++l) {
        const int r0 = __shfl_sync(0xffffffffu, r.x, l), r1 = __shfl_sync(0xffffffffu, r.y, l);
        const int r2 = __shfl_sync(0xffffffffu, r.z, l), r3 = __shfl_sync(0xffffffffu, r.w, l);
        if (lane == 0)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(dst + l * 4 * box_bytes), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
      }
    }
  }
  for (int s = 0; s < kSlots && s < iters; ++s) {
    const int last = ((iters - 1 - s) / kSlots) * kSlots + s;   // last iteration that used slot s ... any it == s mod kSlots
    if (last >= 0) mbar_wait(smem_u32(&bars[warp * kSlots + s]), (last / kSlots) & 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

// LDGSTS reference: the same rows with cp.async 16 B per lane (box_bytes / 16 lanes per row)
__global__ void __launch_bounds__(512, 1)
ldgsts_bench(const uint16_t* __restrict__ x, int cols, const int32_t* __restrict__ idx, int64_t n_idx, int iters, int box_bytes,
             unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const uint32_t slot_bytes = 128 * box_bytes;
  const int lpr = box_bytes / 16;        // lanes per row
  const int rpi = 32 / lpr;              // rows per instruction
  const int sub = lane % lpr, rb = lane / lpr;
  __syncthreads();
  const long long t0 = clock64();
  const int64_t base = ((int64_t)blockIdx.x * nwarps + warp) * iters * 128;
  for (int it = 0; it < iters; ++it) {
    const int s = it % kSlots;
    const uint32_t dst = smem_u32(smem) + (warp * kSlots + s) * slot_bytes;
    const int4 mine = *reinterpret_cast<const int4*>(idx + (base + (int64_t)it * 128 + 4 * lane) % n_idx);
    if (it >= kSlots) asm volatile("cp.async.wait_group %0;" ::"n"(kSlots - 1) : "memory");
    // rows of this lane's instruction j: row = j * rpi + rb; index broadcast by one shuffle, zero-fill form of cp.async
    const int vals[4] = {mine.x, mine.y, mine.z, mine.w};
#pragma unroll 8
    for (int j = 0; j < 128 / rpi; ++j) {
      const int row = j * rpi + rb;
      int r = __shfl_sync(0xffffffffu, vals[0], row >> 2);
      const int r1 = __shfl_sync(0xffffffffu, vals[1], row >> 2), r2 = __shfl_sync(0xffffffffu, vals[2], row >> 2),
                r3 = __shfl_sync(0xffffffffu, vals[3], row >> 2);
      r = (row & 3) == 0 ? r : ((row & 3) == 1 ? r1 : ((row & 3) == 2 ? r2 : r3));
      const uint32_t d = dst + row * box_bytes + ((sub ^ (row & 7)) << 4);
      const uint32_t nbytes = r >= 0 ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(x + (int64_t)max(r, 0) * cols + sub * 8), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

int main() {
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const int cols = 96;
  unsigned long long* dcyc;
  CK(cudaMalloc(&dcyc, sms * 8));
  CK(cudaFuncSetAttribute(gather_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  CK(cudaFuncSetAttribute(ldgsts_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  for (int64_t rows : {(int64_t)150000, (int64_t)2500000}) {   // 29 MB (L2 resident) and 480 MB (HBM)
    uint16_t* dx;
    CK(cudaMalloc(&dx, rows * cols * 2));
    CK(cudaMemset(dx, 1, rows * cols * 2));
    // index stream: locality like a sorted voxel map (neighbours within +-2000 rows of a slowly advancing cursor)
    const int64_t n_idx = 1 << 24;
    std::vector<int32_t> hidx(n_idx);
    uint64_t st = 12345;
    for (int64_t i = 0; i < n_idx; ++i) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      const int64_t cursor = (i / 16) % rows;
      int64_t r = cursor + (int64_t)((st >> 33) % 4001) - 2000;
      if (r < 0) r += rows;
      if (r >= rows) r -= rows;
      hidx[i] = (int32_t)r;
    }
    int32_t* didx;
    CK(cudaMalloc(&didx, n_idx * 4));
    CK(cudaMemcpy(didx, hidx.data(), n_idx * 4, cudaMemcpyHostToDevice));
    printf("table %lld rows x %d ch (%.0f MB)\n", (long long)rows, cols, rows * cols * 2 / 1e6);
    for (int box_bytes : {128, 32}) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
      cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
      cuuint32_t box[2] = {(cuuint32_t)(box_bytes / 2), 1};
      cuuint32_t estr[2] = {1, 1};
      const CUtensorMapSwizzle sw = box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      for (int mode : {0}) {
        for (int nwarps : {2, 4, 6, 8, 12}) {
          if (nwarps * kSlots * 128 * box_bytes > 200 * 1024) continue;
          const int iters = 400;
          const size_t smem = (size_t)nwarps * kSlots * 128 * box_bytes + 1024;
          gather_bench<<<sms, nwarps * 32, smem>>>(tm, didx, n_idx, iters, box_bytes, mode, dcyc);
          CK(cudaDeviceSynchronize());
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          gather_bench<<<sms, nwarps * 32, smem>>>(tm, didx, n_idx, iters, box_bytes, mode, dcyc);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          std::vector<unsigned long long> hc(sms);
          CK(cudaMemcpy(hc.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost));
          double avg = 0;
          for (auto c : hc) avg += (double)c / sms;
          const double bytes_sm = (double)nwarps * iters * 128 * box_bytes;
          printf("  gather4 box %3d B mode %d warps %2d: %.3f ms  %7.1f cyc per gather4 per SM  %6.1f B/clk/SM  %6.2f TB/s chip\n", box_bytes, mode,
                 nwarps, ms, avg / ((double)nwarps * iters * 32), bytes_sm / avg, bytes_sm * sms / (ms * 1e-3) / 1e12);
        }
      }
      for (int nwarps : {4, 6, 8, 12}) {
        if (nwarps * kSlots * 128 * box_bytes > 200 * 1024) continue;
        const int iters = 400;
        const size_t smem = (size_t)nwarps * kSlots * 128 * box_bytes + 1024;
        ldgsts_bench<<<sms, nwarps * 32, smem>>>(dx, cols, didx, n_idx, iters, box_bytes, dcyc);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        ldgsts_bench<<<sms, nwarps * 32, smem>>>(dx, cols, didx, n_idx, iters, box_bytes, dcyc);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes_sm = (double)nwarps * iters * 128 * box_bytes;
        printf("  ldgsts  box %3d B        warps %2d: %.3f ms  %6.2f TB/s chip\n", box_bytes, nwarps, ms, bytes_sm * sms / (ms * 1e-3) / 1e12);
      }
    }
    CK(cudaFree(dx));
    CK(cudaFree(didx));
  }
  return 0;
}
