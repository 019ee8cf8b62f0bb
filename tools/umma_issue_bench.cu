// Micro-benchmark of the MMA issuer's per-stage latency chain (~0.4 us per pipeline stage
// whatever the stage holds). One CTA; warp 0 runs the loop the conv issuers run, with every barrier ALREADY
// complete and operands that are whatever shared memory holds, and reports SM cycles per iteration for each
// ingredient alone and for the combinations, with and without other warps spinning on mbarriers beside it:
//   0 empty loop                       1 try_wait on a completed mbarrier        2 two try_waits
//   3 tcgen05.fence::after_thread_sync 4 one tcgen05.commit                      5 two commits
//   6 nmma MMAs (M=128, N, K=16), no commit          7 nmma MMAs + one commit
//   8 the full stage: 2 try_waits + fence + nmma MMAs + 2 commits (what conv_fwd_kernel does per stage)
//   9 as 8, but the commits arrive on barriers that another warp waits for and re-arms (a real round trip)
// Run first thing next round:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I box2mask_b200/csrc \
//                                    -o build/umma_issue_bench tools/umma_issue_bench.cu && build/umma_issue_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

using namespace b2m;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kThreads = 608;          // the conv kernels' block size: 19 warps

__global__ void __launch_bounds__(kThreads, 1)
issue_bench(int mode, int iters, int ncols, int nmma, int spinners, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t done = smem_u32(&bars[0]);        // completed once before the loop: waits on parity 0 always pass
  const uint32_t sink_a = smem_u32(&bars[1]), sink_b = smem_u32(&bars[2]);   // commits arrive here, nobody waits
  const uint32_t rt_full = smem_u32(&bars[3]), rt_empty = smem_u32(&bars[4]);
  const uint32_t never = smem_u32(&bars[5]);       // never completes: what the spinning warps poll
  if (threadIdx.x == 0) {
    mbar_init(done, 1); mbar_init(sink_a, 1); mbar_init(sink_b, 1); mbar_init(rt_full, 1); mbar_init(rt_empty, 1);
    mbar_init(never, 1);
    mbar_fence_init();
    stop = 0;
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(done);
  __syncthreads();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 0) {
    const bool lead = elect_one_sync();
    const uint32_t idesc = umma_idesc_bf16(128, ncols, 0, 0);
    const uint32_t hi = umma_desc_hi(1024, 128);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16), b_lo = umma_desc_lo(smem_u32(smem) + 65536, 16);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 1 || mode == 2 || mode >= 8) mbar_wait(done, 0);
      if (mode == 2 || mode >= 8) mbar_wait(done, 0);
      if (mode == 9) mbar_wait(rt_full, phase);        // armed by the partner warp once the previous commit arrived
      if (mode == 3 || mode >= 8) tc_fence_after();
      if (lead) {
        if (mode >= 6)
          for (int k = 0; k < nmma; ++k)
            umma_bf16_lohi(tmem_base, a_lo + 2 * (k & 3) + 1024 * (k >> 2), hi, b_lo + 2 * (k & 3), hi, idesc, (it | k) ? 1u : 0u);
        if (mode == 4 || mode == 5 || mode == 7 || mode == 8) umma_commit(sink_a);
        if (mode == 5 || mode == 8) umma_commit(sink_b);
        if (mode == 9) { umma_commit(rt_empty); umma_commit(sink_b); }
      }
      phase ^= (mode == 9) ? 1u : 0u;
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) { out[0] = (unsigned long long)(t1 - t0); stop = 1; }
    if (lead && mode != 9) { }   // outstanding commits simply arrive on the sink barriers
  } else if (warp == 2 && mode == 9) {
    // partner of the round trip: waits for the issuer's commit, re-arms the "full" barrier (what a producer does)
    uint32_t phase = 0;
    if (lane == 0) mbar_arrive(rt_full);               // stage 0 is available at once
    for (int it = 0; it < iters; ++it) {
      mbar_wait(rt_empty, phase);
      if (lane == 0 && it + 1 < iters) mbar_arrive(rt_full);
      phase ^= 1u;
    }
  } else if (warp >= 3 && warp < 3 + spinners) {
    // bystanders: poll a barrier that never completes, the way idle producers / epilogue warps do
    while (!stop) {
      for (int i = 0; i < 64; ++i) (void)mbar_try_wait(never, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, 8));
  CK(cudaFuncSetAttribute(issue_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 20000;
  const char* names[] = {"empty loop", "1 try_wait (hit)", "2 try_waits (hit)", "tcgen05.fence::after", "1 commit", "2 commits",
                         "MMAs only", "MMAs + 1 commit", "full stage (2 waits, fence, MMAs, 2 commits)",
                         "full stage with a real commit->wait round trip"};
  for (int spinners = 0; spinners <= 16; spinners += 16) {
    printf("--- %d other warps polling mbarriers\n", spinners);
    for (int ncols = 32; ncols <= 256; ncols = (ncols == 32 ? 96 : (ncols == 96 ? 256 : 512))) {
      for (int nmma = 4; nmma <= 8; nmma += 4) {
        for (int mode = 0; mode <= 9; ++mode) {
          if (mode < 6 && !(ncols == 32 && nmma == 4)) continue;      // the MMA-free rows do not depend on N / count
          issue_bench<<<1, kThreads, 200 * 1024>>>(mode, iters, ncols, nmma, spinners, d);
          CK(cudaDeviceSynchronize());
          unsigned long long c = 0;
          CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
          printf("N=%3d nmma=%d  %-52s %8.1f cycles / iteration\n", ncols, nmma, names[mode], (double)c / iters);
        }
      }
    }
  }
  return 0;
}
