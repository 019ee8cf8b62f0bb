// SyncBatchNorm statistics exchanged over NVLink peer memory (one node, one process per GPU).
//
// The reference converts every BatchNorm to SyncBatchNorm under cfg.multigpu (/root/reference/models/model.py:25): per
// layer one all-reduce of (sum x, sum x^2, rows) in forward and one of (sum g, sum g * xhat) in backward - 162 dependent
// collectives of <= 8 KB per training step, each a library launch with tens of microseconds of latency. Here every rank
// owns a small exchange buffer that its peers map through CUDA IPC; one single-CTA kernel writes the rank's vector into
// its slot of EVERY peer's buffer with plain stores over NVLink, publishes a sequence number, waits for the peers'
// sequence numbers in its own buffer and adds the slots in rank order (so all ranks get bit-identical sums). No host
// involvement, stream-ordered, and a command of the launch lists (runlist.cu) like any other kernel of the pass.
//
// Buffer layout: uint64 flags[kPeerMaxRanks] (flags[r] = sequence number of the last exchange rank r has written here),
// then double slots[2][kPeerMaxRanks][kPeerSlotDoubles]. Two slot sets alternate with the sequence number's parity: a rank
// can be at most one exchange ahead of a peer (it needs the peer's flag of exchange k to finish k), so while a fast rank
// writes exchange k + 1 into one set the slow one still reads exchange k from the other.
#include "common.cuh"

namespace b2m {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerSlotDoubles = 2048;
constexpr size_t kPeerFlagBytes = 256;    // >= kPeerMaxRanks * 8, keeps the slots 256-byte aligned
constexpr size_t kPeerBytes = kPeerFlagBytes + (size_t)2 * kPeerMaxRanks * kPeerSlotDoubles * sizeof(double);

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double* peer_slot(uint8_t* buf, int parity, int rank) {
  return reinterpret_cast<double*>(buf + kPeerFlagBytes) + ((size_t)parity * kPeerMaxRanks + rank) * kPeerSlotDoubles;
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const double* __restrict__ in, int n, double tail, int use_tail, double* __restrict__ out,
                      uint8_t* const* __restrict__ peers, int rank, int world, unsigned long long seq,
                      int* __restrict__ status) {
  __shared__ uint8_t* bufs[kPeerMaxRanks];
  if (threadIdx.x < world) bufs[threadIdx.x] = peers[threadIdx.x];
  __syncthreads();
  const int par = (int)(seq & 1ull);
  // 1. this rank's vector into slot `rank` of every buffer (its own included)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = (use_tail && i == n - 1) ? tail : in[i];
    for (int p = 0; p < world; ++p) peer_slot(bufs[p], par, rank)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, then wait until every rank has published this exchange here
  if (threadIdx.x < world) {
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(bufs[threadIdx.x]) + rank, seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(bufs[rank]) + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(mine) < seq) {
      if (clock64() - t0 > 6000000000ll) {      // ~3 s: a peer never arrived (crashed rank, mismatched call order)
        atomicExch(status, 1);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  // 3. sum of the slots in rank order, read past L1 (the lines were last read two exchanges ago)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += __ldcg(peer_slot(bufs[rank], par, r) + i);
    out[i] = acc;
  }
}

}  // namespace b2m

using namespace b2m;

extern "C" size_t b2m_peer_buffer_bytes(void) { return kPeerBytes; }
extern "C" int32_t b2m_peer_max_doubles(void) { return kPeerSlotDoubles; }

extern "C" int b2m_peer_buffer_create(void** buffer, void* ipc_handle) {
  if (!buffer || !ipc_handle) return B2M_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  if (cudaMalloc(&p, kPeerBytes) != cudaSuccess) return B2M_ERR_CUDA_LAUNCH;
  if (cudaMemset(p, 0, kPeerBytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle), p) != cudaSuccess) {
    cudaFree(p);
    cudaGetLastError();
    return B2M_ERR_CUDA_LAUNCH;
  }
  *buffer = p;
  return B2M_OK;
}

extern "C" int b2m_peer_buffer_open(const void* ipc_handle, void** buffer) {
  if (!buffer || !ipc_handle) return B2M_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return B2M_ERR_CUDA_LAUNCH;
  }
  *buffer = p;
  return B2M_OK;
}

extern "C" int b2m_peer_buffer_close(void* buffer, int32_t own) {
  if (!buffer) return B2M_OK;
  const cudaError_t e = own ? cudaFree(buffer) : cudaIpcCloseMemHandle(buffer);
  if (e != cudaSuccess) { cudaGetLastError(); return B2M_ERR_CUDA_LAUNCH; }
  return B2M_OK;
}

extern "C" int b2m_peer_allreduce_f64(const double* in, int32_t n, double tail, int32_t use_tail, double* out,
                                      void* const* peer_buffers, int32_t rank, int32_t world, uint64_t seq, int32_t* status,
                                      b2m_stream_t stream) {
  if (!in || !out || !peer_buffers || !status || n <= 0 || world < 1 || rank < 0 || rank >= world || seq == 0)
    return B2M_ERR_INVALID_ARGUMENT;
  if (world > kPeerMaxRanks || n > kPeerSlotDoubles) return B2M_ERR_UNSUPPORTED_SHAPE;
  peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(in, n, tail, use_tail ? 1 : 0, out,
                                                            reinterpret_cast<uint8_t* const*>(peer_buffers), rank, world,
                                                            (unsigned long long)seq, status);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
