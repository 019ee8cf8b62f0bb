// Launch lists: one C call enqueues a whole pass of the U-Net trunk.
//
// The trunk of /root/reference/models/detection_net.py:235-337 is ~250 modules and ~540 kernel calls per training step;
// called one by one through a foreign-function interface the host needs ~50 us per call (argument marshalling, stream
// lookups, event objects for the side stream) for launches that take 10-25 us on the deep levels. The host side
// (box2mask_b200/trunk.py) therefore only *records* the calls of a pass - the arguments of the b2m_* entry points below,
// in the same order - and b2m_run_commands issues them back to back: ~5 us of launch cost per kernel instead of ~50.
// Two streams (weight gradients run beside the dgrad chain) with record / wait commands between them; the events come
// from a per-device pool owned by the library.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b2m {

// dst[r, 0:w] = src[r, 0:w] for bf16 rows with different pitches (channel concatenation and its backward split);
// one 16-byte vector per thread, consecutive threads along the row
__global__ void __launch_bounds__(256)
copy_columns_kernel(const uint4* __restrict__ src, int64_t src_ld16, uint4* __restrict__ dst, int64_t dst_ld16, int64_t n,
                    int w16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * w16) return;
  const int64_t r = i / w16;
  const int c = (int)(i - r * w16);
  dst[r * dst_ld16 + c] = __ldg(src + r * src_ld16 + c);
}

struct EventPool {
  std::mutex mu;
  std::vector<std::vector<cudaEvent_t>> per_device;
  cudaEvent_t get(int64_t slot) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || slot < 0 || slot >= (1 << 20)) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if ((int)per_device.size() <= dev) per_device.resize(dev + 1);
    auto& ev = per_device[dev];
    while ((int64_t)ev.size() <= slot) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      ev.push_back(e);
    }
    return ev[slot];
  }
};
static EventPool g_events;

}  // namespace b2m

using namespace b2m;

extern "C" int b2m_copy_columns(const uint16_t* src, int64_t src_ld, uint16_t* dst, int64_t dst_ld, int64_t n, int32_t width,
                                b2m_stream_t stream) {
  if (n == 0 || width == 0) return B2M_OK;
  if (!src || !dst || n < 0 || width < 0 || src_ld < width || dst_ld < width) return B2M_ERR_INVALID_ARGUMENT;
  if ((width % 8) || (src_ld % 8) || (dst_ld % 8) || (reinterpret_cast<uintptr_t>(src) & 15) ||
      (reinterpret_cast<uintptr_t>(dst) & 15))
    return B2M_ERR_UNSUPPORTED_SHAPE;
  const int w16 = width / 8;
  copy_columns_kernel<<<cdiv(n * w16, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), src_ld / 8,
                                                                         reinterpret_cast<uint4*>(dst), dst_ld / 8, n, w16);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_run_commands(const b2m_command_t* cmds, int64_t n, b2m_stream_t main_stream, b2m_stream_t side_stream,
                                int64_t* failed) {
  if (failed) *failed = -1;
  if (n == 0) return B2M_OK;
  if (!cmds || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  auto P = [](int64_t v) { return reinterpret_cast<void*>(static_cast<uintptr_t>(v)); };
  for (int64_t i = 0; i < n; ++i) {
    const b2m_command_t& c = cmds[i];
    const int64_t* a = c.a;
    b2m_stream_t st = c.stream ? side_stream : main_stream;
    int rc = B2M_OK;
    if (c.stream && !side_stream) rc = B2M_ERR_INVALID_ARGUMENT;
    else switch (c.op) {
      case B2M_CMD_CONV_FORWARD:
        rc = b2m_conv_dgrad_bn_reduce((const uint16_t*)P(a[0]), a[1], (int32_t)a[2], (const int32_t*)P(a[3]), (const int32_t*)P(a[4]),
                                 (const uint32_t*)P(a[5]), (int32_t)a[6], a[7], (const uint16_t*)P(a[8]), (int32_t)a[9],
                                 (uint16_t*)P(a[10]), (double*)P(a[11]), (const float*)P(a[12]), (const float*)P(a[13]),
                                 (const uint16_t*)P(a[14]), (int32_t)a[15], (float*)P(a[16]), (int32_t)a[17], P(a[18]),
                                 (size_t)a[19], (const uint16_t*)P(a[20]), (const uint8_t*)P(a[21]), (const float*)P(a[22]),
                                 (const float*)P(a[23]), (double*)P(a[24]), st);
        break;
      case B2M_CMD_CONV_WGRAD:
        rc = b2m_conv_wgrad_ex((const uint16_t*)P(a[0]), a[1], (int32_t)a[2], (const uint16_t*)P(a[3]), (int32_t)a[4],
                               (const int32_t*)P(a[5]), (const int32_t*)P(a[6]), (const uint32_t*)P(a[7]), (int32_t)a[8], a[9],
                               (float*)P(a[10]), P(a[11]), (size_t)a[12], st);
        break;
      case B2M_CMD_BN_FORWARD:
        rc = b2m_bn_forward((const uint16_t*)P(a[0]), a[1], a[2], (int32_t)a[3], (const double*)P(a[4]), (const float*)P(a[5]),
                            (const float*)P(a[6]), (float*)P(a[7]), (float*)P(a[8]), (float)c.f[0], (float)c.f[1], (int32_t)a[9],
                            (const uint16_t*)P(a[10]), (int32_t)a[11], (uint16_t*)P(a[12]), (float*)P(a[13]), (float*)P(a[14]),
                            (uint8_t*)P(a[15]), st);
        break;
      case B2M_CMD_BN_BACKWARD_REDUCE:
        rc = b2m_bn_backward_reduce((const uint16_t*)P(a[0]), (const uint16_t*)P(a[1]), (const uint16_t*)P(a[2]), a[3],
                                    (int32_t)a[4], (const float*)P(a[5]), (const float*)P(a[6]), (int32_t)a[7], (double*)P(a[8]),
                                    (const uint8_t*)P(a[9]), st);
        break;
      case B2M_CMD_BN_BACKWARD_APPLY:
        rc = b2m_bn_backward_apply((const uint16_t*)P(a[0]), (const uint16_t*)P(a[1]), (const uint16_t*)P(a[2]), a[3], a[4],
                                   (int32_t)a[5], (const float*)P(a[6]), (const float*)P(a[7]), (const float*)P(a[8]),
                                   (const double*)P(a[9]), (const double*)P(a[10]), (const double*)P(a[11]), (int32_t)a[12],
                                   (int32_t)a[13], (uint16_t*)P(a[14]), (uint16_t*)P(a[15]), (float*)P(a[16]), (float*)P(a[17]),
                                   (const uint8_t*)P(a[18]), st);
        break;
      case B2M_CMD_COPY_COLUMNS:
        rc = b2m_copy_columns((const uint16_t*)P(a[0]), a[1], (uint16_t*)P(a[2]), a[3], a[4], (int32_t)a[5], st);
        break;
      case B2M_CMD_RECORD: {
        cudaEvent_t e = g_events.get(a[0]);
        rc = (e && cudaEventRecord(e, (cudaStream_t)st) == cudaSuccess) ? B2M_OK : B2M_ERR_CUDA_LAUNCH;
        break;
      }
      case B2M_CMD_WAIT: {
        cudaEvent_t e = g_events.get(a[0]);
        rc = (e && cudaStreamWaitEvent((cudaStream_t)st, e, 0) == cudaSuccess) ? B2M_OK : B2M_ERR_CUDA_LAUNCH;
        break;
      }
      case B2M_CMD_PEER_ALLREDUCE:
        rc = b2m_peer_allreduce_f64((const double*)P(a[0]), (int32_t)a[1], c.f[0], (int32_t)a[2], (double*)P(a[3]),
                                    (void* const*)P(a[4]), (int32_t)a[5], (int32_t)a[6], (uint64_t)a[7], (int32_t*)P(a[8]), st);
        break;
      default:
        rc = B2M_ERR_INVALID_ARGUMENT;
    }
    if (rc != B2M_OK) {
      if (failed) *failed = i;
      return rc;
    }
  }
  return B2M_OK;
}
