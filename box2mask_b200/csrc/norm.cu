// BatchNorm over sparse-tensor rows (+ residual, + ReLU), column statistics, dtype conversion.
//
// Replaces MinkowskiBatchNorm (-> torch.nn.BatchNorm1d, /root/reference/models/resnet.py:63,66,159 and
// models/detection_net.py:40..187), MinkowskiReLU and `out += residual` (models/resnet.py:67,80-81).
// HBM-bound passes: 16-byte (8 x bf16) vector loads/stores, one read + one write per element, column
// statistics reduced in fp32 per thread, fp32 in shared memory per block, fp64 atomics across blocks.
#include "common.cuh"

namespace b2m {

constexpr int kNormThreads = 256;

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__global__ void cast_pad_kernel(const float* __restrict__ x, int64_t n, int c, int c_pad, uint16_t* __restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * c_pad) return;
  const int64_t r = gid / c_pad;
  const int j = (int)(gid % c_pad);
  const float v = (j < c) ? x[r * c + j] : 0.f;
  reinterpret_cast<__nv_bfloat16*>(out)[gid] = __float2bfloat16_rn(v);
}

// Generic two-quantity column reduction over rows [n, c] of 8-wide vectors.
// MODE 0: (sum x, sum x^2).  MODE 1: (sum g, sum g*xhat) with g = dout * (out > 0 if relu).
template <int MODE>
__global__ void __launch_bounds__(kNormThreads)
colreduce_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ out, const uint16_t* __restrict__ dout,
                 int64_t n, int c, const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                 const uint8_t* __restrict__ relu_mask, double* __restrict__ red) {
  extern __shared__ float sh[];  // [rows_per_pass][c][2]
  const int G = c / 8;
  const int rpp = kNormThreads / G;  // rows per pass
  const int g = threadIdx.x % G;
  const int rl = threadIdx.x / G;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
  float mu[8], is[8];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mu[i] = mean[g * 8 + i]; is[i] = invstd[g * 8 + i]; }
  }
  if (rl < rpp) {
    const int64_t rstep = (int64_t)gridDim.x * rpp;
    constexpr int U = (MODE == 0) ? 4 : 2;   // rows in flight per thread
    for (int64_t r0 = (int64_t)blockIdx.x * rpp + rl; r0 < n; r0 += U * rstep) {
      uint4 xr[U], gr[U], orr[U];
      unsigned mk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * rstep;
        mk[u] = 0xFFu;
        if (r < n) {
          xr[u] = __ldg(reinterpret_cast<const uint4*>(x + r * c) + g);
          if (MODE == 1) {
            gr[u] = __ldg(reinterpret_cast<const uint4*>(dout + r * c) + g);
            if (relu) {
              if (relu_mask) mk[u] = __ldg(relu_mask + r * G + g);      // one byte instead of 16: bit i = (out[8 g + i] > 0)
              else orr[u] = __ldg(reinterpret_cast<const uint4*>(out + r * c) + g);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * rstep;
        if (r >= n) continue;
        float xv[8];
        bf16x8_to_float(xr[u], xv);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i] += xv[i]; b[i] = fmaf(xv[i], xv[i], b[i]); }
        } else {
          float gv[8];
          bf16x8_to_float(gr[u], gv);
          if (relu && relu_mask) {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (!((mk[u] >> i) & 1u)) gv[i] = 0.f;
          } else if (relu) {
            float ov[8];
            bf16x8_to_float(orr[u], ov);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (!(ov[i] > 0.f)) gv[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i] += gv[i]; b[i] = fmaf(gv[i], (xv[i] - mu[i]) * is[i], b[i]); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sh[(rl * c + g * 8 + i) * 2] = a[i];
      sh[(rl * c + g * 8 + i) * 2 + 1] = b[i];
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < c; col += kNormThreads) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < rpp; ++r) { s1 += sh[(r * c + col) * 2]; s2 += sh[(r * c + col) * 2 + 1]; }
    atomicAdd(red + col, (double)s1);
    atomicAdd(red + c + col, (double)s2);
  }
}

// Apply kernels: thread (rl, g) owns the 8 columns of group g for rows rl, rl + rows-per-pass, ... so the per-column
// coefficients live in registers (loaded once) and a warp still reads consecutive 16-byte vectors (row-major rows
// are contiguous: thread t of a pass reads vector r0 * G + t).
// Statistics and apply in one launch: every thread derives mean / invstd of its 8 columns from the fp64 column sums
// (training) or the running statistics (eval); block 0 also writes the saved statistics and updates the running ones.
__global__ void __launch_bounds__(kNormThreads)
bn_apply_kernel(const uint16_t* __restrict__ x, int64_t n, int64_t n_stat, int c, const double* __restrict__ sums,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
                float momentum, float eps, int training, const uint16_t* __restrict__ residual, int relu,
                uint16_t* __restrict__ out, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                uint8_t* __restrict__ relu_mask) {
  const int G = c / 8;
  const int rpp = kNormThreads / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  if (rl >= rpp) return;
  // SyncBatchNorm: the global row count travels with the all-reduced sums (element 2c), no host round trip
  if (training && n_stat <= 0) n_stat = (int64_t)llrint(sums[2 * c]);
  float mu[8], sc[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = g * 8 + i;
    float mean, invstd;
    double var = 0.0;
    if (training) {
      const double m = sums[j] / (double)n_stat;
      var = sums[c + j] / (double)n_stat - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
    } else {
      mean = running_mean[j];
      invstd = (float)(1.0 / sqrt((double)running_var[j] + (double)eps));
    }
    mu[i] = mean;
    sc[i] = gamma[j] * invstd;
    be[i] = beta[j];
    if (blockIdx.x == 0 && rl == 0) {          // one writer per column
      save_mean[j] = mean;
      save_invstd[j] = invstd;
      if (training) {
        if (running_mean) running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * mean;
        if (running_var) {
          const double unbiased = (n_stat > 1) ? var * (double)n_stat / (double)(n_stat - 1) : var;
          running_var[j] = (1.f - momentum) * running_var[j] + momentum * (float)unbiased;
        }
      }
    }
  }
  constexpr int U = 4;  // independent rows (16-byte loads) in flight per thread
  const int64_t rstep = (int64_t)gridDim.x * rpp;
  for (int64_t r0 = (int64_t)blockIdx.x * rpp + rl; r0 < n; r0 += U * rstep) {
    uint4 xr[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * rstep;
      if (r < n) {
        xr[u] = __ldg(reinterpret_cast<const uint4*>(x + r * c) + g);
        if (residual) rr[u] = __ldg(reinterpret_cast<const uint4*>(residual + r * c) + g);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * rstep;
      if (r >= n) continue;
      float xv[8], o[8];
      bf16x8_to_float(xr[u], xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(xv[i] - mu[i], sc[i], be[i]);
      if (residual) {
        float rv[8];
        bf16x8_to_float(rr[u], rv);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += rv[i];
      }
      if (relu) {
        if (relu_mask) {             // the gate of the backward pass as one byte per 8 columns (it re-read `out` before)
          unsigned m = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) m |= (o[i] > 0.f ? 1u : 0u) << i;
          relu_mask[r * G + g] = (uint8_t)m;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
      }
      reinterpret_cast<uint4*>(out + r * c)[g] = float_to_bf16x8(o);
    }
  }
}

__global__ void __launch_bounds__(kNormThreads)
bn_bwd_apply_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ out, const uint16_t* __restrict__ dout,
                    int64_t n, int64_t n_stat, int c, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const double* __restrict__ red,
                    const double* __restrict__ red_local, const double* __restrict__ n_stat_dev, int relu, int training,
                    uint16_t* __restrict__ dx, uint16_t* __restrict__ dres, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, const uint8_t* __restrict__ relu_mask) {
  const int G = c / 8;
  // The affine gradients come from THIS rank's reduction (red_local): torch's SyncBatchNorm keeps grad_weight /
  // grad_bias local and lets the gradient all-reduce average them like every other parameter; only dx needs the
  // global (sum g, sum g*xhat) in `red`.
  if (blockIdx.x == 0) {
    const double* rl_ = red_local ? red_local : red;
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
      if (dbeta) dbeta[j] = (float)rl_[j];
      if (dgamma) dgamma[j] = (float)rl_[c + j];
    }
  }
  if (n_stat_dev) n_stat = (int64_t)llrint(*n_stat_dev);
  const int rpp = kNormThreads / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  if (rl >= rpp) return;
  // dx = sc * (g - sg - xhat * sgx), xhat = (x - mean) * invstd  ==  sc * (g - sg - (x - mean) * kx)
  const float inv_n = 1.f / (float)n_stat;
  float mu[8], sc[8], sg[8], kx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = g * 8 + i;
    const float is = invstd[j];
    mu[i] = mean[j];
    sc[i] = gamma[j] * is;
    if (training) {
      sg[i] = (float)red[j] * inv_n;
      kx[i] = is * ((float)red[c + j] * inv_n);
    } else {
      sg[i] = 0.f;
      kx[i] = 0.f;
    }
  }
  constexpr int U = 2;
  const int64_t rstep = (int64_t)gridDim.x * rpp;
  for (int64_t r0 = (int64_t)blockIdx.x * rpp + rl; r0 < n; r0 += U * rstep) {
    uint4 xr[U], gr[U], orr[U];
    unsigned mk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * rstep;
      mk[u] = 0xFFu;
      if (r < n) {
        xr[u] = __ldg(reinterpret_cast<const uint4*>(x + r * c) + g);
        gr[u] = __ldg(reinterpret_cast<const uint4*>(dout + r * c) + g);
        if (relu) {
          if (relu_mask) mk[u] = __ldg(relu_mask + r * G + g);
          else orr[u] = __ldg(reinterpret_cast<const uint4*>(out + r * c) + g);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * rstep;
      if (r >= n) continue;
      float xv[8], gv[8], o[8];
      bf16x8_to_float(xr[u], xv);
      bf16x8_to_float(gr[u], gv);
      if (relu && relu_mask) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (!((mk[u] >> i) & 1u)) gv[i] = 0.f;
      } else if (relu) {
        float ov[8];
        bf16x8_to_float(orr[u], ov);
#pragma unroll
        for (int i = 0; i < 8; ++i) if (!(ov[i] > 0.f)) gv[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = sc[i] * (gv[i] - sg[i] - (xv[i] - mu[i]) * kx[i]);
      reinterpret_cast<uint4*>(dx + r * c)[g] = float_to_bf16x8(o);
      if (dres) reinterpret_cast<uint4*>(dres + r * c)[g] = float_to_bf16x8(gv);
    }
  }
}

static int reduce_grid(int64_t n, int c) {
  const int rpp = kNormThreads / (c / 8);
  int64_t blocks = (n + rpp * 16 - 1) / (rpp * 16);  // >= 16 rows per thread lane
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int apply_grid(int64_t n, int c) {
  const int rpp = kNormThreads / (c / 8);
  int64_t blocks = (n + rpp * 4 - 1) / (rpp * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace b2m

using namespace b2m;

extern "C" int b2m_cast_pad_bf16(const float* x, int64_t n, int32_t c, int32_t c_pad, uint16_t* out, b2m_stream_t stream) {
  if (!x || !out || n < 0 || c <= 0 || c_pad < c) return B2M_ERR_INVALID_ARGUMENT;
  if (n == 0) return B2M_OK;
  cast_pad_kernel<<<cdiv(n * c_pad, 256), 256, 0, (cudaStream_t)stream>>>(x, n, c, c_pad, out);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

static bool norm_shape_ok(int c) { return c > 0 && c % 8 == 0 && c <= 2048 && (c / 8) <= kNormThreads; }

extern "C" int b2m_colstats(const uint16_t* x, int64_t n, int32_t c, double* sums, b2m_stream_t stream) {
  if (!x || !sums || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (!norm_shape_ok(c)) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  const int rpp = kNormThreads / (c / 8);
  const size_t sh = (size_t)rpp * c * 2 * sizeof(float);
  colreduce_kernel<0><<<reduce_grid(n, c), kNormThreads, sh, (cudaStream_t)stream>>>(x, nullptr, nullptr, n, c, nullptr,
                                                                                    nullptr, 0, nullptr, sums);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_bn_forward(const uint16_t* x, int64_t n, int64_t n_stat, int32_t c, const double* sums, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                              int32_t training, const uint16_t* residual, int32_t relu, uint16_t* out, float* save_mean,
                              float* save_invstd, uint8_t* relu_mask, b2m_stream_t stream) {
  if (!x || !gamma || !beta || !out || !save_mean || !save_invstd || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (training && !sums) return B2M_ERR_INVALID_ARGUMENT;
  if (!training && (!running_mean || !running_var)) return B2M_ERR_INVALID_ARGUMENT;
  if (!norm_shape_ok(c)) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // Running statistics are updated by block 0 while the other blocks may still read them in eval mode only, where they
  // are not written; in training mode the normalisation uses the batch sums, so there is no read/write hazard.
  bn_apply_kernel<<<apply_grid(n, c), kNormThreads, 0, st>>>(x, n, n_stat, c, sums, gamma, beta, running_mean, running_var,
                                                            momentum, eps, training, residual, relu, out, save_mean,
                                                            save_invstd, relu ? relu_mask : nullptr);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_bn_backward_reduce(const uint16_t* x, const uint16_t* out, const uint16_t* dout, int64_t n, int32_t c,
                                      const float* save_mean, const float* save_invstd, int32_t relu, double* red,
                                      const uint8_t* relu_mask, b2m_stream_t stream) {
  if (!x || !dout || !save_mean || !save_invstd || !red || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (relu && !out && !relu_mask) return B2M_ERR_INVALID_ARGUMENT;
  if (!norm_shape_ok(c)) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  const int rpp = kNormThreads / (c / 8);
  const size_t sh = (size_t)rpp * c * 2 * sizeof(float);
  colreduce_kernel<1><<<reduce_grid(n, c), kNormThreads, sh, (cudaStream_t)stream>>>(x, out, dout, n, c, save_mean,
                                                                                    save_invstd, relu, relu_mask, red);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}

extern "C" int b2m_bn_backward_apply(const uint16_t* x, const uint16_t* out, const uint16_t* dout, int64_t n,
                                     int64_t n_stat, int32_t c,
                                     const float* save_mean, const float* save_invstd, const float* gamma,
                                     const double* red, const double* red_local, const double* n_stat_dev,
                                     int32_t relu, int32_t training, uint16_t* dx,
                                     uint16_t* dresidual, float* dgamma, float* dbeta, const uint8_t* relu_mask,
                                     b2m_stream_t stream) {
  if (!x || !dout || !save_mean || !save_invstd || !gamma || !red || !dx || n < 0) return B2M_ERR_INVALID_ARGUMENT;
  if (relu && !out && !relu_mask) return B2M_ERR_INVALID_ARGUMENT;
  if (!norm_shape_ok(c)) return B2M_ERR_UNSUPPORTED_SHAPE;
  if (n == 0) return B2M_OK;
  bn_bwd_apply_kernel<<<apply_grid(n, c), kNormThreads, 0, (cudaStream_t)stream>>>(
      x, out, dout, n, n_stat, c, save_mean, save_invstd, gamma, red, red_local, n_stat_dev, relu, training, dx, dresidual,
      dgamma, dbeta, relu_mask);
  B2M_CHECK_LAUNCH();
  return B2M_OK;
}
