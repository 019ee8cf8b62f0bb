/*
 * b2m.h — C-ABI of the B200-native (sm_100a) hot path of Box2Mask.
 *
 * This is the boundary the Python host (box2mask_b200/, importable as `MinkowskiEngine`) binds with
 * ctypes. It replaces what the reference reaches through `import MinkowskiEngine as ME`
 * (MinkowskiEngine==0.5.4, pinned at /root/reference/docs/installation.md:6,42; not vendored) and the
 * CPU torch loops of /root/reference/models/iou_nms.py. Each entry point cites the reference call site
 * (file:line, relative to /root/reference) whose work it performs.
 *
 * Conventions
 *  - plain pointers + sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - no allocation inside: outputs and workspaces are caller-allocated; `*_workspace_bytes` queries.
 *  - stream-ordered on `stream` (a cudaStream_t passed as void*); no host synchronisation inside
 *    unless stated; re-entrant. The only process state: a per-device cache (SM count, shared-memory
 *    opt-in of the kernels) and the explicit tuning options of b2m_set_option.
 *  - return 0 (B2M_OK) or a negative error code; never throws. `b2m_error_string` names a code.
 *  - "bf16" = __nv_bfloat16 bit patterns carried as uint16_t.
 *  - Kernel maps are dense neighbour tables `nbr[K][pitch]` (int32, -1 = no neighbour), pitch =
 *    b2m_map_pitch(n_out) = n_out rounded up to 128 (columns >= n_out hold -1, so a 128-row MMA tile and
 *    the 16-byte index loads of the TMA row gathers never run off a row of the table): entry
 *    (k, o) is the input row whose coordinate equals coord(o) + delta_k. Offsets enumerate with the
 *    first spatial axis fastest, k = ix + K*iy + K*K*iz; odd kernels are centred, even kernels start
 *    at 0 (MinkowskiEngine region-iterator convention).
 */
#ifndef B2M_H_
#define B2M_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2M_OK 0
#define B2M_ERR_INVALID_ARGUMENT (-1)
#define B2M_ERR_CUDA_LAUNCH (-2)
#define B2M_ERR_WORKSPACE_TOO_SMALL (-3)
#define B2M_ERR_UNSUPPORTED_SHAPE (-4)
#define B2M_ERR_COORD_RANGE (-5)

typedef void* b2m_stream_t; /* cudaStream_t */

int b2m_version(void);
const char* b2m_error_string(int code);

/* Process-wide tuning options (no environment variables are read inside the library).
 *  B2M_OPT_MAX_CTAS          cap on the CTAs of the persistent convolution kernels (0 = one per SM); lowered by
 *                            the data-parallel host while a gradient all-reduce overlaps the backward pass
 *  B2M_OPT_CHUNKS_PER_STAGE  force 1 or 2 reduction chunks per pipeline stage of the forward kernel (0 = auto)
 *  B2M_OPT_SPLIT_OFFSETS     1 = never split the kernel offsets of a convolution over CTAs (0 = automatic)
 *  B2M_OPT_GATHER_MODE       0 = forward / dgrad gather feature rows with cp.async, wgrad with TMA gather4 (default),
 *                            1 = TMA gather4 everywhere, 2 = cp.async with the generic->async proxy fence on the
 *                            producer side (forward kernel), 3 = cp.async in wgrad too
 *  B2M_OPT_ISSUER            1 = general MMA issue loop of the forward kernel (0 = lean loop where it applies)
 *  B2M_OPT_WGRAD_BSLOTS      dY ring depth of the wgrad kernel (0 = automatic)
 *  B2M_OPT_WGRAD_GROUP       2 = two gather warps per wgrad pipeline stage (0 = one)
 *  B2M_OPT_WGRAD_ROWS        64 = 64 reduction rows per wgrad pipeline stage always (0 = 128 on large levels) */
#define B2M_OPT_MAX_CTAS 1
#define B2M_OPT_CHUNKS_PER_STAGE 2
#define B2M_OPT_SPLIT_OFFSETS 3
#define B2M_OPT_GATHER_MODE 4
#define B2M_OPT_WGRAD_ROWS 5
#define B2M_OPT_ISSUER 6
#define B2M_OPT_WGRAD_GROUP 8
#define B2M_OPT_WGRAD_BSLOTS 9
int b2m_set_option(int32_t option, int64_t value);

/* ------------------------------------------------------------------------------------------------
 * Coordinate hash (open addressing, 64-bit packed keys)           reference: ME.SparseTensor(...)
 * models/model.py:43, models/detection_net.py:499,503 (coordinate-manager insert)
 * ---------------------------------------------------------------------------------------------- */
/* number of slots for n coordinates: power of two >= 2n, >= 1024 */
int64_t b2m_hash_capacity(int64_t n);
/* Insert rows 0..n-1 of coords int32[n,4] = (b,x,y,z). table_keys uint64[capacity], table_vals
 * int32[capacity] are (re)initialised inside. status int32[2] (device): [0] = number of rows whose
 * coordinate was already present (duplicates keep the LOWEST row index, i.e. first occurrence),
 * [1] = number of rows outside the packable range [-32768, 32767]. */
int b2m_hash_build(const int32_t* coords, int64_t n, uint64_t* table_keys, int32_t* table_vals,
                   int64_t capacity, int32_t* status, b2m_stream_t stream);
/* rows[m] = row index of each query coordinate, -1 if absent */
int b2m_hash_query(const int32_t* query_coords, int64_t m, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t capacity, int32_t* rows, b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Strided coordinate map          reference: MinkowskiConvolution(kernel_size=2, stride=2)
 * models/detection_net.py:42,48,54,61,68,74,81 (implied coordinate-map stride)
 * out = unique rows of floor(c / new_stride) * new_stride (batch kept), sorted lexicographically by
 * (b,x,y,z) — the same rows and order as np.unique(axis=0). parent_row[i] = output row of input i.
 * n_out (device int32[1]) receives the number of output rows; out_coords has room for n rows.
 * ---------------------------------------------------------------------------------------------- */
size_t b2m_downsample_workspace_bytes(int64_t n);
int b2m_downsample_coords(const int32_t* coords, int64_t n, int32_t new_stride, int32_t* out_coords,
                          int32_t* parent_row, int32_t* n_out, void* workspace, size_t workspace_bytes,
                          b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Kernel maps
 * ---------------------------------------------------------------------------------------------- */
/* row pitch of every neighbour table / `order` array over n output rows: n rounded up to a multiple of 128 */
int64_t b2m_map_pitch(int64_t n);
/* stride-1 (submanifold) map for an odd kernel (3 or 5): nbr int32[K^3, pitch(n)].
 * reference: MinkowskiConvolution(kernel_size=3) models/resnet.py:61-65; kernel_size=5
 * models/detection_net.py:37 */
int b2m_kernel_map_submanifold(const int32_t* coords, int64_t n, int32_t tensor_stride,
                               int32_t kernel_size, const uint64_t* table_keys,
                               const int32_t* table_vals, int64_t capacity, int32_t* nbr,
                               b2m_stream_t stream);
/* The same table as b2m_kernel_map_submanifold (3^3 or 5^3, odd kernels only), built WITHOUT hashing from the next
 * coarser level: parent_row int32[n] and nbr_down int32[8, pitch(n_coarse)] from b2m_downsample_coords /
 * b2m_kernel_map_stride2 (unsorted), nbr3_coarse int32[27, pitch(n_coarse)] = the coarse level's own 3^3 table
 * (unsorted). nbr[k][o] = nbr_down[slot][nbr3_coarse[kP][parent_row[o]]]: two reads of small dense tables per entry.
 * group_mask (optional, uint32[ceil(n/64), ceil(K/32)], zeroed inside) receives the offsets present in each 64-row group
 * of the UNSORTED table (what a map that is consumed in its original row order needs, e.g. the 125-offset map).
 * workspace (optional, b2m_kernel_map_from_coarse_workspace_bytes(n_coarse) bytes, 16-byte aligned): room for the child
 * table transposed to [coarse row][8]; with it the table is built one thread per row (27 + 27 table reads per row and
 * coalesced writes) instead of one thread per entry; the result is identical. */
size_t b2m_kernel_map_from_coarse_workspace_bytes(int64_t n_coarse);
int b2m_kernel_map_from_coarse(const int32_t* coords, int64_t n, int32_t tensor_stride, int32_t kernel_size,
                               const int32_t* parent_row, const int32_t* nbr3_coarse, const int32_t* nbr_down,
                               int64_t n_coarse, int32_t* nbr, uint32_t* group_mask, void* workspace,
                               size_t workspace_bytes, b2m_stream_t stream);
/* kernel 2 / stride 2 maps from the parent relation of b2m_downsample_coords.
 * nbr_down int32[8, pitch(n_coarse)]: child row of coarse row o at offset k (strided conv, coarse output).
 * nbr_up   int32[8, pitch(n_fine)]  : parent row if offset(f)==k else -1 (transposed conv, fine output).
 * reference: models/detection_net.py:42-82 (down) and :88-133 (MinkowskiConvolutionTranspose) */
int b2m_kernel_map_stride2(const int32_t* fine_coords, int64_t n_fine, const int32_t* parent_row,
                           int64_t n_coarse, int32_t fine_stride, int32_t* nbr_down, int32_t* nbr_up,
                           b2m_stream_t stream);
/* ME-style pair lists from a dense table: counts int32[K] (pairs per offset). With in_rows/out_rows
 * non-null and offsets int32[K] (exclusive prefix of counts) also fills the lists, each offset's pairs
 * ordered by output row. Used by tests/inspection; the convolutions consume the dense table. */
int b2m_kernel_map_count(const int32_t* nbr, int32_t kvol, int64_t n_out, int32_t* counts,
                         b2m_stream_t stream);

/* Sorted kernel map, the form the convolutions consume. Rows are ordered by (row / block_rows, occupancy bit
 * mask of the row) with a stable radix sort, so that the 128 rows of an MMA tile share their set of present
 * offsets and (tile, offset) blocks without any pair are skipped. Outputs:
 *   order      int32[pitch]        position -> original output row (-1 in the padding)
 *   nbr_sorted int32[K, pitch]     nbr_sorted[k][j] = nbr[k][order[j]]
 *   group_mask uint32[ceil(n_out/64), ceil(K/32)]  offsets present in each 64-row group of the sorted order
 * block_rows <= 0 or K > 32 keeps the original order (identity `order`), masks are still produced.
 * The set of (k, in, out) pairs is unchanged — this is a traversal order, not a different map. */
size_t b2m_kernel_map_sort_workspace_bytes(int64_t n_out);
int b2m_kernel_map_sort(const int32_t* nbr, int32_t kvol, int64_t n_out, int32_t block_rows, int32_t* order,
                        int32_t* nbr_sorted, uint32_t* group_mask, void* workspace, size_t workspace_bytes,
                        b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse convolution (gather -> tcgen05 implicit GEMM, fp32 accumulation in TMEM)
 * reference: MinkowskiConvolution.forward / MinkowskiConvolutionTranspose.forward and their autograd
 * backward — models/detection_net.py:235-337, models/resnet.py:70-83, models/training.py:68
 * ---------------------------------------------------------------------------------------------- */
/* fp32 rows -> bf16 rows with the channel count zero-padded to c_pad (multiple of 16) */
int b2m_cast_pad_bf16(const float* x, int64_t n, int32_t c, int32_t c_pad, uint16_t* out,
                      b2m_stream_t stream);
/* Weight packing: fp32 kernel[kvol, c_in, c_out] -> bf16 UMMA B-operand image (K-major, 128B swizzle).
 * mode 0: forward operand   B[k]  = W[k]            (reduction dim = c_in,  N = c_out)
 * mode 1: dgrad, same-coords B[k]  = W[kvol-1-k]^T  (reduction dim = c_out, N = c_in)
 * mode 2: dgrad, strided     B[k]  = W[k]^T         (reduction dim = c_out, N = c_in)
 * c_red (the reduction dim) is padded to a multiple of 64 with zeros; for c_red in {16, 32} the slices are
 * "flat": 64 / c_red consecutive offsets share one 64-wide slice (what b2m_conv_forward expects). */
size_t b2m_packed_weight_bytes(int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode);
int b2m_pack_weights(const float* kernel, int32_t kvol, int32_t c_in, int32_t c_out, int32_t mode,
                     uint16_t* packed, b2m_stream_t stream);
/* The same for every convolution of a network in ONE launch (one call per training step instead of one per layer).
 * All arrays are DEVICE arrays: kernels[j] fp32 [kvol, c_in_src, c_out], packed[j] = destination of
 * b2m_packed_weight_bytes(kvol, c_in, c_out, mode) bytes, meta int32[n_jobs, 5] = {kvol, c_in_src, c_in, c_out, mode}
 * (input channels c_in_src..c_in-1 are zero: the 6-channel input padded to 8), group_prefix int64[n_jobs + 1] =
 * running sum of packed bytes / 16; total_groups = group_prefix[n_jobs]. */
int b2m_pack_weights_batched(const float* const* kernels, uint16_t* const* packed, const int32_t* meta,
                             const int64_t* group_prefix, int32_t n_jobs, int64_t total_groups,
                             b2m_stream_t stream);
/* y[order[j], :] = sum_k x[nbr[k][j], :] * B[k]   (entries < 0 contribute zero). (nbr, order, group_mask) is a
 * sorted kernel map from b2m_kernel_map_sort; nbr == NULL means the identity map with kvol == 1 (order and
 * group_mask ignored). x bf16[n_in, c_red], y bf16[n_out, c_n]. c_red % 16 == 0, c_n % 16 == 0, c_n <= 512.
 * colsum (optional, double[2*c_n], caller-zeroed) accumulates per-column sum and sum of squares of the fp32
 * results (BatchNorm batch statistics). */
int b2m_conv_forward(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, const int32_t* order,
                     const uint32_t* group_mask, int32_t kvol, int64_t n_out, const uint16_t* packed_w,
                     int32_t c_n, uint16_t* y, double* colsum, b2m_stream_t stream);
/* The same with a fused epilogue and the offset-split path (all extras optional / NULL):
 *   v = acc * scale[col] + shift[col] (+ residual[row, col]) (ReLU if relu)      scale, shift float[c_n]; residual bf16[n_out, c_n]
 *   - eval-mode BatchNorm folded into the convolution (scale = gamma / sqrt(var + eps), shift = beta - mean * scale),
 *     the residual add and ReLU of a BasicBlock (models/resnet.py:70-83), bias + ReLU of the MLP heads
 *     (models/detection_net.py:170-194; scale NULL, shift = bias);
 *   - y32 != NULL: the result is written as fp32 rows of c_store <= c_n real columns (class logits of a head whose
 *     width was padded to a multiple of 16) instead of bf16 y;
 *   - colsum accumulates the statistics of v (after the epilogue);
 *   - workspace of b2m_conv_forward_workspace_bytes(n_out, c_red, kvol, c_n) bytes (0 = not needed): on levels with few
 *     128-row tiles the kernel offsets are dealt to several CTAs per tile, each writes an fp32 partial tile into the
 *     workspace and a finalize kernel sums them in a fixed order (deterministic) and applies the epilogue. Without a
 *     workspace the launch runs unsplit. */
size_t b2m_conv_forward_workspace_bytes(int64_t n_out, int32_t c_red, int32_t kvol, int32_t c_n);
int b2m_conv_forward_ex(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, const int32_t* order,
                        const uint32_t* group_mask, int32_t kvol, int64_t n_out, const uint16_t* packed_w,
                        int32_t c_n, uint16_t* y, double* colsum, const float* scale, const float* shift,
                        const uint16_t* residual, int32_t relu, float* y32, int32_t c_store, void* workspace,
                        size_t workspace_bytes, b2m_stream_t stream);
/* b2m_conv_forward_ex used as the dgrad of a unit (x = dL/d(conv output), packed_w = the mirrored / transposed weights,
 * residual = the gradient already pending for the unit's input) with the BatchNorm-backward reduction of the layer that
 * PRODUCED that input fused into its epilogue: bn_red double[2 * c_n] (zeroed by the caller) += (sum g, sum g * xhat), g =
 * the bf16 result row gated by bn_relu_mask (uint8[n_out, c_n / 8], NULL = no ReLU), xhat = (bn_x - bn_mean) * bn_invstd
 * with bn_x bf16[n_out, c_n] the producer's pre-BatchNorm rows - exactly what b2m_bn_backward_reduce computes in a pass of
 * its own (models/resnet.py:63-67 backward). bn_red == NULL: plain b2m_conv_forward_ex. With bn_red: no colsum, scale,
 * shift, relu, y32. */
int b2m_conv_dgrad_bn_reduce(const uint16_t* x, int64_t n_in, int32_t c_red, const int32_t* nbr, const int32_t* order,
                             const uint32_t* group_mask, int32_t kvol, int64_t n_out, const uint16_t* packed_w,
                             int32_t c_n, uint16_t* y, double* colsum, const float* scale, const float* shift,
                             const uint16_t* residual, int32_t relu, float* y32, int32_t c_store, void* workspace,
                             size_t workspace_bytes, const uint16_t* bn_x, const uint8_t* bn_relu_mask,
                             const float* bn_mean, const float* bn_invstd, double* bn_red, b2m_stream_t stream);
/* dw[k, ci, co] = sum_j x[nbr[k][j], ci] * dy[order[j], co]   (fp32; dw is OVERWRITTEN: with one CTA per element
 * - few row groups - by plain stores, otherwise zero-filled inside and accumulated with fp32 vector reductions)
 * over a sorted kernel map. x bf16[n_in, c_in], dy bf16[n_out, c_out]; c_in % 8 == 0, c_out % 16 == 0,
 * c_out <= 256. */
int b2m_conv_wgrad(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                   const int32_t* nbr, const int32_t* order, const uint32_t* group_mask, int32_t kvol,
                   int64_t n_out, float* dw, b2m_stream_t stream);
/* The same with a caller-provided workspace (b2m_conv_wgrad_workspace_bytes; 0 = none needed): the row splits then write
 * partial sums with plain stores and a second kernel adds them in a fixed order - deterministic, and no contended
 * atomics. Without (or with too small) a workspace it behaves like b2m_conv_wgrad. */
size_t b2m_conv_wgrad_workspace_bytes(int64_t n_out, int32_t c_in, int32_t c_out, int32_t kvol);
int b2m_conv_wgrad_ex(const uint16_t* x, int64_t n_in, int32_t c_in, const uint16_t* dy, int32_t c_out,
                      const int32_t* nbr, const int32_t* order, const uint32_t* group_mask, int32_t kvol,
                      int64_t n_out, float* dw, void* workspace, size_t workspace_bytes, b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm (+ residual, + ReLU) over rows          reference: MinkowskiBatchNorm -> BatchNorm1d
 * models/resnet.py:63,66,159; models/detection_net.py:40..187; ReLU / `out += residual`
 * models/resnet.py:67,80-81
 * ---------------------------------------------------------------------------------------------- */
/* sums double[2c] += (sum_x, sum_x^2) per column of x bf16[n,c] */
int b2m_colstats(const uint16_t* x, int64_t n, int32_t c, double* sums, b2m_stream_t stream);
/* n_stat = number of rows the sums were taken over (== n on one GPU; the all-rank total under SyncBN;
 * n_stat <= 0: read it from sums[2c], where the SyncBN all-reduce carries the global row count).
 * training: mean/var from sums (biased var for normalisation), running stats updated with momentum
 * (unbiased var), save_mean/save_invstd float[c] written. eval (training==0): uses running stats.
 * out = act( (x-mean)*invstd*gamma + beta (+ residual) ), act = ReLU if relu else identity.
 * relu_mask (optional, uint8[n, c/8], written when relu): bit i of byte (row, g) = (out[row, 8 g + i] > 0), the gate
 * the backward passes need - 1/16 of the bytes of `out`, which they re-read otherwise. */
int b2m_bn_forward(const uint16_t* x, int64_t n, int64_t n_stat, int32_t c, const double* sums, const float* gamma,
                   const float* beta, float* running_mean, float* running_var, float momentum,
                   float eps, int32_t training, const uint16_t* residual, int32_t relu, uint16_t* out,
                   float* save_mean, float* save_invstd, uint8_t* relu_mask, b2m_stream_t stream);
/* pass 1: red double[2c] += (sum_g, sum_g*xhat), g = dout masked by (out > 0) when relu (read from relu_mask when
 *         given - `out` may then be NULL - else from `out`).
 * pass 2: dx = gamma*invstd*(g - sum_g/n - xhat*sum_gxhat/n); dresidual = g (optional);
 *         dgamma = sum_gxhat, dbeta = sum_g. In eval mode dx = gamma*invstd*g. */
int b2m_bn_backward_reduce(const uint16_t* x, const uint16_t* out, const uint16_t* dout, int64_t n,
                           int32_t c, const float* save_mean, const float* save_invstd, int32_t relu,
                           double* red, const uint8_t* relu_mask, b2m_stream_t stream);
/* SyncBatchNorm (models/model.py:25): `red` is the all-reduced (global) reduction that enters dx; `red_local`
 * (NULL = red) is this rank's own reduction, which becomes dgamma / dbeta (torch's SyncBatchNorm keeps the affine
 * gradients local; the gradient all-reduce averages them like any other parameter); `n_stat_dev` (NULL = use
 * n_stat) points at the global row count as a double on the device (no host round trip). */
int b2m_bn_backward_apply(const uint16_t* x, const uint16_t* out, const uint16_t* dout, int64_t n,
                          int64_t n_stat, int32_t c, const float* save_mean, const float* save_invstd,
                          const float* gamma, const double* red, const double* red_local,
                          const double* n_stat_dev, int32_t relu, int32_t training,
                          uint16_t* dx, uint16_t* dresidual, float* dgamma, float* dbeta,
                          const uint8_t* relu_mask, b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Superpoint pooling                reference: models/detection_net.py:345-352 (re-keyed
 * SparseTensor + MinkowskiGlobalAvgPooling / MinkowskiGlobalMaxPooling), ids utils/util.py:123-130
 * ---------------------------------------------------------------------------------------------- */
/* out float[s, c] = mean over rows v with ids[v] == s of f bf16[n, c]; counts float[s] written.
 * out and counts are zeroed inside. */
int b2m_segment_mean_forward(const uint16_t* f, const int64_t* ids, int64_t n, int32_t c, int64_t s,
                             float* out, float* counts, b2m_stream_t stream);
/* df bf16[n, c] = dout[ids[v], :] / counts[ids[v]] */
int b2m_segment_mean_backward(const float* dout, const int64_t* ids, const float* counts, int64_t n,
                              int32_t c, int64_t s, uint16_t* df, b2m_stream_t stream);
/* out float[s, c] = max over the segment, argmax int32[s, c] = row that attained it */
int b2m_segment_max_forward(const uint16_t* f, const int64_t* ids, int64_t n, int32_t c, int64_t s,
                            float* out, int32_t* argmax, b2m_stream_t stream);
/* df bf16[n, c] = gradient of the segment max: row argmax[s, j] receives dout[s, j], all else zero */
int b2m_segment_max_backward(const float* dout, const int32_t* argmax, int64_t s, int32_t c, int64_t n,
                             uint16_t* df, b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Box-vote decoding                 reference: models/iou_nms.py and models/detection_net.py:369-488
 * ---------------------------------------------------------------------------------------------- */
/* NMS_clustering (models/iou_nms.py:68-105). boxes float[m,7] = (score, min xyz, max xyz).
 * Greedy in order of descending score (ties: lower index first). Outputs:
 *  n_clusters int32[1]; representatives int32[m] (first n_clusters valid, box index of each cluster);
 *  cluster_of int32[m] (cluster index each box was assigned to);
 *  heatmaps float[max_clusters, m] (row c = IoU of representative c with ALL boxes, self = 1) — rows
 *  beyond max_clusters are not written (n_clusters still counts them).
 * IoU arithmetic is the reference's fp32 operation order with IEEE rounding and no FMA contraction:
 *  I = (dx*dy)*dz with d = max(min(max_a,max_b) - max(min_a,min_b), 0);
 *  U = ((A_r + A_j) - I) + 1e-6f;  iou = I / U. */
size_t b2m_nms_workspace_bytes(int64_t m);
int b2m_aabb_nms(const float* boxes, int64_t m, float cluster_th, int32_t* n_clusters,
                 int32_t* representatives, int32_t* cluster_of, float* heatmaps, int64_t max_clusters,
                 void* workspace, size_t workspace_bytes, b2m_stream_t stream);
/* Heat-map rows of the first k clusters, heat float[k, m], for callers that read n_clusters back first and allocate k
 * rows instead of m (b2m_aabb_nms with heatmaps == NULL, then this). */
int b2m_aabb_heatmaps(const float* boxes, int64_t m, const int32_t* n_clusters, const int32_t* representatives,
                      int64_t k, float* heatmaps, b2m_stream_t stream);
/* Heat-map rows -> bit-packed voxel masks (models/detection_net.py:436-446):
 *  mask[c, v] = heat[c, fg_rank[seg2vox[v]]] > mask_bin_th, 0 where fg_rank < 0 (background).
 *  heat float[k, m_fg], fg_rank int32[s] (rank of superpoint among foreground ones or -1),
 *  seg2vox int64[n_vox]; masks uint32[k, words], words = ceil(n_vox/32); bit v%32 of word v/32. */
int b2m_heatmap_project(const float* heat, int64_t k, int64_t m_fg, const int32_t* fg_rank,
                        const int64_t* seg2vox, int64_t n_vox, float mask_bin_th, uint32_t* masks,
                        b2m_stream_t stream);
/* mask_NMS (models/iou_nms.py:130-144) on bit-packed masks given in score order.
 *  keep uint8[k] = 1 for kept masks; n_keep int32[1]. IoU = float(|a&b|) / float(|a|b|) > th suppresses.
 *  workspace: int32[k*k] intersections + int32[k] areas. */
size_t b2m_mask_nms_workspace_bytes(int64_t k);
int b2m_mask_nms(const uint32_t* masks, int64_t k, int64_t words, float th, uint8_t* keep,
                 int32_t* n_keep, void* workspace, size_t workspace_bytes, b2m_stream_t stream);
/* Per-instance majority label (np.bincount + argmax, models/detection_net.py:461-466) on bit-packed masks:
 * counts int32[k, n_labels] (zeroed inside), best int32[k] = arg-max, lowest label on ties. label int32[n_vox]. */
int b2m_mask_label_vote(const uint32_t* masks, int64_t k, int64_t words, int64_t n_vox, const int32_t* label,
                        int32_t n_labels, int32_t* counts, int32_t* best, b2m_stream_t stream);
/* Per-segment mode of per-voxel labels (torch.mode, models/detection_net.py:398-410; S3DIS branch):
 * counts int32[n_seg, n_labels] (zeroed inside), best int32[n_seg] = mode, lowest label on ties. */
int b2m_segment_label_vote(const int64_t* seg, const int32_t* label, int64_t n, int64_t n_seg, int32_t n_labels,
                           int32_t* counts, int32_t* best, b2m_stream_t stream);
/* unpack bit masks to bool bytes uint8[k, n_vox] */
int b2m_unpack_masks(const uint32_t* masks, int64_t k, int64_t words, int64_t n_vox, uint8_t* out,
                     b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Voxelisation of a scene's points (the step before the path)
 * reference: models/dataloader.py:61-77 (numpy round + unique, scikit-learn ball-tree 1-NN per voxel)
 * ---------------------------------------------------------------------------------------------- */
/* coords int32[n,4] = (0, rint((p - min(0, *min_position)) / voxel_size)) for positions f64[n,3] (fp64, the
 * reference's operation order; rint = np.round's half-to-even). The sorted unique voxel set and the
 * point -> voxel map follow from b2m_downsample_coords(coords, stride 1). status int32[1]: points outside the
 * packable coordinate range. min_position: DEVICE scalar, the minimum over all 3n position components. */
int b2m_voxel_coords(const double* positions, int64_t n, const double* min_position, double voxel_size,
                     int32_t* coords, int32_t* status, b2m_stream_t stream);
/* nearest[v] = index of the scene point closest to the centre of voxel v (ties: lower index), searched
 * among the points of the 27 neighbouring voxels, which is exact (see csrc/voxel.cu). vox_coords
 * int32[n_vox,4]; nbr = its k=3 table from b2m_kernel_map_submanifold (pitch b2m_map_pitch(n_vox));
 * start int64[n_vox+1], point_order int64[n]: points grouped by voxel (CSR). */
int b2m_nearest_point(const double* positions, const double* min_position, double voxel_size,
                      const int32_t* vox_coords, int64_t n_vox, const int32_t* nbr, const int64_t* start,
                      const int64_t* point_order, int64_t* nearest, b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Weak-supervision label association (next row 2)      reference: ScanNet.approx_association,
 * models/dataloader.py:203-314 (ARKitScenes :539-621 and S3DIS :805-927 use the same steps)
 * ---------------------------------------------------------------------------------------------- */
/* Point-in-box occupancy. positions f64[n,3]; box_min / box_max f64[n_boxes,3] (inclusive on both sides, float64
 * comparisons like numpy); volume f64[n_boxes]. num[i] = boxes containing point i; first[i] = the lowest such box index
 * (-1 if none); smallest[i] = the containing box of least volume, lowest index on ties (-1 if none).
 * reference: dataloader.py:236-242 (`is_within_bb`, `bb_occupancy.sum`, `np.argmin(bb_volume[box_ids])`). */
int b2m_point_box_occupancy(const double* positions, int64_t n, const double* box_min, const double* box_max,
                            const double* volume, int32_t n_boxes, int32_t* num, int32_t* first, int32_t* smallest,
                            b2m_stream_t stream);
/* inst[i] = -1 (no box), the instance id of the single box, or for several boxes -2 / the smallest box's instance
 * (smallest_heuristic). reference: dataloader.py:243-259. */
int b2m_point_instances(const int32_t* num, const int32_t* first, const int32_t* smallest, const int64_t* instance_ids,
                        int64_t n, int32_t smallest_heuristic, int64_t* inst, b2m_stream_t stream);
/* Per-superpoint decision pooled back to the points. seg_rank int32[n]: index of the point's superpoint in the caller's
 * unique-segment list, -1 if it is not in the list (such points get -2). majority_vote = 0: the superpoint follows its
 * least-covered point (lowest point index on ties): 1 box -> that instance, 0 boxes -> -1, several -> -2 or the smallest
 * box (dataloader.py:278-312). majority_vote = 1: most common per-point value, smallest value on ties (:264-274); needs
 * sorted_ids int64[n_boxes] (instance ids ascending) and id_rank int32[n_boxes] (rank of each box's id in that order). */
size_t b2m_segment_association_workspace_bytes(int64_t n_segs, int32_t n_boxes, int32_t majority_vote);
int b2m_segment_association(const int32_t* num, const int32_t* first, const int32_t* smallest, const int32_t* seg_rank,
                            int64_t n, int64_t n_segs, const int64_t* instance_ids, const int64_t* sorted_ids,
                            const int32_t* id_rank, int32_t n_boxes, int32_t majority_vote, int32_t smallest_heuristic,
                            int64_t* per_seg, int64_t* per_point, void* workspace, size_t workspace_bytes,
                            b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SyncBatchNorm statistics over NVLink peer memory      reference: MinkowskiSyncBatchNorm under cfg.multigpu,
 * models/model.py:25 (per layer one all-reduce of the batch statistics in forward and one of the gradient sums in
 * backward). One node, one process per GPU: every rank creates an exchange buffer, hands its 64-byte CUDA-IPC handle
 * to its peers (the host does that once, e.g. through torch.distributed) and opens theirs.
 * b2m_peer_allreduce_f64: out[0..n) = sum over ranks of in[0..n) (added in rank order: bit-identical on all ranks), as ONE
 * single-CTA kernel that stores this rank's vector into its slot of every peer's buffer, publishes sequence number
 * `seq` and waits for the peers' - no library collective, no host involvement. use_tail: element n-1 of this rank's
 * vector is `tail` instead of in[n-1] (the row count that travels with the sums). peer_buffers: DEVICE array of `world`
 * buffer pointers (index = rank; the own buffer at index `rank`). seq: 1, 2, 3, ... identical on all ranks, consecutive
 * calls on one stream. status (device int32): set to 1 if a peer did not arrive within ~3 s. n <= b2m_peer_max_doubles().
 * ---------------------------------------------------------------------------------------------- */
size_t b2m_peer_buffer_bytes(void);
int32_t b2m_peer_max_doubles(void);
int b2m_peer_buffer_create(void** buffer, void* ipc_handle_64_bytes);       /* synchronises the device once */
int b2m_peer_buffer_open(const void* ipc_handle_64_bytes, void** buffer);
int b2m_peer_buffer_close(void* buffer, int32_t own);
int b2m_peer_allreduce_f64(const double* in, int32_t n, double tail, int32_t use_tail, double* out,
                           void* const* peer_buffers, int32_t rank, int32_t world, uint64_t seq, int32_t* status,
                           b2m_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Launch lists: a whole pass of the U-Net trunk (models/detection_net.py:235-337 - ~250 modules, ~540 calls per training
 * step when bound call by call) in ONE call. A command carries the arguments of one of the entry points above in their
 * declared order (pointers and integers in a[], floating-point arguments in f[]), without the stream; `stream` selects
 * the main (0) or the side (1) stream passed to b2m_run_commands. RECORD / WAIT order the two streams through events
 * of a per-device pool owned by the library (a[0] = event slot): weight gradients run on the side stream beside the
 * dgrad chain.
 * ---------------------------------------------------------------------------------------------- */
#define B2M_CMD_CONV_FORWARD 1       /* b2m_conv_dgrad_bn_reduce: a[0..24] (a[20..24] = bn_x .. bn_red, 0 = plain b2m_conv_forward_ex) */
#define B2M_CMD_CONV_WGRAD 2         /* b2m_conv_wgrad_ex: a[0..12] */
#define B2M_CMD_BN_FORWARD 3         /* b2m_bn_forward: a[0..8] = x..running_var, f[0] = momentum, f[1] = eps, a[9..15] = training..relu_mask */
#define B2M_CMD_BN_BACKWARD_REDUCE 4 /* b2m_bn_backward_reduce: a[0..9] */
#define B2M_CMD_BN_BACKWARD_APPLY 5  /* b2m_bn_backward_apply: a[0..18] */
#define B2M_CMD_COPY_COLUMNS 6       /* b2m_copy_columns: a[0..5] */
#define B2M_CMD_RECORD 7             /* record event slot a[0] on the selected stream */
#define B2M_CMD_WAIT 8               /* the selected stream waits for event slot a[0] */
#define B2M_CMD_PEER_ALLREDUCE 9     /* b2m_peer_allreduce_f64: a[0] = in, a[1] = n, f[0] = tail, a[2] = use_tail, a[3..8] = out..status */
typedef struct b2m_command {
  int32_t op;
  int32_t stream;
  int64_t a[28];
  double f[2];
} b2m_command_t;
/* dst[r, 0:width] = src[r, 0:width] for bf16 rows of pitch src_ld / dst_ld elements (width, pitches multiples of 8,
 * 16-byte aligned pointers): the channel concatenation `ME.cat` of the decoder (models/detection_net.py:286-336) and
 * the split of its gradient. */
int b2m_copy_columns(const uint16_t* src, int64_t src_ld, uint16_t* dst, int64_t dst_ld, int64_t n, int32_t width,
                     b2m_stream_t stream);
/* Issues cmds[0..n) in order. On an error returns its code and, if `failed` is given, the index of the command. */
int b2m_run_commands(const b2m_command_t* cmds, int64_t n, b2m_stream_t main_stream, b2m_stream_t side_stream,
                     int64_t* failed);

#ifdef __cplusplus
}
#endif
#endif /* B2M_H_ */
