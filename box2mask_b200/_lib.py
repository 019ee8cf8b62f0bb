"""ctypes binding of the C-ABI library (include/b2m.h -> box2mask_b200/lib/libb2m.so).

The product path has no CPU fallback: if the library is missing every op raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2M_LIB selects a variant built by tools/build_variant.py (debug / experiment builds); the product library otherwise
LIB_PATH = os.path.abspath(os.environ["B2M_LIB"]) if os.environ.get("B2M_LIB") else os.path.join(_HERE, "lib", "libb2m.so")

_P = c_void_p  # every device pointer travels as void*

# name -> (restype, argtypes); mirrors include/b2m.h one to one
SIGNATURES = {
    "b2m_version": (c_int32, []),
    "b2m_error_string": (c_char_p, [c_int32]),
    "b2m_set_option": (c_int32, [c_int32, c_int64]),
    "b2m_hash_capacity": (c_int64, [c_int64]),
    "b2m_hash_build": (c_int32, [_P, c_int64, _P, _P, c_int64, _P, _P]),
    "b2m_hash_query": (c_int32, [_P, c_int64, _P, _P, c_int64, _P, _P]),
    "b2m_downsample_workspace_bytes": (c_size_t, [c_int64]),
    "b2m_downsample_coords": (c_int32, [_P, c_int64, c_int32, _P, _P, _P, _P, c_size_t, _P]),
    "b2m_kernel_map_submanifold": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, c_int64, _P, _P]),
    "b2m_kernel_map_from_coarse_workspace_bytes": (c_size_t, [c_int64]),
    "b2m_kernel_map_from_coarse": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, c_int64, _P, _P, _P, c_size_t, _P]),
    "b2m_kernel_map_stride2": (c_int32, [_P, c_int64, _P, c_int64, c_int32, _P, _P, _P]),
    "b2m_kernel_map_count": (c_int32, [_P, c_int32, c_int64, _P, _P]),
    "b2m_cast_pad_bf16": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P]),
    "b2m_packed_weight_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "b2m_pack_weights": (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "b2m_pack_weights_batched": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, _P]),
    "b2m_map_pitch": (c_int64, [c_int64]),
    "b2m_kernel_map_sort_workspace_bytes": (c_size_t, [c_int64]),
    "b2m_kernel_map_sort": (c_int32, [_P, c_int32, c_int64, c_int32, _P, _P, _P, _P, c_size_t, _P]),
    "b2m_conv_forward": (c_int32, [_P, c_int64, c_int32, _P, _P, _P, c_int32, c_int64, _P, c_int32, _P, _P, _P]),
    "b2m_conv_forward_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "b2m_conv_forward_ex": (c_int32, [_P, c_int64, c_int32, _P, _P, _P, c_int32, c_int64, _P, c_int32, _P, _P, _P, _P, _P,
                                      c_int32, _P, c_int32, _P, c_size_t, _P]),
    "b2m_conv_dgrad_bn_reduce": (c_int32, [_P, c_int64, c_int32, _P, _P, _P, c_int32, c_int64, _P, c_int32, _P, _P, _P, _P,
                                           _P, c_int32, _P, c_int32, _P, c_size_t, _P, _P, _P, _P, _P, _P]),
    "b2m_conv_wgrad": (c_int32, [_P, c_int64, c_int32, _P, c_int32, _P, _P, _P, c_int32, c_int64, _P, _P]),
    "b2m_conv_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "b2m_conv_wgrad_ex": (c_int32, [_P, c_int64, c_int32, _P, c_int32, _P, _P, _P, c_int32, c_int64, _P, _P, c_size_t, _P]),
    "b2m_colstats": (c_int32, [_P, c_int64, c_int32, _P, _P]),
    "b2m_bn_forward": (c_int32, [_P, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, c_float, c_float, c_int32, _P, c_int32,
                                 _P, _P, _P, _P, _P]),
    "b2m_bn_backward_reduce": (c_int32, [_P, _P, _P, c_int64, c_int32, _P, _P, c_int32, _P, _P, _P]),
    "b2m_bn_backward_apply": (c_int32, [_P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P, c_int32, c_int32, _P,
                                        _P, _P, _P, _P, _P]),
    "b2m_segment_mean_forward": (c_int32, [_P, _P, c_int64, c_int32, c_int64, _P, _P, _P]),
    "b2m_segment_mean_backward": (c_int32, [_P, _P, _P, c_int64, c_int32, c_int64, _P, _P]),
    "b2m_segment_max_backward": (c_int32, [_P, _P, c_int64, c_int32, c_int64, _P, _P]),
    "b2m_segment_max_forward": (c_int32, [_P, _P, c_int64, c_int32, c_int64, _P, _P, _P]),
    "b2m_nms_workspace_bytes": (c_size_t, [c_int64]),
    "b2m_aabb_nms": (c_int32, [_P, c_int64, c_float, _P, _P, _P, _P, c_int64, _P, c_size_t, _P]),
    "b2m_aabb_heatmaps": (c_int32, [_P, c_int64, _P, _P, c_int64, _P, _P]),
    "b2m_mask_label_vote": (c_int32, [_P, c_int64, c_int64, c_int64, _P, c_int32, _P, _P, _P]),
    "b2m_segment_label_vote": (c_int32, [_P, _P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "b2m_heatmap_project": (c_int32, [_P, c_int64, c_int64, _P, _P, c_int64, c_float, _P, _P]),
    "b2m_mask_nms_workspace_bytes": (c_size_t, [c_int64]),
    "b2m_mask_nms": (c_int32, [_P, c_int64, c_int64, c_float, _P, _P, _P, c_size_t, _P]),
    "b2m_unpack_masks": (c_int32, [_P, c_int64, c_int64, c_int64, _P, _P]),
    "b2m_peer_buffer_bytes": (c_size_t, []),
    "b2m_peer_max_doubles": (c_int32, []),
    "b2m_peer_buffer_create": (c_int32, [_P, _P]),
    "b2m_peer_buffer_open": (c_int32, [_P, _P]),
    "b2m_peer_buffer_close": (c_int32, [_P, c_int32]),
    "b2m_peer_allreduce_f64": (c_int32, [_P, c_int32, c_double, c_int32, _P, _P, c_int32, c_int32, ctypes.c_uint64, _P, _P]),
    "b2m_copy_columns": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int32, _P]),
    "b2m_run_commands": (c_int32, [_P, c_int64, _P, _P, _P]),
    "b2m_point_box_occupancy": (c_int32, [_P, c_int64, _P, _P, _P, c_int32, _P, _P, _P, _P]),
    "b2m_point_instances": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P, _P]),
    "b2m_segment_association_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "b2m_segment_association": (c_int32, [_P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P,
                                          c_size_t, _P]),
    "b2m_voxel_coords": (c_int32, [_P, c_int64, _P, c_double, _P, _P, _P]),
    "b2m_nearest_point": (c_int32, [_P, _P, c_double, _P, c_int64, _P, _P, _P, _P, _P]),
}

_lib = None


class B2MError(RuntimeError):
    pass


def load():
    """Load libb2m.so (once). Raises B2MError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2MError(
            "libb2m.so not found at %s - build it with `python -m box2mask_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().b2m_error_string(code).decode()
        raise B2MError("%s failed: %s (%d)" % (what, msg, code))


def ptr(t):
    """Device (or host) pointer of a torch tensor as a void*; None -> NULL."""
    if t is None:
        return None
    return t.data_ptr()          # ctypes converts an int for a c_void_p parameter; no wrapper object per argument


OPT_MAX_CTAS, OPT_CHUNKS_PER_STAGE, OPT_SPLIT_OFFSETS, OPT_GATHER_MODE, OPT_WGRAD_ROWS, OPT_ISSUER, OPT_WGRAD_GROUP, OPT_WGRAD_BSLOTS = 1, 2, 3, 4, 5, 6, 8, 9


def set_option(option, value):
    check(load().b2m_set_option(int(option), int(value)), "set_option")


def stream_ptr(device=None):
    """Current CUDA stream of `device` (default: the current device) as a void*."""
    # the raw handle straight from torch's C layer: `torch.cuda.current_stream()` builds a Stream object and resolves
    # the device through several Python layers (3 400 calls = 5 ms of host time per training step)
    import torch
    C = torch._C
    if device is None:
        idx = C._cuda_getDevice()
    elif isinstance(device, int):
        idx = device
    else:
        idx = torch.device(device).index
        if idx is None:
            idx = C._cuda_getDevice()
    return C._cuda_getCurrentRawStream(idx)
