"""ctypes loader for lib/libvscuda.so (built by csrc/Makefile via __graft_entry__.build())."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libvscuda.so")

VS_OK, VS_EINVAL, VS_EEMPTY, VS_EDIM, VS_ECUDA, VS_ENODEV, VS_ENOMEM, VS_ERANGE = 0, -1, -2, -3, -4, -5, -6, -7

# every symbol include/vscuda.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "vs_init", "vs_shutdown", "vs_last_error", "vs_device_info",
    "vs_ctx_create", "vs_ctx_create_on_stream", "vs_ctx_profile_enable", "vs_ctx_profile_read", "vs_ctx_trace_enable", "vs_ctx_trace_read", "vs_ctx_destroy", "vs_ctx_sync", "vs_ctx_stream", "vs_ctx_launch_count",
    "vs_ctx_slowpath_count", "vs_debug_set_certify_scale", "vs_ctx_timer_start", "vs_ctx_timer_stop",
    "vs_quantize_f32", "vs_quantize_f64", "vs_dequantize_f32", "vs_dequantize_f64",
    "vs_quantize_f32_dev", "vs_quantize_f64_dev",
    "vs_matrix_create", "vs_matrix_create_empty", "vs_matrix_fill_f32_dev", "vs_matrix_load_rows", "vs_matrix_create_dev", "vs_matrix_from_f32_dev", "vs_matrix_retain",
    "vs_matrix_release", "vs_matrix_rows", "vs_matrix_cols", "vs_matrix_read_rows", "vs_matrix_load_spool", "vs_matrix_save_spool", "vs_matrix_gather", "vs_matrix_split_dev", "vs_matrix_split", "vs_reassign_recenter", "vs_recenter_clusters_dev",
    "vs_cosine_1xN", "vs_dot_1xN", "vs_argmax_MxN", "vs_argmax_MxN_dev",
    "vs_index_build", "vs_index_build_assigned", "vs_index_build_dev", "vs_index_create_empty", "vs_index_fill_dev", "vs_index_fill", "vs_index_release",
    "vs_index_rows", "vs_index_lists", "vs_index_cols", "vs_index_list_offsets", "vs_index_read_rows", "vs_index_upload", "vs_index_with_room", "vs_index_append", "vs_index_capacity", "vs_index_list_lengths", "vs_release_cached_memory", "vs_search", "vs_search_flat", "vs_search_flat_gemm", "vs_search_batch_dev", "vs_index_search_batch_dev", "vs_search_dev", "vs_probe_dev", "vs_search_dev_probed",
    "vs_search_resolve", "vs_select_probes", "vs_topk_merge_dev", "vs_topk_merge_packed_dev",
    "vs_kmeans_step", "vs_kmeans", "vs_kmeans_accumulate_dev", "vs_kmeans_finish_dev", "vs_recenter", "vs_debug_set_argmax_gemm_min", "vs_debug_set_fused", "vs_debug_set_list_major", "vs_debug_set_lm_dense_min",
    "vs_sharded_create", "vs_sharded_release", "vs_sharded_rows", "vs_sharded_shards", "vs_sharded_shard_rows", "vs_sharded_build_assigned",
    "vs_sharded_upload", "vs_sharded_search", "vs_sharded_ctx_create", "vs_sharded_ctx_destroy", "vs_sharded_search_ctx",
    "vs_exchange_create", "vs_exchange_handle", "vs_exchange_connect", "vs_exchange_slot", "vs_exchange_merge", "vs_exchange_release",
]


class BackendUnavailable(RuntimeError):
    """libvscuda.so is missing or there is no B200: the product path has no CPU fallback."""


_lib = None
_inited = False
_lock = threading.Lock()


def lib_path():
    return _SO


def load():
    """dlopen libvscuda.so and declare prototypes. Does not touch the GPU."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_SO):
            raise BackendUnavailable(
                f"{_SO} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(_SO)
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        L.vs_last_error.restype = C.c_char_p
        L.vs_ctx_stream.restype = vp
        L.vs_ctx_launch_count.restype = u64
        L.vs_ctx_slowpath_count.restype = u64
        L.vs_matrix_rows.restype = sz
        L.vs_matrix_cols.restype = sz
        L.vs_index_rows.restype = sz
        L.vs_index_lists.restype = sz
        L.vs_index_cols.restype = sz
        L.vs_matrix_retain.restype = None
        L.vs_matrix_release.restype = None
        L.vs_index_release.restype = None
        L.vs_ctx_destroy.restype = None
        L.vs_shutdown.restype = None
        for name in ("vs_ctx_stream", "vs_ctx_launch_count", "vs_ctx_slowpath_count", "vs_ctx_destroy", "vs_ctx_sync",
                     "vs_ctx_timer_start", "vs_matrix_retain", "vs_matrix_release", "vs_matrix_rows", "vs_matrix_cols",
                     "vs_index_release", "vs_index_rows", "vs_index_lists", "vs_index_cols"):
            getattr(L, name).argtypes = [vp]
        L.vs_init.argtypes = [C.c_int]
        L.vs_debug_set_certify_scale.argtypes = [C.c_float]
        L.vs_ctx_create.argtypes = [C.POINTER(vp)]
        L.vs_ctx_create_on_stream.argtypes = [vp, C.POINTER(vp)]
        L.vs_ctx_profile_enable.argtypes = [vp, C.c_int]
        L.vs_ctx_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
        L.vs_ctx_trace_enable.argtypes = [vp, C.c_int]
        L.vs_ctx_trace_read.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
        L.vs_matrix_create_empty.argtypes = [vp, sz, sz, C.POINTER(vp)]
        L.vs_matrix_fill_f32_dev.argtypes = [vp, vp, sz, vp, sz]
        L.vs_matrix_load_rows.argtypes = [vp, vp, sz, vp, sz]
        L.vs_index_list_offsets.argtypes = [vp, vp, vp]
        L.vs_index_read_rows.argtypes = [vp, vp, sz, sz, vp, vp]
        L.vs_index_upload.argtypes = [vp, vp, vp, sz, sz, vp, vp, C.POINTER(vp)]
        L.vs_index_create_empty.argtypes = [vp, vp, vp, C.POINTER(vp)]
        L.vs_exchange_create.argtypes = [C.c_int, C.c_int, sz, C.POINTER(vp)]
        L.vs_exchange_handle.argtypes = [vp, vp, sz]
        L.vs_exchange_connect.argtypes = [vp, vp]
        L.vs_exchange_slot.argtypes = [vp, C.c_int]
        L.vs_exchange_slot.restype = vp
        L.vs_exchange_merge.argtypes = [vp, vp, C.c_uint32, sz, sz, sz, sz, sz, vp, vp, vp]
        L.vs_exchange_release.argtypes = [vp]
        L.vs_exchange_release.restype = None
        L.vs_index_with_room.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
        L.vs_index_append.argtypes = [vp, vp, vp, sz, sz, vp, vp]
        L.vs_index_capacity.argtypes = [vp]
        L.vs_index_capacity.restype = sz
        L.vs_index_list_lengths.argtypes = [vp, vp, vp]
        L.vs_index_fill_dev.argtypes = [vp, vp, vp, vp, vp, u64]
        L.vs_index_fill.argtypes = [vp, vp, vp, sz, sz, vp, vp, u64]
        L.vs_ctx_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
        L.vs_device_info.argtypes = [C.c_char_p, sz, C.POINTER(C.c_int), C.POINTER(sz)]
        L.vs_quantize_f32.argtypes = [vp, vp, sz, sz, vp]
        L.vs_quantize_f64.argtypes = [vp, vp, sz, sz, vp]
        L.vs_dequantize_f32.argtypes = [vp, vp, sz, sz, vp]
        L.vs_dequantize_f64.argtypes = [vp, vp, sz, sz, vp]
        L.vs_quantize_f32_dev.argtypes = [vp, vp, sz, sz, vp]
        L.vs_quantize_f64_dev.argtypes = [vp, vp, sz, sz, vp]
        L.vs_matrix_create.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
        L.vs_matrix_create_dev.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
        L.vs_matrix_from_f32_dev.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
        L.vs_matrix_read_rows.argtypes = [vp, vp, sz, sz, vp]
        L.vs_cosine_1xN.argtypes = [vp, vp, sz, vp, vp]
        L.vs_dot_1xN.argtypes = [vp, vp, sz, vp, vp]
        L.vs_argmax_MxN.argtypes = [vp, vp, vp, vp, vp]
        L.vs_argmax_MxN_dev.argtypes = [vp, vp, vp, vp]
        L.vs_index_build.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz, C.POINTER(vp)]
        L.vs_index_build_assigned.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz, C.POINTER(vp)]
        L.vs_index_build_dev.argtypes = [vp, vp, vp, vp, u64, vp, C.POINTER(vp)]
        L.vs_search.argtypes = [vp, vp, vp, sz, sz, sz, vp, vp, vp]
        L.vs_search_flat.argtypes = [vp, vp, vp, vp, sz, sz, vp, vp, vp]
        L.vs_search_flat_gemm.argtypes = [vp, vp, vp, vp, sz, sz, vp, vp, vp]
        L.vs_search_batch_dev.argtypes = [vp, vp, vp, u64, vp, sz, vp, vp, vp, vp]
        L.vs_debug_set_argmax_gemm_min.argtypes = [sz]
        L.vs_debug_set_fused.argtypes = [C.c_int]
        L.vs_debug_set_list_major.argtypes = [C.c_int]
        L.vs_debug_set_lm_dense_min.argtypes = [C.c_int]
        L.vs_sharded_create.argtypes = [vp, sz, C.POINTER(vp)]
        L.vs_sharded_release.argtypes = [vp]
        L.vs_sharded_release.restype = None
        L.vs_sharded_rows.argtypes = [vp]
        L.vs_sharded_rows.restype = sz
        L.vs_sharded_shards.argtypes = [vp]
        L.vs_sharded_shards.restype = sz
        L.vs_sharded_shard_rows.argtypes = [vp, sz]
        L.vs_sharded_shard_rows.restype = sz
        L.vs_sharded_build_assigned.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz]
        L.vs_sharded_upload.argtypes = [vp, vp, sz, sz, vp, vp]
        L.vs_sharded_search.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
        L.vs_sharded_ctx_create.argtypes = [vp, C.POINTER(vp)]
        L.vs_sharded_ctx_destroy.argtypes = [vp]
        L.vs_sharded_ctx_destroy.restype = None
        L.vs_sharded_search_ctx.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
        L.vs_matrix_gather.argtypes = [vp, vp, vp, sz, C.POINTER(vp)]
        L.vs_matrix_split_dev.argtypes = [vp, vp, vp, sz, vp, vp]
        L.vs_matrix_split.argtypes = [vp, vp, vp, sz, vp, vp]
        L.vs_reassign_recenter.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp]
        L.vs_recenter_clusters_dev.argtypes = [vp, vp, vp, sz, vp, vp]
        L.vs_matrix_load_spool.argtypes = [vp, C.c_char_p, sz, sz, sz, C.POINTER(vp)]
        L.vs_matrix_save_spool.argtypes = [vp, vp, sz, sz, C.c_char_p, C.c_int]
        L.vs_kmeans.argtypes = [vp, vp, sz, vp, sz, sz, vp, vp]
        L.vs_kmeans_accumulate_dev.argtypes = [vp, vp, sz, vp, vp, vp]
        L.vs_kmeans_finish_dev.argtypes = [vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_int)]
        L.vs_index_search_batch_dev.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp]
        L.vs_search_dev.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp]
        L.vs_probe_dev.argtypes = [vp, vp, vp, sz, vp, vp]
        L.vs_search_dev_probed.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp]
        L.vs_search_resolve.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, C.POINTER(C.c_int)]
        L.vs_select_probes.argtypes = [vp, vp, vp, sz, sz, vp, vp]
        L.vs_topk_merge_dev.argtypes = [vp, vp, vp, vp, sz, sz, sz, vp, vp, vp]
        L.vs_topk_merge_packed_dev.argtypes = [vp, vp, sz, sz, sz, sz, sz, sz, sz, vp, vp, vp]
        L.vs_kmeans_step.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, C.POINTER(C.c_int)]
        L.vs_recenter.argtypes = [vp, vp, vp]
        _lib = L
        return L


def last_error():
    return load().vs_last_error().decode("utf-8", "replace")


def init(device=None):
    """vs_init on LOCAL_RANK (or `device`). Raises BackendUnavailable without a B200."""
    global _inited
    L = load()
    if _inited and device is None:
        return L
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    rc = L.vs_init(int(device))
    if rc != VS_OK:
        raise BackendUnavailable(f"vs_init({device}) failed: {last_error()}")
    _inited = True
    return L
