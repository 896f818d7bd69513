"""ctypes binding of libkm_b200.so (include/km_b200.h).  Fails loudly when the library has not
been built: there is no Python or CPU fallback for the hot path."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# KM_B200_LIB: an alternative build of the same library (tuning experiments: tools/build_variant.py)
LIB_PATH = os.environ.get("KM_B200_LIB") or os.path.join(HERE, "libkm_b200.so")

u64, u32, i64, i32, u8 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint8
vp, cp, ci, cd, cf = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_double, ctypes.c_float


class KmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("km_b200 error %d: %s" % (code, msg))
        self.code = code


class TableInfo(ctypes.Structure):
    _fields_ = [("k", i32), ("canonical", i32), ("device", i32), ("reserved", i32),
                ("n_keys", u64), ("n_buckets", u64), ("bytes", u64)]


class FindParams(ctypes.Structure):
    _fields_ = [("ratio", cd), ("count", i64), ("steps", i32), ("branchs", i32), ("nodes", i32),
                ("extra_nodes", i32), ("flags", i32), ("reserved", i32)]


ROW_DTYPE = np.dtype([
    ("target", "<i4"), ("kind", "<i4"), ("type", "<i4"), ("name_start", "<i4"), ("name_end", "<i4"),
    ("path_id", "<i4"), ("var_begin", "<i4"), ("var_end", "<i4"), ("ref_begin", "<i4"), ("ref_end", "<i4"),
    ("del_begin", "<i4"), ("del_len", "<i4"), ("ins_begin", "<i4"), ("ins_len", "<i4"), ("start_off", "<i4"),
    ("cluster_id", "<i4"), ("cluster_n", "<i4"), ("n_iter", "<i4"), ("min_cov", "<i8"),
    ("rvaf", "<f8"), ("expr", "<f8"), ("ref_rvaf", "<f8"), ("ref_expr", "<f8")], align=True)
assert ROW_DTYPE.itemsize == 112, ROW_DTYPE.itemsize


class ResultView(ctypes.Structure):
    _fields_ = [("n_targets", i32), ("n_paths", i32), ("n_rows", i32), ("k", i32),
                ("status", vp), ("n_nodes", vp), ("node_off", vp), ("node_kmer", vp), ("node_count", vp),
                ("path_first", vp), ("path_count", vp), ("path_off", vp), ("path_len", vp), ("path_pool", vp),
                ("row_first", vp), ("row_count", vp), ("rows", vp), ("lookups", vp),
                ("ms_h2d", cf), ("ms_walk", cf), ("ms_graph", cf), ("ms_d2h", cf), ("ms_total", cf),
                ("n_launches", i32), ("n_retries", i32), ("has_graph", i32), ("reserved", i32),
                ("bytes_h2d", u64), ("bytes_d2h", u64)]


_LIB = None


def lib():
    """Load the CUDA library.  Raises if it is missing -- build it with
    ``python -m km_b200.build`` (or __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("km_b200: %s is missing; the CUDA library must be built (python -m km_b200.build). "
                          "There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    L.km_last_error.restype = cp
    L.km_version.restype = cp
    L.km_device_count.restype = ci
    L.km_table_open_jf.argtypes = [cp, ci, P(vp)]
    L.km_table_create.argtypes = [ci, ci, ci, u64, P(vp)]
    L.km_table_create_layout.argtypes = [ci, ci, ci, u64, ci, P(vp)]
    L.km_table_create_layout.restype = ci
    L.km_table_export.argtypes = [vp, vp, vp, u64, P(u64)]
    L.km_table_write_jf.argtypes = [vp, cp, ctypes.c_uint32]
    L.km_table_insert.argtypes = [vp, vp, vp, u64, ci]
    L.km_table_build_synthetic.argtypes = [vp, u64, u64]
    L.km_table_count_reads.argtypes = [vp, cp, vp, i64]
    L.km_table_count_file.argtypes = [vp, cp, ci, P(u64), P(u64)]
    L.km_table_count_text.argtypes = [vp, vp, vp, u64, ci]
    L.km_table_link.argtypes = [vp]
    L.km_table_link.restype = ci
    L.km_table_set_routing.argtypes = [vp, ci]
    L.km_table_recount.argtypes = [vp, P(u64)]
    L.km_shard_owner_device.argtypes = [vp, vp, u64, vp, vp]
    L.km_route_partition.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.km_route_unpermute.argtypes = [vp, vp, vp, u64, vp, vp]
    for _n in ("km_table_link", "km_table_count_text", "km_table_set_routing", "km_table_recount", "km_shard_owner_device", "km_route_partition",
               "km_route_unpermute"):
        getattr(L, _n).restype = ci
    L.km_table_drop_below.argtypes = [vp, u32, P(u64)]
    L.km_table_get_info.argtypes = [vp, P(TableInfo)]
    L.km_table_close.argtypes = [vp]
    L.km_table_close.restype = None
    L.km_query_batch.argtypes = [vp, vp, u64, vp]
    L.km_query_batch_device.argtypes = [vp, vp, u64, vp, vp]
    L.km_query_ascii.argtypes = [vp, cp, u64, vp]
    L.km_get_child_batch.argtypes = [vp, vp, u64, ci, cd, i64, vp, vp]
    L.km_find_batch.argtypes = [vp, cp, vp, i32, P(FindParams), P(vp)]
    L.km_result_get.argtypes = [vp, P(ResultView)]
    L.km_result_free.argtypes = [vp]
    L.km_result_free.restype = None
    L.km_result_format_target.argtypes = [vp, i32, cp, cp, vp, i64]
    L.km_result_format_target.restype = i64
    L.km_result_format_all.argtypes = [vp, cp, cp, vp, i32, vp, i64]
    L.km_result_format_all.restype = i64
    L.km_table_create_shard.argtypes = [ci, ci, ci, u64, ci, ci, P(vp)]
    L.km_table_shard_export_fd.argtypes = [vp, P(ci)]
    L.km_table_shard_attach_fd.argtypes = [vp, ci, ci]
    L.km_shard_owner.argtypes = [vp, u64, ci, ci, ci, vp]
    for _n in ("km_table_create_shard", "km_table_shard_export_fd", "km_table_shard_attach_fd", "km_shard_owner"):
        getattr(L, _n).restype = ci
    L.km_debug_format_fixed.argtypes = [ctypes.c_double, ci, cp]
    L.km_debug_format_fixed.restype = ci
    L.km_debug_nat_cmp.argtypes = [cp, cp]
    L.km_debug_nat_cmp.restype = ci
    L.km_find_text.argtypes = [vp, cp, vp, i32, P(FindParams), cp, cp, vp, i32, P(vp)]
    L.km_find_text.restype = ci
    L.km_result_text.argtypes = [vp, cp, cp, vp, i32, P(vp)]
    L.km_result_text.restype = i64
    L.km_find_plan_create.argtypes = [vp, cp, vp, i32, P(FindParams), P(vp)]
    L.km_find_plan_launch.argtypes = [vp, vp]
    L.km_find_plan_fetch.argtypes = [vp, ci, P(vp)]
    L.km_find_plan_last_ms.argtypes = [vp, P(cf), P(cf)]
    L.km_find_plan_kernel_ms.argtypes = [vp, P(cf)]
    L.km_find_plan_kernel_ms.restype = ci
    L.km_find_plan_free.argtypes = [vp]
    L.km_find_plan_free.restype = None
    L.km_bench_random_gather.argtypes = [ci, u64, u64, ci, P(cf)]
    L.km_bench_count.argtypes = [vp, u64, ci, u64, u64, ci, P(cf), P(u64)]
    L.km_bench_count.restype = ci
    L.km_bench_make_queries.argtypes = [vp, vp, u64, u64, u64, u64, vp]
    L.km_bench_make_queries.restype = ci
    L.km_bench_lookup.argtypes = [vp, u64, u64, u64, u64, ci, P(cf), P(cf), P(u64)]
    for name in ("km_table_open_jf", "km_table_create", "km_table_insert", "km_table_build_synthetic",
                 "km_table_count_reads", "km_table_count_file", "km_table_drop_below", "km_table_get_info", "km_query_batch",
                 "km_query_batch_device", "km_query_ascii", "km_get_child_batch", "km_find_batch",
                 "km_result_get", "km_bench_random_gather", "km_bench_lookup", "km_find_plan_create",
                 "km_find_plan_launch", "km_find_plan_fetch", "km_find_plan_last_ms"):
        getattr(L, name).restype = ci
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise KmError(rc, lib().km_last_error().decode("utf-8", "replace"))


EXPORTS = ["km_table_link", "km_last_error", "km_device_count", "km_version", "km_table_open_jf", "km_table_create",
           "km_table_insert", "km_table_build_synthetic", "km_table_count_reads", "km_table_count_file", "km_table_drop_below",
           "km_table_get_info", "km_table_close", "km_query_batch", "km_query_batch_device", "km_query_ascii",
           "km_get_child_batch", "km_find_batch", "km_result_get", "km_result_free", "km_result_format_target", "km_result_format_all", "km_result_text", "km_find_text", "km_table_create_layout", "km_table_export", "km_table_write_jf", "km_table_create_shard", "km_table_shard_export_fd", "km_table_shard_attach_fd", "km_shard_owner", "km_debug_format_fixed", "km_debug_nat_cmp",
           "km_table_count_text", "km_table_set_routing", "km_table_recount", "km_shard_owner_device", "km_route_partition", "km_route_unpermute",
           "km_find_plan_create", "km_find_plan_launch", "km_find_plan_fetch", "km_find_plan_free", "km_find_plan_last_ms", "km_find_plan_kernel_ms",
           "km_bench_random_gather", "km_bench_lookup", "km_bench_make_queries", "km_bench_count", "km_debug_walk_cycles", "km_debug_phase_cycles", "km_debug_target_cycles", "km_debug_timeline"]
