"""ctypes binding of libvecgo_cuda.so (include/vecgo_cuda.h).

The shared library is the product; this module is the same thin binding a
Go cgo shim would be.  There is NO fallback: if the library
is missing the import fails, and every compute call fails with VecgoError
when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvecgo_cuda.so")

VG_OK = 0
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_FORMAT, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
METRIC_L2, METRIC_COSINE, METRIC_DOT, METRIC_HAMMING = 0, 1, 2, 3
CODEC_F32, CODEC_PQ, CODEC_OPQ, CODEC_SQ8, CODEC_BQ, CODEC_RABITQ, CODEC_INT4 = 0, 1, 2, 3, 4, 5, 6


class VecgoError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[vg_status {status}] {message}")
        self.status = status
        self.message = message


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m vecgo_b200.build` (nvcc, sm_100a). "
        "vecgo_b200 has no CPU or PyTorch fallback."
    )
lib = C.CDLL(LIB_PATH)

f32p, u8p, i8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int8)
i32p, u32p, i64p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)
i64, i32, u64, f32, vp, sz = C.c_int64, C.c_int32, C.c_uint64, C.c_float, C.c_void_p, C.c_size_t


class IndexDesc(C.Structure):
    _fields_ = [
        ("codec", i32), ("metric", i32), ("dim", i64), ("rows", i64), ("segment_id", C.c_uint32),
        ("reserved", C.c_uint32), ("row_base", u64),
        ("sq8_mins", f32p), ("sq8_inv_scales", f32p), ("int4_min", f32p), ("int4_diff", f32p),
        ("pq_m", i64), ("pq_k", i64), ("pq_codebooks", i8p), ("pq_scales", f32p), ("pq_offsets", f32p),
        ("opq_rotation", f32p), ("opq_block", i64), ("bq_threshold", f32), ("reserved2", C.c_uint32),
        ("num_partitions", i64), ("centroids", f32p), ("partition_offsets", u32p),
    ]


class SearchStats(C.Structure):
    _fields_ = [("queries", u64), ("distance_computations", u64), ("filter_queries", u64), ("second_chance_queries", u64),
                ("exact_rerun_queries", u64), ("threshold_pass_queries", u64)]


class FlatHeader(C.Structure):
    _fields_ = [("segment_id", u64), ("row_count", C.c_uint32), ("dim", C.c_uint32), ("metric", C.c_uint32),
                ("num_partitions", C.c_uint32), ("quantization_type", C.c_uint32), ("checksum", C.c_uint32)]


_SIGS = {
    "vg_device_count": [i32p],
    "vg_init": [i32],
    "vg_synchronize": [],
    "vg_dev_alloc": [C.POINTER(vp), sz],
    "vg_dev_free": [vp],
    "vg_memcpy_h2d": [vp, vp, sz],
    "vg_memcpy_d2h": [vp, vp, sz],
    "vg_set_stream": [u64],
    "vg_simd_dot": [f32p, f32p, i64, i64, f32p],
    "vg_simd_squared_l2": [f32p, f32p, i64, i64, f32p],
    "vg_simd_dot_batch": [f32p, i64, f32p, i64, i64, f32p],
    "vg_simd_squared_l2_batch": [f32p, i64, f32p, i64, i64, f32p],
    "vg_simd_sq8u_l2_batch": [f32p, i64, u8p, i64, i64, f32p, f32p, f32p],
    "vg_simd_int4_l2_batch": [f32p, i64, u8p, i64, i64, f32p, f32p, f32p],
    "vg_simd_pq_adc_lookup": [f32p, i64, u8p, i64, i64, f32p],
    "vg_simd_hamming": [u8p, i64, u8p, i64, i64, i32p],
    "vg_simd_squared_l2_bounded": [f32p, f32p, i64, i64, f32p, f32p, u8p],
    "vg_simd_scale": [f32p, i64, f32],
    "vg_normalize_l2": [f32p, i64, i64, u8p],
    "vg_sq8_train": [f32p, i64, i64, f32p, f32p, f32p, f32p],
    "vg_sq8_set_bounds": [f32p, f32p, i64, f32p, f32p],
    "vg_sq8_encode": [f32p, i64, i64, f32p, f32p, f32p, u8p],
    "vg_sq8_decode": [u8p, i64, i64, f32p, f32p, f32p],
    "vg_int4_train": [f32p, i64, i64, f32p, f32p],
    "vg_int4_encode": [f32p, i64, i64, f32p, f32p, u8p],
    "vg_int4_decode": [u8p, i64, i64, f32p, f32p, f32p],
    "vg_bq_train": [f32p, i64, i64, f32p],
    "vg_bq_encode": [f32p, i64, i64, f32, u8p],
    "vg_rabitq_encode": [f32p, i64, i64, u8p],
    "vg_pq_encode": [f32p, i64, i64, i64, i64, i8p, f32p, f32p, u8p],
    "vg_pq_decode": [u8p, i64, i64, i64, i64, i8p, f32p, f32p, f32p],
    "vg_pq_build_distance_table": [f32p, i64, i64, i64, i64, i8p, f32p, f32p, f32p],
    "vg_pq_train": [f32p, i64, i64, i64, i64, i64, u64, i8p, f32p, f32p, f32p],
    "vg_pq_train_dev": [vp, i64, i64, i64, i64, i64, u64, i8p, f32p, f32p, f32p],
    "vg_pq_train_range_dev": [vp, i64, i64, i64, i64, i64, u64, i64, i64, i8p, f32p, f32p, f32p],
    "vg_pq_assign_tc_stats": [u64p, u64p],
    "vg_opq_block_size": [i64, i64, i64p],
    "vg_opq_train": [f32p, i64, i64, i64, i64, i64, i64, u64, f32p, i8p, f32p, f32p],
    "vg_opq_rotate": [f32p, i64, i64, i64, f32p, i32, f32p],
    "vg_opq_procrustes": [f32p, i64, i64, f32p, f32p],
    "vg_minmax_dev": [vp, i64, i64, f32p, f32p],
    "vg_sq8_encode_dev": [vp, i64, i64, f32p, f32p, f32p, vp],
    "vg_int4_encode_dev": [vp, i64, i64, f32p, f32p, vp],
    "vg_rabitq_encode_dev": [vp, i64, i64, vp],
    "vg_pq_encode_dev": [vp, i64, i64, i64, i64, i8p, f32p, f32p, vp],
    "vg_kmeans_train": [f32p, i64, i64, i64, i32, i64, i64p, u64, f32p, i32p, i64p],
    "vg_kmeans_assign": [f32p, i64, i64, f32p, i64, i32, i32p],
    "vg_kmeans_find_closest": [f32p, i64, i64, f32p, i64, i64, i32, i32p],
    "vg_index_create": [C.POINTER(IndexDesc), u64p],
    "vg_index_create_on": [i32, C.POINTER(IndexDesc), u64p],
    "vg_index_set_stream": [u64, u64],
    "vg_index_device": [u64, i32p],
    "vg_index_search_dev_async": [u64, vp, i64, i64, i64, vp, vp, vp, vp, vp],
    "vg_index_search_resolve": [u64, vp, i64, i64, i64, vp, vp, vp, vp, vp, i64p],
    "vg_last_search_stats": [C.POINTER(SearchStats)],
    "vg_index_set_int4_score_mode": [u64, i32],
    "vg_int4_build_lookup_table": [f32p, f32p, i64, f32p],
    "vg_index_l2_bounded": [u64, f32p, i64, u32p, i64, f32p, i32, f32p, u8p],
    "vg_index_l2_bounded_dev": [u64, vp, i64, vp, i64, vp, i32, vp, vp],
    "vg_index_upload": [u64, i64, i64, vp, f32p],
    "vg_index_upload_dev": [u64, i64, i64, vp, vp],
    "vg_index_set_host_vectors": [u64, vp, i64],
    "vg_index_close": [u64],
    "vg_index_info": [u64, i64p, i64p, i64p, i64p],
    "vg_index_search": [u64, f32p, i64, i64, i64, u8p, u32p, f32p, i32p],
    "vg_index_search_dev": [u64, vp, i64, i64, i64, vp, vp, vp, vp],
    "vg_index_rerank": [u64, f32p, i64, u32p, i64, f32p],
    "vg_index_rerank_dev": [u64, vp, i64, vp, i64, vp],
    "vg_index_search_rerank": [u64, f32p, i64, i64, i64, u32p, f32p, i32p],
    "vg_index_score": [u64, f32p, i64, u32p, i64, f32p],
    "vg_index_score_dev": [u64, vp, i64, vp, i64, vp],
    "vg_flat_tc_enable": [i32],
    "vg_flat_tc_stats": [u64p, u64p],
    "vg_index_search_blocks": [u64, f32p, i64, i64, i64, u8p, u8p, u32p, f32p, i32p],
    "vg_index_search_blocks_dev": [u64, vp, i64, i64, i64, vp, u8p, vp, vp, vp],
    "vg_tile_skip_enable": [i32],
    "vg_quant_tc_i8_enable": [i32],
    "vg_quant_tc_i8_state": [i32p],
    "vg_ivf_grouped_enable": [i32],
    "vg_quant_tc_stats": [u64p, u64p],
    "vg_quant_tc_profile": [i32, C.POINTER(C.c_double), u64p],
    "vg_flat_tc_candidates": [u64, f32p, i64, i64, u32p, i32p, f32p, i64p],
    "vg_flat_open": [u8p, sz, i32, u64p],
    "vg_flat_open_on": [i32, u8p, sz, i32, u64p],
    "vg_flat_decode_header": [u8p, sz, C.POINTER(FlatHeader)],
    "vg_index_fetch_ids": [u64, u32p, i64, u64p],
    "vg_topk_merge_dev": [vp, vp, i64, i64, i64, i32, i64, vp, vp, vp],
    "vg_nccl_unique_id": [u8p],
    "vg_shard_group_create": [i32p, i32, u64p],
    "vg_shard_group_create_rank": [u8p, i32, i32, i32, u64p],
    "vg_shard_group_info": [u64, i32p, i32p, i32p],
    "vg_shard_group_destroy": [u64],
    "vg_shard_group_search": [u64, u64p, f32p, i64, i64, u32p, f32p, i32p],
    "vg_shard_group_search_rerank": [u64, u64p, f32p, i64, i64, i64, u32p, f32p, i32p],
    "vg_shard_group_search_dev": [u64, u64p, C.POINTER(vp), i64, i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
    "vg_shard_group_search_rerank_dev": [u64, u64p, C.POINTER(vp), i64, i64, i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
    "vg_topk_pack_dev": [vp, vp, i64, i32, vp],
    "vg_topk_merge_keys_dev": [vp, i64, i64, i64, i32, i64, vp, vp, vp],
    "vg_topk_merge": [u32p, f32p, i64, i64, i64, i32, i64, u32p, f32p, i32p],
}
for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = i32
    _fn.argtypes = _args
lib.vg_last_error.restype = C.c_char_p
lib.vg_last_error.argtypes = []
lib.vg_version.restype = C.c_char_p
lib.vg_version.argtypes = []
lib.vg_launch_count.restype = u64
lib.vg_launch_count.argtypes = []


def check(status: int) -> None:
    if status != VG_OK:
        raise VecgoError(status, lib.vg_last_error().decode("utf-8", "replace"))


def call(name: str, *args) -> None:
    check(getattr(lib, name)(*args))


def last_search_stats() -> dict:
    """Counters of the calling thread's last search call (vg_last_search_stats)."""
    st = SearchStats()
    call("vg_last_search_stats", C.byref(st))
    return {name: int(getattr(st, name)) for name, _ in SearchStats._fields_}


def launch_count() -> int:
    return int(lib.vg_launch_count())


# ------------------------------------------------------------------ numpy glue
def as_f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def as_u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def ptr(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)
