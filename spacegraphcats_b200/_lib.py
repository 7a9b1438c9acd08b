"""ctypes binding of ``libsgc_b200.so`` (C-ABI: ``include/sgc_b200.h``).

The library holds every kernel of the hot path; there is no CPU fallback.  Loading
fails loudly when the shared object has not been built (``python -c 'import
__graft_entry__ as g; g.build()'`` or ``make -C spacegraphcats_b200/csrc``), and
:func:`check` raises :class:`SgcError` with the library's own message on any error.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgc_b200.so")

SGC_MISS = 0xFFFFFFFF
SGC_MAX_ID = 0xFFFFFFFD
SGC_MAX_K = 255
SGC_OK, SGC_ERR_CUDA, SGC_ERR_ARG, SGC_ERR_CAPACITY, SGC_ERR_NOMEM = 0, -1, -2, -3, -4


class SgcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsgc_b200: %s (status %d)" % (msg, code))
        self.code = code


class SgcCapacityError(SgcError):
    pass


_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int
_u32 = ctypes.c_uint32
_pi64 = ctypes.POINTER(ctypes.c_int64)
_pu32 = ctypes.POINTER(ctypes.c_uint32)

# name -> (restype, argtypes); every symbol include/sgc_b200.h declares
PROTOTYPES = {
    "sgc_last_error": (ctypes.c_char_p, []),
    "sgc_abi_version": (_i32, []),
    "sgc_ctx_create": (_i32, [_i32, _vp, ctypes.POINTER(_vp)]),
    "sgc_ctx_destroy": (None, [_vp]),
    "sgc_ctx_sync": (_i32, [_vp]),
    "sgc_ctx_stream": (_vp, [_vp]),
    "sgc_ctx_launches": (_i64, [_vp]),
    "sgc_hash_batch": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _pi64]),
    "sgc_minhash_batch": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, ctypes.c_uint64, _vp, _i64, _vp, _vp, _vp, _pi64]),
    "sgc_table_create": (_i32, [_vp, _i64, ctypes.c_float, ctypes.POINTER(_vp)]),
    "sgc_table_free": (None, [_vp]),
    "sgc_table_insert_pairs": (_i32, [_vp, _vp, _vp, _i64]),
    "sgc_table_insert_sequences": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _u32, _vp]),
    "sgc_table_stats": (_i32, [_vp, _pi64, _pi64, _pu32]),
    "sgc_table_export": (_i32, [_vp, _vp, _vp, _i64, _pi64]),
    "sgc_table_bytes": (_i64, [_vp]),
    "sgc_table_set_home_shift": (_i32, [_vp, _i32]),
    "sgc_table_enable_side": (_i32, [_vp]),
    "sgc_table_side_bytes": (_i64, [_vp]),
    "sgc_lookup": (_i32, [_vp, _vp, _i64, _vp]),
    "sgc_count_matches": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _pi64, _pi64]),
    "sgc_query_count_batch": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _pi64, _pi64, _pi64]),
    "sgc_lookup_sequences": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _i64, _vp, _pi64, _pi64]),
    "sgc_label_reads": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _pi64, _pi64, _pi64]),
    "sgc_label_reads_launch": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    "sgc_label_reads_finish": (_i32, [_vp, _pi64, _pi64, _pi64]),
    "sgc_packed_bytes": (_i64, [_vp, _i32, _i64]),
    "sgc_offsets_to_counts8": (_i32, [_vp, _vp, _i64, _vp]),
    "sgc_counts8_to_offsets": (_i32, [_vp, _i64, _i32, _vp]),
    "sgc_label_reads_packed_launch": (_i32, [_vp, _vp, _i64, _vp, _i32, _i64, _i32, _vp, _vp, _i64]),
    "sgc_label_reads_packed": (_i32, [_vp, _vp, _i64, _vp, _i32, _i64, _i32, _vp, _vp, _i64, _pi64, _pi64, _pi64]),
    "sgc_pack_reads_device": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _pi64]),
    "sgc_pack_reads": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _pi64]),
    "sgc_count_doms": (_i32, [_vp, _vp, _vp, _vp, _vp, _pi64, _i32, _vp, _i64]),
    "sgc_hash_partition": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _pi64]),
    "sgc_partition_pairs": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _pi64]),
    "sgc_labels_from_ids": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _pi64, _pi64]),
    "sgc_unpermute_u32": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "sgc_partition_hashes": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _pi64]),
    "sgc_partition_hashes_p2p": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _pi64]),
    "sgc_lookup_regions_p2p": (_i32, [_vp, _vp, _pi64, _i64, _i32, _i32, _vp, _i32]),
    "sgc_label_reads_partitioned": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _u32, _vp, _vp, _vp, _vp, _vp, _i64,
                                            _vp, _i64, _vp, _vp, _i64, _vp, _pi64, _pi64, _pi64, ctypes.POINTER(ctypes.c_int32)]),
    "sgc_hash_positions": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _pi64]),
    "sgc_table_insert_triples": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _i64, _pi64]),
    "sgc_store_build": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "sgc_store_apply_marks": (_i32, [_vp, _vp, _i64, _vp, _i64]),
    "sgc_table_create_shared": (_i32, [_vp, _i64, ctypes.c_float, ctypes.POINTER(_vp)]),
    "sgc_table_export_fd": (_i32, [_vp, ctypes.POINTER(_vp), _pi64, ctypes.POINTER(ctypes.c_int)]),
    "sgc_peer_import_fd": (_i32, [_vp, _i32, _i64, ctypes.POINTER(_vp)]),
    "sgc_peer_unmap": (_i32, [_vp, _vp, _i64]),
    "sgc_table_set_peers": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _vp, _i64]),
    "sgc_filter_insert": (_i32, [_vp, _vp, _i64, _vp, _i32]),
    "sgc_table_set_filter": (_i32, [_vp, _vp, _i32, _i32]),
    "sgc_peer_alloc": (_i32, [_vp, _i64, ctypes.POINTER(_vp), _vp]),
    "sgc_peer_open": (_i32, [_vp, _vp, ctypes.POINTER(_vp)]),
    "sgc_peer_close": (_i32, [_vp, _vp]),
    "sgc_peer_free": (_i32, [_vp, _vp]),
    "sgc_labels_from_replies": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _pi64, _pi64, _pi64]),
    "sgc_ctx_set_timing": (_i32, [_vp, _i32]),
    "sgc_ctx_tile_ms": (_i32, [_vp, ctypes.POINTER(ctypes.c_double), _pi64]),
    "sgc_synth_genome": (_i32, [_vp, _vp, _i64, ctypes.c_uint64]),
    "sgc_synth_reads": (_i32, [_vp, _vp, _i64, _i64, _i32, ctypes.c_uint64, _u32, _i32, _vp, _vp]),
    "sgc_synth_unitigs": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp]),
    "sgc_mphf_build": (_i32, [_vp, _vp, _i64, ctypes.c_double, _i32, ctypes.POINTER(_vp)]),
    "sgc_mphf_load": (_i32, [_vp, _vp, _i64, ctypes.POINTER(_vp)]),
    "sgc_mphf_free": (_i32, [_vp]),
    "sgc_mphf_info": (_i32, [_vp, _pi64, ctypes.POINTER(ctypes.c_double)]),
    "sgc_mphf_lookup": (_i32, [_vp, _vp, _i64, _vp]),
    "sgc_mphf_order": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "sgc_mphf_dump_bytes": (_i32, [_vp, _pi64]),
    "sgc_mphf_dump": (_i32, [_vp, _vp, _i64]),
    "sgc_bgzf_index": (_i32, [_vp, _i64, _vp, _vp, _i64, _pi64, _pi64]),
    "sgc_bgzf_inflate": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _i64]),
    "sgc_reads_count": (_i32, [_vp, _i64, _i32, _vp]),
    "sgc_reads_cut": (_i32, [_vp, _i64, _i32, _pi64]),
    "sgc_reads_scan": (_i32, [_vp, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "sgc_reads_pair_flags": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "sgc_reads_rows": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _pi64]),
}

STORE_PAD_WORDS = 8  # SGC_STORE_PAD_WORDS
BRK_PAD_BITS = 32  # SGC_BRK_PAD_BITS

_lib = None


def lib():
    """The loaded library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "spacegraphcats_b200: %s is missing -- build it with __graft_entry__.build(); "
                "this package has no CPU fallback" % LIB_PATH
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status == SGC_OK:
        return
    msg = lib().sgc_last_error().decode("utf-8", "replace")
    if status == SGC_ERR_CAPACITY:
        raise SgcCapacityError(status, msg)
    raise SgcError(status, msg)
