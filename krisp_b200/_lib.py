"""ctypes binding of libkrisp_b200.so (C ABI: include/krisp_b200.h).

There is no CPU fallback: if the CUDA library is missing or no sm_100a device is present, importing
callers get a loud error.  Nothing here touches ``oracle/``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkrisp_b200.so")

KB_OK, KB_EINVAL, KB_ECUDA, KB_ENOMEM, KB_EUNSUPPORTED, KB_EINTERNAL = 0, -1, -2, -3, -4, -5
_ERRNAMES = {KB_EINVAL: "KB_EINVAL", KB_ECUDA: "KB_ECUDA", KB_ENOMEM: "KB_ENOMEM",
             KB_EUNSUPPORTED: "KB_EUNSUPPORTED", KB_EINTERNAL: "KB_EINTERNAL"}

# every symbol include/krisp_b200.h declares (checked by tests/test_abi.py)
EXPORTS = ["kb_version", "kb_create", "kb_destroy", "kb_last_error", "kb_set_stream", "kb_configure",
           "kb_set_option", "kb_clear_sequences", "kb_reserve", "kb_add_sequence", "kb_add_fasta", "kb_fasta_flags",
           "kb_get_sequence", "kb_synchronize",
           "kb_search", "kb_shard_plan", "kb_shard_child_counts", "kb_shard_set_child_counts", "kb_sequence_buffer", "kb_shard_own_files", "kb_shard_ipc_export", "kb_shard_ipc_import", "kb_shard_ipc_close", "kb_shard_count",
           "kb_shard_scatter", "kb_shard_slab_plan", "kb_shard_slab_extract", "kb_shard_slab_send", "kb_shard_slab_buffers", "kb_shard_slab_own", "kb_shard_slab_level", "kb_shard_slab_finish", "kb_shard_extract", "kb_shard_recv_buffer", "kb_shard_search", "kb_result_get", "kb_result_rows",
           "kb_result_free", "kb_last_profile", "kb_last_counters", "kb_extract_sorted", "kb_table_get",
           "kb_table_free"]


class KrispB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {message}")
        self.code = code


class UnsupportedError(KrispB200Error):
    pass


class ResultView(ctypes.Structure):
    _fields_ = [("n_groups", ctypes.c_uint64), ("n_records", ctypes.c_uint64), ("n_run_records", ctypes.c_uint64),
                ("L", ctypes.c_int32), ("D", ctypes.c_int32), ("R", ctypes.c_int32),
                ("flank_words", ctypes.c_int32), ("mask_words", ctypes.c_int32), ("record_words", ctypes.c_int32),
                ("reserved", ctypes.c_int32),
                ("flank", ctypes.POINTER(ctypes.c_uint64)), ("in_mask", ctypes.POINTER(ctypes.c_uint32)),
                ("out_mask", ctypes.POINTER(ctypes.c_uint32)), ("group_size", ctypes.POINTER(ctypes.c_uint32)),
                ("run_offset", ctypes.POINTER(ctypes.c_uint64)), ("records", ctypes.POINTER(ctypes.c_uint64)),
                ("stats", ctypes.c_uint64 * 4)]


_lib = None


def load():
    """Load the shared library (built by ``python -m krisp_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m krisp_b200.build` "
                          "(nvcc, sm_100a).  krisp_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i, u64, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_longlong
    pvp = ctypes.POINTER(ctypes.c_void_p)
    L.kb_version.restype = ctypes.c_char_p
    L.kb_version.argtypes = []
    L.kb_create.argtypes = [i, pvp]
    L.kb_destroy.argtypes = [vp]
    L.kb_destroy.restype = None
    L.kb_last_error.argtypes = [vp]
    L.kb_last_error.restype = ctypes.c_char_p
    L.kb_set_stream.argtypes = [vp, vp]
    L.kb_configure.argtypes = [vp, i, i, i, i, i, vp]
    L.kb_set_option.argtypes = [vp, ctypes.c_char_p, ll]
    L.kb_clear_sequences.argtypes = [vp]
    L.kb_reserve.argtypes = [vp, u64]
    L.kb_add_sequence.argtypes = [vp, i, vp, u64, i]
    L.kb_add_fasta.argtypes = [vp, i, vp, u64]
    L.kb_fasta_flags.argtypes = [vp, ctypes.POINTER(ctypes.c_uint)]
    L.kb_get_sequence.argtypes = [vp, i, vp, u64, ctypes.POINTER(u64)]
    L.kb_synchronize.argtypes = [vp]
    L.kb_search.argtypes = [vp, pvp]
    L.kb_shard_plan.argtypes = [vp, i, i, u64, ctypes.POINTER(i)]
    L.kb_shard_child_counts.argtypes = [vp, vp, u64, ctypes.POINTER(u64)]
    L.kb_shard_set_child_counts.argtypes = [vp, vp, u64]
    L.kb_shard_ipc_close.argtypes = [vp]
    L.kb_sequence_buffer.argtypes = [vp, pvp, ctypes.POINTER(u64)]
    L.kb_shard_own_files.argtypes = [vp, i, i]
    L.kb_shard_ipc_export.argtypes = [vp, u64, ctypes.c_char_p]
    L.kb_shard_ipc_import.argtypes = [vp, i, ctypes.c_char_p]
    L.kb_shard_count.argtypes = [vp, ctypes.POINTER(u64)]
    L.kb_shard_scatter.argtypes = [vp, ctypes.POINTER(u64)]
    L.kb_shard_slab_plan.argtypes = [vp, i, i, u64, u64, ctypes.POINTER(i), ctypes.POINTER(u64), ctypes.POINTER(i)]
    L.kb_shard_slab_extract.argtypes = [vp, pvp]
    L.kb_shard_slab_send.argtypes = [vp, i, i, i, i, vp]
    L.kb_shard_slab_buffers.argtypes = [vp, pvp, pvp, ctypes.POINTER(u64), ctypes.POINTER(i)]
    L.kb_shard_slab_own.argtypes = [vp, vp]
    L.kb_shard_slab_level.argtypes = [vp, vp, i, i]
    L.kb_shard_slab_finish.argtypes = [vp, ctypes.POINTER(i), pvp]
    L.kb_shard_extract.argtypes = [vp, pvp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.kb_shard_recv_buffer.argtypes = [vp, u64, pvp]
    L.kb_shard_search.argtypes = [vp, u64, ctypes.POINTER(u64), pvp]
    L.kb_result_get.argtypes = [vp, ctypes.POINTER(ResultView)]
    L.kb_result_rows.argtypes = [vp, pvp, ctypes.POINTER(u64), ctypes.POINTER(i)]
    L.kb_result_free.argtypes = [vp]
    L.kb_result_free.restype = None
    L.kb_last_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float), i]
    L.kb_last_counters.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(i)]
    L.kb_extract_sorted.argtypes = [vp, i, pvp]
    L.kb_table_get.argtypes = [vp, ctypes.POINTER(ctypes.POINTER(u64)), ctypes.POINTER(u64), ctypes.POINTER(i)]
    L.kb_table_free.argtypes = [vp]
    L.kb_table_free.restype = None
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int and name not in ("kb_version", "kb_last_error"):
            fn.restype = i
    _lib = L
    return L


def check(ctx, rc):
    if rc == KB_OK:
        return
    msg = load().kb_last_error(ctx)
    msg = msg.decode(errors="replace") if msg else ""
    raise (UnsupportedError if rc == KB_EUNSUPPORTED else KrispB200Error)(rc, msg)
