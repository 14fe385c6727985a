"""ctypes binding of include/duet_b200.h.  Loading fails loudly: there is no fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libduet_b200.so")

DUET_OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_HASH_COLLISION, ERR_BAD_HP, ERR_ZERO_DIVISION, ERR_STATE = range(1, 8)
MEM_HOST, MEM_DEVICE, MEM_HOST_MAPPED = 0, 1, 2
SV_GT_MISSING = 1
CLS_FILTERED = 255
N_FEATURES = 6
N_COUNTERS = 8
FEATURE_NAMES = ("hapread_ratio", "sv_ratio", "hap1_avgsc", "hap2_avgsc", "totsc_ratio", "hap_avgsc_diff")
COUNTER_NAMES = ("n_sv", "n_kept", "n_emitted", "n_1|0", "n_0|1", "n_1|1", "n_joins", "n_hits")


INPUT_COLUMNS = ("read_off", "sv_off", "read_key", "read_tag", "sv_pos", "sv_svlen", "sv_svread", "sv_refread",
                 "sv_flags", "sv_group", "csr_off", "csr_key", "csr_chk")


class Thresholds(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("svlen_thres", "suppread_thres", "pc_max", "c0_sv_num_min",
                                         "c2_sv_num_min", "c2_hap0_min", "c1_ref_num_max", "_pad")] + \
               [(n, C.c_double) for n in ("c2_sv_ratio_min", "c2_avgsc_diff_max", "c1_one_ratio_lo",
                                          "c1_one_ratio_hi", "c1_hapread_ratio", "c1_avgsc_diff_max",
                                          "c1_two_ratio_a", "c1_two_ratio_b", "c1_two_ratio_c",
                                          "c1_totsc_ratio_max")]


class PhaseInput(C.Structure):
    _fields_ = [("mem", C.c_int32), ("n_shards", C.c_int32), ("n_reads", C.c_int64), ("n_svs", C.c_int64),
                ("n_joins", C.c_int64)] + \
               [(n, C.c_void_p) for n in INPUT_COLUMNS]


class PhaseOutput(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("gt", "ps", "cls", "hap1", "hap2", "hap0", "allhap", "totsc1", "totsc2",
                                          "features", "join_row", "order", "shard_counts")] + \
               [("n_emitted", C.c_int64)]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("device_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_ms", C.c_float * 8)]


class ClusterParams(C.Structure):
    _fields_ = [("max_distance", C.c_double), ("position_normalizer", C.c_double), ("partition_window", C.c_int32),
                ("_pad", C.c_int32)]


class ClusterInput(C.Structure):
    _fields_ = [("mem", C.c_int32), ("_pad", C.c_int32), ("n", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("contig", "type", "start", "end")]


KERNEL_NAMES = ("init", "build", "probe", "reduce", "tail", "oneps", "predict", "order")

# every symbol include/duet_b200.h declares
SYMBOLS = (
    "duet_abi_version", "duet_default_thresholds", "duet_create", "duet_destroy", "duet_last_error",
    "duet_set_thresholds", "duet_set_stream", "duet_phase_upload", "duet_phase_execute",
    "duet_phase_download", "duet_phase_run", "duet_host_alloc", "duet_host_free", "duet_sync",
    "duet_get_timings", "duet_launch_count", "duet_default_cluster_params", "duet_cluster_run", "duet_hash_names",
    "duet_pack_tags", "duet_hash_name_lists", "duet_debug_timers", "duet_decode_bam", "duet_set_decode_threads", "duet_free", "duet_decode_sam_text", "duet_count_lines",
    "duet_host_is_pinned", "duet_phase_input_layout", "duet_phase_output_layout", "duet_cluster_timings", "duet_decode_reads", "duet_rows_take", "duet_rows_free", "duet_decode_sv_vcf", "duet_svs_take", "duet_svs_free",
)
DECODE_ERR_INDEX, DECODE_ERR_VALUE, DECODE_ERR_ASCII, DECODE_ERR_RANGE, DECODE_ERR_CAPACITY, DECODE_ERR_FORMAT = 20, 21, 22, 23, 24, 25
DECODE_FALLBACK = 26
READS_SAM_TEXT, READS_BAM = 0, 1

_lib = None


class DuetError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"duet_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


def load() -> C.CDLL:
    """dlopen the in-tree CUDA library.  Raises if it has not been built -- the product
    never computes on the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:                                   # a fresh checkout: compile the library (needs nvcc), never fall back
            from . import build as _build
            _build.build()
        except Exception as e:
            raise ImportError(
                f"{LIB_PATH} is missing and could not be built ({e}): run `python -m duet_b200.build` "
                "(duet_b200 has no CPU fallback)") from e
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.duet_abi_version.restype = C.c_int
    lib.duet_default_thresholds.argtypes = [C.POINTER(Thresholds)]
    lib.duet_default_thresholds.restype = None
    lib.duet_create.argtypes = [C.c_int, C.POINTER(H)]
    lib.duet_destroy.argtypes = [H]
    lib.duet_destroy.restype = None
    lib.duet_last_error.argtypes = [H]
    lib.duet_last_error.restype = C.c_char_p
    lib.duet_set_thresholds.argtypes = [H, C.POINTER(Thresholds)]
    lib.duet_set_stream.argtypes = [H, C.c_void_p]
    lib.duet_phase_upload.argtypes = [H, C.POINTER(PhaseInput)]
    lib.duet_phase_execute.argtypes = [H, C.c_int]
    lib.duet_phase_download.argtypes = [H, C.POINTER(PhaseOutput)]
    lib.duet_phase_run.argtypes = [H, C.POINTER(PhaseInput), C.POINTER(PhaseOutput)]
    lib.duet_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    lib.duet_host_free.argtypes = [C.c_void_p]
    lib.duet_sync.argtypes = [H]
    lib.duet_get_timings.argtypes = [H, C.POINTER(Timings)]
    lib.duet_launch_count.argtypes = [H]
    lib.duet_launch_count.restype = C.c_int64
    lib.duet_default_cluster_params.argtypes = [C.POINTER(ClusterParams)]
    lib.duet_default_cluster_params.restype = None
    lib.duet_cluster_run.argtypes = [H, C.POINTER(ClusterInput), C.POINTER(ClusterParams), C.c_void_p,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    lib.duet_hash_names.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.duet_hash_names.restype = None
    lib.duet_decode_sam_text.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 2 + \
                                       [C.POINTER(C.c_int64)] * 3
    lib.duet_hash_name_lists.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.duet_hash_name_lists.restype = C.c_int64
    lib.duet_set_decode_threads.argtypes = [C.c_int]
    lib.duet_set_decode_threads.restype = C.c_int
    lib.duet_decode_bam.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)] + \
                                  [C.POINTER(C.c_int64)] * 3
    lib.duet_free.argtypes = [C.c_void_p]
    lib.duet_free.restype = None
    lib.duet_debug_timers.argtypes = [H, C.c_int, C.c_void_p]
    lib.duet_pack_tags.argtypes = [C.c_int64] + [C.c_void_p] * 5
    lib.duet_pack_tags.restype = None
    lib.duet_count_lines.argtypes = [C.c_void_p, C.c_int64]
    lib.duet_count_lines.restype = C.c_int64
    lib.duet_host_is_pinned.argtypes = [C.c_void_p]
    lib.duet_phase_input_layout.argtypes = [C.c_int64, C.c_int64, C.c_void_p]
    lib.duet_phase_input_layout.restype = C.c_int64
    lib.duet_phase_output_layout.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
    lib.duet_phase_output_layout.restype = C.c_int64
    lib.duet_cluster_timings.argtypes = [H, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    lib.duet_decode_reads.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p)] + [C.POINTER(C.c_int64)] * 3
    lib.duet_rows_take.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.duet_rows_free.argtypes = [C.c_void_p]
    lib.duet_rows_free.restype = None
    lib.duet_decode_sv_vcf.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.duet_svs_take.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.POINTER(C.c_int32)] + [C.c_void_p] * 4
    lib.duet_svs_free.argtypes = [C.c_void_p]
    lib.duet_svs_free.restype = None
    _lib = lib
    return lib
