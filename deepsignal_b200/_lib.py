"""ctypes binding of libdsb200.so (include/dsb200.h).  Fails loudly: there is no
Python / PyTorch / CPU fallback for any function of the hot path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSB_LIB") or os.path.join(_HERE, "libdsb200.so")   # DSB_LIB: another build of the same ABI (A/B timing)

DSB_OK = 0
DSB_ERR_CAPACITY = -5
DSB_ERR_NONFINITE = -8
ABI_VERSION = 2
FLAG_CNN, FLAG_RNN, FLAG_BASE = 1, 2, 4
PREC_FP32, PREC_BF16, PREC_FP16, PREC_FP16X2, PREC_FP16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16, "fp16x2": PREC_FP16X2, "fp16x3": PREC_FP16X3,
              0: PREC_FP32, 1: PREC_BF16, 2: PREC_FP16, 3: PREC_FP16X2, 4: PREC_FP16X3}
# the tensor-core mode that meets BASELINE's 2e-3 / 99.9 % tolerance on spread-logit checkpoints
DEFAULT_PRECISION = "fp16x2"

EXPORTS = [
    "dsb_abi_version", "dsb_create", "dsb_destroy", "dsb_last_error", "dsb_load_weights",
    "dsb_set_max_chunk", "dsb_get_plan_options", "dsb_set_plan_options", "dsb_workspace_bytes", "dsb_forward_dev",
    "dsb_forward_host", "dsb_nonfinite_count",
    "dsb_set_profiling", "dsb_get_profile", "dsb_launch_count", "dsb_extract_dev",
    "dsb_extract_host", "dsb_extract_max_sites", "dsb_extract_call_dev", "dsb_site_frequency_host",
    "dsb_tsv_parse_features", "dsb_tsv_format_calls",
]


# classes of dsb_get_profile (common.cuh KClass): forward kernels, then the extraction chain
PROFILE_CLASSES = ("gemm", "lstm", "elementwise", "incept_a", "incept_b", "front",
                   "x_read_norm", "x_sites", "x_lut", "x_unused", "x_site_feature")


class DsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libdsb200 error {}: {}".format(code, msg))
        self.code = code


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class PlanOptions(C.Structure):
    _fields_ = [("fused_blocks", C.c_int32), ("fused_front", C.c_int32), ("lstm_seq", C.c_int32),
                ("ablate_a", C.c_int32), ("ablate_b", C.c_int32), ("ablate_gemm", C.c_int32), ("ablate_lstm", C.c_int32),
                ("lstm_sms", C.c_int32)]


class Reads(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("raw", C.c_void_p), ("raw_off", C.c_void_p), ("scaling", C.c_void_p),
                ("offset", C.c_void_p), ("ev_start", C.c_void_p), ("ev_len", C.c_void_p), ("ev_base", C.c_void_p),
                ("ev_off", C.c_void_p), ("chrom_start", C.c_void_p), ("strand_minus", C.c_void_p),
                ("chrom_len", C.c_void_p)]


class ExtractParams(C.Structure):
    _fields_ = [("kmer_len", C.c_int32), ("signal_len", C.c_int32), ("normalize_method", C.c_int32),
                ("mod_loc", C.c_int32), ("n_motifs", C.c_int32), ("motif_len", C.c_int32), ("motifs", C.c_char_p)]


SITE_FIELDS = ["site_read", "site_loc", "pos", "pos_in_strand", "kmer", "means", "stds", "lens", "signals",
               "means64", "stds64", "lens_i32", "signals64", "window"]


class Sites(C.Structure):
    _fields_ = [("capacity", C.c_int64)] + [(f, C.c_void_p) for f in SITE_FIELDS]


_lib = None


def load():
    """Load the shared library (built by ``__graft_entry__.build()`` / csrc/build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("{} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(deepsignal_b200/csrc/build.sh). There is no fallback path.".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.dsb_abi_version.restype = C.c_int
    lib.dsb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.dsb_destroy.argtypes = [vp]
    lib.dsb_last_error.argtypes = [vp]
    lib.dsb_last_error.restype = C.c_char_p
    lib.dsb_load_weights.argtypes = [vp, C.POINTER(Tensor), i32, C.c_uint32, i32, i32]
    lib.dsb_set_max_chunk.argtypes = [vp, i64]
    lib.dsb_get_plan_options.argtypes = [vp, C.POINTER(PlanOptions)]
    lib.dsb_set_plan_options.argtypes = [vp, C.POINTER(PlanOptions)]
    lib.dsb_nonfinite_count.argtypes = [vp, C.POINTER(i64)]
    lib.dsb_workspace_bytes.argtypes = [vp, i64, i32]
    lib.dsb_workspace_bytes.restype = sz
    lib.dsb_forward_dev.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.dsb_forward_host.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.dsb_set_profiling.argtypes = [vp, i32]
    lib.dsb_get_profile.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]
    lib.dsb_launch_count.argtypes = [vp]
    lib.dsb_launch_count.restype = i64
    lib.dsb_extract_dev.argtypes = [vp, C.POINTER(Reads), C.POINTER(ExtractParams), C.POINTER(Sites),
                                    C.POINTER(i64), vp, vp, vp]
    lib.dsb_extract_host.argtypes = [vp, C.POINTER(Reads), C.POINTER(ExtractParams), C.POINTER(Sites),
                                     C.POINTER(i64), vp, vp]
    lib.dsb_extract_max_sites.argtypes = [C.POINTER(Reads)]
    lib.dsb_extract_max_sites.restype = i64
    lib.dsb_extract_call_dev.argtypes = [vp, i32, C.POINTER(Reads), C.POINTER(ExtractParams), C.POINTER(Sites),
                                         C.POINTER(i64), vp, vp, vp]
    lib.dsb_site_frequency_host.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, i64, C.POINTER(i64), vp, vp, vp, vp, vp]
    lib.dsb_tsv_parse_features.argtypes = [vp, i64, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64), i32]
    lib.dsb_tsv_parse_features.restype = i64
    lib.dsb_tsv_format_calls.argtypes = [vp, vp, vp, i32, vp, vp, i64, vp, i64]
    lib.dsb_tsv_format_calls.restype = i64
    for name in EXPORTS:
        if name not in ("dsb_tsv_parse_features", "dsb_tsv_format_calls", "dsb_last_error", "dsb_workspace_bytes", "dsb_launch_count", "dsb_extract_max_sites",
                        "dsb_abi_version"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
