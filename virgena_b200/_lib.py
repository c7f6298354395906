"""ctypes binding of libvirgena_gpu.so (include/virgena_gpu.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a). There is no CPU
fallback: if the shared object is missing, or no CUDA device is usable, this raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvirgena_gpu.so")

VG_OK = 0
STATUS = {0: "VG_OK", -1: "VG_ERR_INVALID", -2: "VG_ERR_CUDA", -3: "VG_ERR_STATE", -4: "VG_ERR_CAPACITY",
          -5: "VG_ERR_ALPHABET", -6: "VG_ERR_INDEX", -7: "VG_ERR_OVERFLOW"}

CLS_CONCORDANT_FR, CLS_CONCORDANT_RF, CLS_DISCORDANT, CLS_LEFT, CLS_RIGHT, CLS_UNMAPPED = range(6)

MATE_DTYPE = np.dtype(
    [
        ("present", "<i4"), ("start", "<i4"), ("end", "<i4"), ("count", "<i4"),
        ("reverse", "<i4"), ("fragment_id", "<i4"),
        ("score", "<i4"), ("start1", "<i4"), ("end1", "<i4"), ("start2", "<i4"), ("end2", "<i4"),
        ("identity", "<i4"), ("length", "<i4"), ("pad_", "<i4"),
        ("seq1_offset", "<i8"),
    ],
    align=True,
)
PAIR_DTYPE = np.dtype([("cls", "<i4"), ("pad_", "<i4"), ("m", MATE_DTYPE, (2,))], align=True)
assert MATE_DTYPE.itemsize == 64 and PAIR_DTYPE.itemsize == 136
BAM_DTYPE = np.dtype([("flag", "<i4"), ("ref_index", "<i4"), ("pos", "<i4"), ("mapq", "<i4"), ("mate_ref_index", "<i4"),
                      ("mate_pos", "<i4"), ("xn", "<i4"), ("n_cigar", "<i4"), ("cigar_offset", "<i8")], align=True)
assert BAM_DTYPE.itemsize == 40


class VgParams(C.Structure):
    _fields_ = [("K", C.c_int32), ("low_coef", C.c_float), ("high_coef", C.c_float),
                ("cutoffs", C.c_void_p), ("n_cutoffs", C.c_int32), ("insert_len", C.c_int32),
                ("match", C.c_int32), ("mismatch", C.c_int32), ("gop", C.c_int32), ("gep", C.c_int32)]


class VgMapStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("total_pairs", "both_exist", "reads_total", "total_mapped", "mapped_forward",
                                         "mapped_reverse", "both_mapped", "total_score", "total_concordant")]

    def as_array(self):
        return np.array([getattr(self, n) for n, _ in self._fields_], np.int64)


# every symbol include/virgena_gpu.h declares, with its signature
_SIGS = {
    "vg_abi_version": (C.c_int, []),
    "vg_device_count": (C.c_int, []),
    "vg_ctx_create": (C.c_int, [C.POINTER(VgParams), C.c_int, C.POINTER(C.c_void_p)]),
    "vg_ctx_destroy": (None, [C.c_void_p]),
    "vg_last_error": (C.c_char_p, [C.c_void_p]),
    "vg_last_required": (C.c_int64, [C.c_void_p]),
    "vg_set_reference": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vg_set_msa_index": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vg_upload_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int64]),
    "vg_upload_reads_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64]),
    "vg_pack_reads": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vg_get_results_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vg_map_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(VgMapStats)]),
    "vg_map_pairs_msa": (C.c_int, [C.c_void_p, C.POINTER(VgMapStats)]),
    "vg_result_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vg_get_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "vg_pileup": (C.c_int, [C.c_void_p, C.c_int32]),
    "vg_pileup_counts_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "vg_pileup_counts_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "vg_consensus": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vg_sw_align_batch": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vg_sw_max_score_batch": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]),
    "vg_random_model_counts": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_int64,
                                         C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "vg_seed_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                C.c_void_p]),
    "vg_last_timings": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vg_seed_timings": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vg_best_window_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vg_map_to_regions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "vg_pileup_alignments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p]),
    "vg_set_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]),
    "vg_identity_counting": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vg_bam_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "vg_comm_unique_id": (C.c_int, [C.c_void_p]),
    "vg_comm_init_rank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "vg_comm_init_all": (C.c_int, [C.c_void_p, C.c_int]),
    "vg_comm_destroy": (None, [C.c_void_p]),
    "vg_comm_size": (C.c_int, [C.c_void_p]),
    "vg_pileup_allreduce": (C.c_int, [C.c_void_p]),
    "vg_last_allreduce_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vg_multi_create": (C.c_int, [C.POINTER(VgParams), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "vg_multi_destroy": (None, [C.c_void_p]),
    "vg_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "vg_multi_last_required": (C.c_int64, [C.c_void_p]),
    "vg_multi_size": (C.c_int32, [C.c_void_p]),
    "vg_multi_ctx": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "vg_multi_set_reference": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vg_multi_set_msa_index": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vg_multi_upload_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64]),
    "vg_multi_upload_reads_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_int64]),
    "vg_multi_get_results_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vg_multi_map_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(VgMapStats)]),
    "vg_multi_map_pairs_msa": (C.c_int, [C.c_void_p, C.POINTER(VgMapStats)]),
    "vg_multi_result_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vg_multi_get_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "vg_multi_pileup": (C.c_int, [C.c_void_p, C.c_int32]),
    "vg_multi_pileup_counts_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "vg_multi_consensus": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vg_dpx_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vg_kernel_launches": (C.c_int64, [C.c_void_p]),
    "vg_ctx_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def lib():
    """Load the shared object and bind every declared symbol. Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvirgena_gpu.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class VgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, str(code)), msg))
        self.code = code


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
