"""ctypes binding of libraiko_kzg.so (the symbols include/raiko_kzg.h declares)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libraiko_kzg.so")

RK_OK, RK_ERR_BAD_LENGTH, RK_ERR_NONCANONICAL_FE, RK_ERR_BAD_SETTINGS, RK_ERR_BAD_POINT, RK_ERR_CUDA, RK_ERR_ARG = range(7)


class RkKzgStats(ctypes.Structure):
    _fields_ = [("msm_ms", ctypes.c_double), ("fr_ms", ctypes.c_double), ("sha_ms", ctypes.c_double),
                ("finalize_ms", ctypes.c_double), ("msm_launches", ctypes.c_uint64),
                ("total_launches", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("msm_point_adds", ctypes.c_uint64), ("msm_affine_launches", ctypes.c_uint64),
                ("msm_affine_point_adds", ctypes.c_uint64)]


_P = ctypes.c_void_p
_SZ = ctypes.c_size_t
_I = ctypes.c_int

# name -> (restype, argtypes); every entry of include/raiko_kzg.h
SIGNATURES = {
    "rk_kzg_ctx_create": (_I, [_P, _SZ, ctypes.POINTER(_I), _I, ctypes.POINTER(_P)]),
    "rk_kzg_ctx_create_ex": (_I, [_P, _SZ, ctypes.POINTER(_I), _I, _I, ctypes.POINTER(_P)]),
    "rk_kzg_ctx_destroy": (None, [_P]),
    "rk_kzg_ctx_window_bits": (_I, [_P]),
    "rk_kzg_ctx_num_devices": (_I, [_P]),
    "rk_kzg_ctx_table_bytes": (ctypes.c_uint64, [_P]),
    "rk_kzg_ctx_window_reduced": (_I, [_P]),
    "rk_kzg_ctx_export_settings": (_I, [_P, _I, _P, ctypes.POINTER(_SZ)]),
    "rk_blob_to_kzg_commitment": (_I, [_P, _P, _SZ, _P]),
    "rk_kzg_to_versioned_hash": (_I, [_P, _P]),
    "rk_get_evaluation_point": (_I, [_P, _P, _SZ, _P, _P]),
    "rk_proof_of_equivalence": (_I, [_P, _P, _SZ, _P, _P, _P]),
    "rk_compute_kzg_proof": (_I, [_P, _P, _SZ, _P, _P, _P]),
    "rk_calc_kzg_proof": (_I, [_P, _P, _SZ, _P, _P]),
    "rk_commit_batch": (_I, [_P, _P, _SZ, _P, _P, _P]),
    "rk_commit_prove_batch": (_I, [_P, _P, _SZ, _P, _P, _P, _P, _P, _P]),
    "rk_compute_kzg_proof_batch": (_I, [_P, _P, _P, _SZ, _P, _P, _P]),
    "rk_compute_blob_kzg_proof_batch": (_I, [_P, _P, _P, _SZ, _P, _P]),
    "rk_verify_kzg_proof": (_I, [_P, _P, _P, _P, _P, ctypes.POINTER(_I)]),
    "rk_verify_blob_kzg_proof_batch": (_I, [_P, _P, _P, _P, _SZ, ctypes.POINTER(_I)]),
    "rk_verify_kzg_proof_batch": (_I, [_P, _P, _P, _P, _P, _SZ, ctypes.POINTER(_I)]),
    "rk_decode_blob_data_batch": (_I, [_P, _P, _SZ, _P, _P]),
    "rk_synth_blobs": (_I, [_P, ctypes.c_uint64, ctypes.c_uint32, _SZ, _P]),
    "rk_kzg_stats_enable": (None, [_P, _I]),
    "rk_kzg_stats_reset": (None, [_P]),
    "rk_kzg_stats_get": (None, [_P, ctypes.POINTER(RkKzgStats)]),
    "rk_measure_imad_peak": (_I, [_I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "rk_last_error": (ctypes.c_char_p, []),
    "rk_version": (ctypes.c_char_p, []),
}

_lib = None


def load():
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "raiko_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C raiko_b200/csrc).  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().rk_last_error().decode("utf-8", "replace")
