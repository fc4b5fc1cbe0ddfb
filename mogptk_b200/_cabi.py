"""ctypes binding of libmogp_b200.so (the C ABI declared in include/mogp_b200.h).

The library is the product: if it is missing this module raises -- there is no CPU or
PyTorch fallback anywhere in the package.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmogp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mogp_b200.h")

FAMILY = {"MOSM": 0, "SM": 1, "CONV": 2, "CSM": 3, "SMLMC": 4, "UMOSM": 5, "MOHSM": 6}


def kind_code(kind):
    """`kind` argument of the C ABI: family in the low 8 bits, Rq (CSM / SM-LMC sub-components) above
    (MOGP_KIND_WITH_RQ).  Host-side kinds are strings: "MOSM", "SM", "CONV", "CSM:<Rq>", "SMLMC:<Rq>", "UMOSM", "MOHSM"."""
    fam, _, rq = str(kind).partition(":")
    return FAMILY[fam] | ((int(rq) << 8) if rq else 0)


class _Kinds(dict):
    def __missing__(self, kind):
        return kind_code(kind)


KIND = _Kinds(FAMILY)

_lib = None

c_dp = C.c_void_p      # device / host double*
c_ip = C.POINTER(C.c_int32)

class ParamEntry(C.Structure):
    """mogp_param_entry of include/mogp_b200.h."""
    _fields_ = [("raw", C.c_void_p), ("grad", C.c_void_p), ("lower", C.c_void_p), ("upper", C.c_void_p),
                ("n", C.c_int64), ("off", C.c_int64), ("type", C.c_int32), ("lower_n", C.c_int32),
                ("upper_n", C.c_int32), ("pad", C.c_int32), ("beta", C.c_double)]


_SIGNATURES = {
    "mogp_version": (C.c_int, []),
    "mogp_num_params": (C.c_int, [C.c_int] * 4),
    "mogp_create": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "mogp_destroy": (C.c_int, [C.c_void_p]),
    "mogp_last_error": (C.c_char_p, [C.c_void_p]),
    "mogp_kbuild": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_dp, c_ip,
                              c_dp, c_dp, C.c_double, c_dp, C.c_int64, C.c_void_p]),
    "mogp_kdiag": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_dp, C.c_void_p]),
    "mogp_kdiag_x": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_dp, C.c_void_p]),
    "mogp_potrf": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mogp_lml_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_dp, c_dp, c_dp,
                                C.c_double, C.c_int, c_dp, C.c_void_p]),
    "mogp_lml_grad_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_dp, c_dp, c_dp,
                                     C.c_double, C.c_int, c_dp]),
    "mogp_predict": (C.c_int, [C.c_void_p, c_dp, c_ip, C.c_int, c_dp, c_dp, C.c_void_p]),
    "mogp_alpha": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "mogp_params_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_dp, c_dp, C.c_void_p]),
    "mogp_params_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "mogp_loss_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, c_dp, c_ip, c_dp, c_dp,
                                 C.c_double, c_dp, c_dp, C.c_void_p]),
    "mogp_train_adam": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, c_dp, c_ip, c_dp, c_dp,
                                  C.c_double, c_dp, c_dp, c_dp, C.c_longlong, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_double, c_dp, C.c_void_p, C.c_void_p]),
    "mogp_dgemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_dp, C.c_int64,
                             c_dp, C.c_int64, C.c_double, c_dp, C.c_int64, C.c_void_p]),
    "mogp_trtri_kinv": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, C.c_int64, C.c_void_p, C.c_void_p]),
    "mogp_peak_fp64": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mogp_early_loss": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_double))]),
    "mogp_early_expected": (C.c_ulonglong, [C.c_void_p]),
}

# not part of the public header: tuning / host self-check hooks
_EXTRA = {
    "mogp_set_gemm_config": (None, [C.c_int]),
    "mogp_set_i8": (C.c_int, [C.c_longlong, C.c_int]),
    "mogp_set_i8_trtri_min": (C.c_int, [C.c_longlong]),
    "mogp_get_i8_min_np": (C.c_longlong, []),
    "mogp_get_i8_slices": (C.c_int, []),
    "mogp_get_i8_trtri_min": (C.c_longlong, []),
    "mogp_set_i8_potrf_min": (C.c_int, [C.c_longlong]),
    "mogp_set_i8_ts": (C.c_int, [C.c_int]),
    "mogp_get_i8_ts": (C.c_int, []),
    "mogp_set_i8_wide": (C.c_int, [C.c_int]),
    "mogp_get_i8_wide": (C.c_int, []),
    "mogp_i8_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "mogp_set_graphs": (None, [C.c_int]),
    "mogp_set_cov_minb": (C.c_int, [C.c_int]),
    "mogp_set_graph_max_np": (None, [C.c_longlong]),
    "mogp_test_fail_capture": (None, [C.c_int]),
    "mogp_set_small_tile_threshold": (None, [C.c_longlong]),
    "mogp_launch_count": (C.c_longlong, []),
    "mogp_panel_debug": (C.c_int, [C.POINTER(C.c_longlong)]),
    "mogp_probe_latency": (C.c_int, [C.POINTER(C.c_double)]),
    "mogp_probe_contention": (C.c_int, [C.POINTER(C.c_double)]),
    "mogp_probe_issue": (C.c_int, [C.POINTER(C.c_double)]),
    "mogp_set_panel_variant": (C.c_int, [C.c_int]),
    "mogp_set_trtri_pipe": (C.c_int, [C.c_int]),
    "mogp_set_rowpipe": (C.c_int, [C.c_int, C.c_longlong, C.c_int]),
    "mogp_get_rowpipe": (C.c_int, []),
    "mogp_set_launch_prio": (C.c_int, [C.c_int]),
    "mogp_set_gemm_k64": (C.c_int, [C.c_int]),
    "mogp_set_xfuse": (C.c_int, [C.c_int]),
    "mogp_set_rchol": (C.c_int, [C.c_int, C.c_longlong, C.c_longlong]),
    "mogp_get_rchol": (C.c_int, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "mogp_rchol_applies": (C.c_int, [C.c_longlong]),
    "mogp_rchol_leaf_for": (C.c_longlong, [C.c_longlong]),
    "mogp_padded_size": (C.c_longlong, [C.c_longlong]),
    "mogp_set_pad_for_rchol": (C.c_int, [C.c_int]),
    "mogp_set_rchol_overlap": (C.c_int, [C.c_int]),
    "mogp_set_rowpipe_kinv": (C.c_int, [C.c_int]),
    "mogp_set_rowpipe_wmin": (C.c_int, [C.c_int]),
    "mogp_host_kinv_chunk_start": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "mogp_set_rowpipe_super": (C.c_int, [C.c_int, C.c_int]),
    "mogp_host_rowpipe_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_ip, C.c_int]),
    "mogp_set_stamps": (C.c_int, [C.c_int]),
    "mogp_set_skip_bulk": (C.c_int, [C.c_int]),
    "mogp_set_two_level_above": (C.c_int, [C.c_longlong]),
    "mogp_set_panel_pdl": (C.c_int, [C.c_int]),
    "mogp_get_panel_pdl": (C.c_int, []),
    "mogp_panel_spans": (C.c_int, [C.c_int, C.POINTER(C.c_ulonglong), C.c_int]),
    "mogp_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "mogp_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mogp_host_inverse_plan": (C.c_int, [C.c_int, c_ip, C.c_int]),
    "mogp_host_pair_comps": (C.c_int, [C.c_int] * 4 + [c_dp, c_dp]),
    "mogp_host_chain": (C.c_int, [C.c_int] * 4 + [c_dp, c_dp, c_dp, c_dp]),
}


def header_symbols():
    """Function names declared in include/mogp_b200.h."""
    with open(HEADER_PATH) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mogp_[a-z0-9_]+)\s*\(", txt)))


def load(check_symbols=False):
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "mogptk_b200: %s is missing -- build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH, mode=os.RTLD_NOW)
        for name, (res, args) in {**_SIGNATURES, **_EXTRA}.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if check_symbols:
        missing = [s for s in header_symbols() if not hasattr(_lib, s)]
        if missing:
            raise RuntimeError("libmogp_b200.so does not export: %s" % ", ".join(missing))
        undeclared = [s for s in _SIGNATURES if s not in header_symbols()]
        if undeclared:
            raise RuntimeError("bound but not declared in the header: %s" % ", ".join(undeclared))
    return _lib
