"""ctypes binding of libpcaa_sm100.so (the C ABI declared in include/pcaa.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpcaa_sm100.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_ELU = 0, 1
EW_MUL, EW_ADD, EW_ELU, EW_ELU_GRAD, EW_ELU_GRAD2, EW_ADD_ROWVEC = range(6)
TC_BIAS_STATS, TC_BIAS_ELU, TC_PLAIN, TC_DGRAD_ELUBN, TC_WGRAD_ACC, TC_DGRAD_ELUOUT, TC_WGRAD_STORE = range(7)
TC_T_BIAS_STATS, TC_T_AFFINE_ELU, TC_T_DGRAD_ELUBN, TC_T_AFFINE_ELU_POOL = 7, 8, 9, 10
OP_K, OP_MN, OP_T256_K, OP_T256_MN = range(4)          # pcaa_operand_layout

_p, _i, _l, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> argument ctypes (return type is always int except the two string getters / sm_count)
SIGNATURES = {
    "pcaa_gemm_simt": [_p, _i, _l, _l, _p, _i, _l, _l, _p, _i, _l, _l, _l, _l, _l, _p, _i, _i, _p],
    "pcaa_gemm_tc": [_p, _l, _i, _p, _l, _i, _p, _l, _i, _l, _l, _l, _i, _p, _p, _p, _l, _p, _p, _p, _p, _p],
    "pcaa_gemm_tc_tn": [_p, _l, _p, _l, _p, _l, _l, _l, _l, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "pcaa_gemm_tc_nt_wgrad": [_p, _l, _p, _l, _p, _l, _l, _l, _l, _p],
    "pcaa_pointnet_l1_fwd": [_p, _p, _p, _p, _p, _l, _l, _i, _p],
    "pcaa_pointnet_l1_wgrad": [_p, _p, _p, _l, _l, _i, _p],
    "pcaa_pointnet_l1_fwd_t": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _p],
    "pcaa_pointnet_l1_wgrad_t": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _p],
    "pcaa_pointnet_l1_fwd_bn_t": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _p],
    "pcaa_input_moments": [_p, _l, _l, _p, _p],
    "pcaa_bn_from_input_moments": [_p, _l, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p],
    "pcaa_bn_elu_apply_t": [_p, _p, _p, _p, _l, _i, _p],
    "pcaa_bn_bwd_apply_t": [_p, _p, _p, _p, _p, _p, _l, _i, _p],
    "pcaa_bn_elu_meanpool_t": [_p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _p],
    "pcaa_pool_bwd_stats": [_p, _p, _p, _p, _l, _i, _i, _p],
    "pcaa_pool_bwd_apply_t": [_p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _p],
    "pcaa_colstats": [_p, _i, _l, _i, _p, _p],
    "pcaa_bn_finalize": [_p, _l, _i, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p],
    "pcaa_bn_eval_coeffs": [_p, _p, _p, _p, _f, _p, _p, _i, _p],
    "pcaa_bn_elu_apply": [_p, _i, _p, _p, _p, _i, _l, _i, _p],
    "pcaa_bn_elu_meanpool": [_p, _i, _p, _p, _p, _l, _i, _i, _p],
    "pcaa_elu_bwd_colstats": [_p, _i, _i, _p, _i, _p, _p, _p, _p, _p, _i, _p, _l, _i, _p],
    "pcaa_bn_bwd_finalize": [_p, _l, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "pcaa_bn_bwd_apply": [_p, _i, _p, _i, _p, _p, _p, _p, _i, _l, _i, _p],
    "pcaa_elu_bwd_from_out": [_p, _p, _p, _l, _p],
    "pcaa_colsum": [_p, _l, _i, _p, _p],
    "pcaa_colsum_ld": [_p, _i, _l, _i, _l, _p, _p],
    "pcaa_convert": [_p, _i, _p, _i, _l, _p],
    "pcaa_pack_bf16": [_p, _l, _l, _l, _p, _l, _i, _p],
    "pcaa_tcn_im2col": [_p, _p, _i, _l, _i, _i, _i, _p],
    "pcaa_tcn_col2im": [_p, _p, _l, _i, _i, _i, _p],
    "pcaa_tcn_bn_elu_next": [_p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _l, _i, _i, _i, _p, _p, _p],
    "pcaa_tcn_elu_bwd_stats": [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _p],
    "pcaa_tcn_bn_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _p],
    "pcaa_mean_rows": [_p, _p, _l, _i, _i, _p],
    "pcaa_mean_rows_bwd": [_p, _p, _l, _i, _i, _p],
    "pcaa_softmax_ce": [_p, _p, _p, _p, _f, _p, _l, _i, _p],
    "pcaa_chamfer_fwd": [_p, _p, _l, _i, _i, _i, _p, _p, _p, _p],
    "pcaa_chamfer_reduce": [_p, _l, _i, _i, _p, _p],
    "pcaa_chamfer_bwd": [_p, _p, _p, _p, _p, _i, _l, _i, _i, _i, _p, _p],
    "pcaa_pairwise_dist": [_p, _p, _l, _i, _i, _i, _p, _p],
    "pcaa_ew": [_i, _p, _p, _p, _l, _i, _p],
    "pcaa_wgangp_dstep": [_p] * 11 + [_f] + [_p] * 7 + [_l, _i, _p],
    "pcaa_disc_fwd": [_p] * 10 + [_f, _p, _f, _l, _i, _p],
    "pcaa_adam_flat": [_p, _p, _p, _p, _l, _f, _f, _f, _f, _i, _f, _p, _p],
    "pcaa_sum_into": [_p, _p, _l, _l, _i, _p],
    "pcaa_sum_rows": [_p, _p, _l, _l, _i, _p],
    "pcaa_gather_rows": [_p, _p, _p, _l, _l, _l, _p],
    "pcaa_adam_advance": [_p, _p, _f, _f, _f, _p],
    "pcaa_adam_flat_dev": [_p, _p, _p, _p, _l, _f, _f, _f, _p, _f, _p, _p],
    "pcaa_openset_score": [_p, _p, _l, _i, _i, _p, _p],
    "pcaa_openset_vote": [_p, _p, _l, _i, _d, _i, _p, _p],
}

_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m opensetgaitrecognition_pcaa_b200.build` "
                "(the PCAA B200 path has no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.pcaa_version.restype = C.c_char_p
        lib.pcaa_last_error.restype = C.c_char_p
        lib.pcaa_sm_count.restype = C.c_int
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = lib
    return _lib


CALLS = 0          # number of C-ABI launches issued by this process (bench.py reports it as gpu_launches)


def call(name: str, *args) -> None:
    global CALLS
    lib = load()
    CALLS += 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {lib.pcaa_last_error().decode()}")


def version() -> str:
    return load().pcaa_version().decode()
