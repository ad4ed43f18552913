"""ctypes binding of libsaige_b200.so (the C ABI declared in include/saige_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is present, importing the
binding / creating a context raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsaige_b200.so")

NCCL_ID_BYTES = 128
ENGINE_TENSOR, ENGINE_F64, ENGINE_UMMA = 0, 1, 2

PROBE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double))
CHROM_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_crossprod_calls", "n_crossprod_columns", "n_pcg_solves",
                                         "n_pcg_iterations", "n_kernel_launches", "n_allreduce",
                                         "bytes_h2d", "bytes_d2h", "n_probe_product_reuse",
                                         "n_probe_batches_resident")]


class SaigeB200Error(RuntimeError):
    pass


_lib = None

P = C.c_void_p
I64 = C.c_int64
DP = C.c_void_p     # double* (numpy arrays passed by address)

_SIGS = {
    "sgb_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "sgb_nccl_unique_id": (C.c_int, [P]),
    "sgb_create_dist": (C.c_int, [C.c_int, C.c_int, C.c_int, P, C.POINTER(P)]),
    "sgb_destroy": (None, [P]),
    "sgb_last_error": (C.c_char_p, [P]),
    "sgb_set_engine": (C.c_int, [P, C.c_int]),
    "sgb_set_rhs_limbs": (C.c_int, [P, C.c_int]),
    "sgb_set_verbose": (C.c_int, [P, C.c_int]),
    "sgb_step2_set_batched": (C.c_int, [P, C.c_int]),
    "sgb_step2_set_chunk_bytes": (C.c_int, [P, I64]),
    "sgb_set_product_tolerance": (C.c_int, [P, C.c_double]),
    "sgb_device_sync": (C.c_int, [P]),
    "sgb_set_min_maf_for_grm": (C.c_int, [P, C.c_float]),
    "sgb_set_max_missing_rate_for_grm": (C.c_int, [P, C.c_float]),
    "sgb_set_min_mac_variance_ratio": (C.c_int, [P, C.c_float, C.c_float, C.c_int]),
    "sgb_setgeno": (C.c_int, [P, C.c_char_p, C.c_char_p, C.c_char_p, P, I64, P, I64, C.c_int, P, I64]),
    "sgb_setgeno_mem": (C.c_int, [P, P, I64, I64, P, I64, P, C.c_int, P, I64]),
    "sgb_setgeno_synth": (C.c_int, [P, I64, I64, C.c_uint64, P, P]),
    "sgb_synth_bed_rows": (C.c_int, [P, I64, I64, I64, C.c_uint64, P, P, C.c_double, P]),
    "sgb_get_total_marker": (I64, [P]),
    "sgb_get_num_qc_markers": (I64, [P]),
    "sgb_get_num_local_markers": (I64, [P]),
    "sgb_get_nnomissing": (I64, [P]),
    "sgb_get_num_vr_markers": (I64, [P]),
    "sgb_get_allele_freq_vec": (C.c_int, [P, DP]),
    "sgb_get_mac_vec": (C.c_int, [P, P]),
    "sgb_get_allele_count_vec": (C.c_int, [P, P]),
    "sgb_get_mac_vec_for_var_ratio": (C.c_int, [P, P]),
    "sgb_get_index_vec_for_var_ratio": (C.c_int, [P, P]),
    "sgb_get_is_var_ratio_geno": (C.c_int, [P]),
    "sgb_get_qcd_marker_index": (C.c_int, [P, P]),
    "sgb_get_one_snp_geno": (C.c_int, [P, I64, P]),
    "sgb_get_one_snp_geno_for_var_ratio": (C.c_int, [P, I64, P]),
    "sgb_get_one_snp_stdgeno": (C.c_int, [P, I64, DP]),
    "sgb_set_start_end_index_vec": (C.c_int, [P, P, P, C.c_int]),
    "sgb_set_start_end_index": (C.c_int, [P, C.c_int, C.c_int, C.c_int]),
    "sgb_set_diag_of_stdgeno_loco": (C.c_int, [P]),
    "sgb_get_diag_of_kin": (C.c_int, [P, DP]),
    "sgb_get_crossprod_mat_and_kin": (C.c_int, [P, DP, C.c_int, DP]),
    "sgb_get_crossprod_mat_and_kin_loco": (C.c_int, [P, DP, C.c_int, DP]),
    "sgb_get_diag_of_sigma": (C.c_int, [P, DP, DP, C.c_int, DP]),
    "sgb_get_crossprod": (C.c_int, [P, DP, C.c_int, DP, DP, C.c_int, DP]),
    "sgb_get_pcg1_of_sigma_and_vector": (C.c_int, [P, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_int, DP, P]),
    "sgb_get_coefficients": (C.c_int, [P, DP, DP, C.c_int, DP, DP, C.c_int, C.c_double, C.c_int, DP, DP, DP, DP, DP]),
    "sgb_get_ai_score": (C.c_int, [P, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double,
                                   PROBE_FN, P, DP, DP]),
    "sgb_get_ai_score_q": (C.c_int, [P, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double,
                                     PROBE_FN, P, DP, DP]),
    "sgb_fit_glmmai_rpcg": (C.c_int, [P, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_double, PROBE_FN, P]),
    "sgb_fit_glmmai_rpcg_q": (C.c_int, [P, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double,
                                        C.c_double, C.c_double, PROBE_FN, P]),
    "sgb_get_sigma_x": (C.c_int, [P, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_int, DP]),
    "sgb_get_sigma_g": (C.c_int, [P, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_int, DP]),
    "sgb_get_coef": (C.c_int, [P, C.c_int, DP, DP, C.c_int, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_int,
                               DP, DP, DP, DP, DP, DP, DP, DP, P]),
    "sgb_get_coef_loco_all": (C.c_int, [P, C.c_int, DP, DP, C.c_int, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double,
                                        DP, DP, DP, DP, DP, P, CHROM_FN, P]),
    "sgb_glmmkin_ai_pcg": (C.c_int, [P, C.c_int, DP, DP, C.c_int, DP, DP, DP, DP, C.c_int, C.c_double, C.c_int, C.c_double,
                                     C.c_int, C.c_double, C.c_int, PROBE_FN, P, DP, DP, DP, DP, DP, DP, P, P, DP, DP, DP, DP, DP, P,
                                     CHROM_FN, P]),
    "sgb_variance_ratio_markers": (C.c_int, [P, P, C.c_int, C.c_int, DP, DP, DP, C.c_int, DP, DP, DP, DP, C.c_int, C.c_double,
                                             DP, DP, DP]),
    "sgb_set_probe_stream_fixed": (C.c_int, [P, C.c_int]),
    "sgb_cal_cv": (C.c_double, [DP, C.c_int]),
    "sgb_inner_product": (C.c_double, [DP, DP, I64]),
    "sgb_step2_set_model": (C.c_int, [P, I64, C.c_int, C.c_int, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP, C.c_double, C.c_double, P]),
    "sgb_step2_set_firth": (C.c_int, [P, C.c_int, C.c_double, DP, C.c_int]),
    "sgb_step2_set_er": (C.c_int, [P, C.c_double]),
    "sgb_step2_set_condition": (C.c_int, [P, C.c_int, DP, DP, DP, DP]),
    "sgb_step2_set_variance_ratios": (C.c_int, [P, C.c_int, DP, DP, DP]),
    "sgb_step2_test_markers": (C.c_int, [P, P, I64, I64, C.c_double, C.c_double, C.c_double, C.c_int, DP]),
    "sgb_bgen_open": (C.c_int, [C.c_char_p, C.POINTER(P), C.POINTER(I64), C.POINTER(I64), C.POINTER(C.c_int)]),
    "sgb_bgen_sample_id": (C.c_int, [P, I64, C.c_char_p, C.c_int]),
    "sgb_bgen_read": (C.c_int, [P, I64, C.c_int, C.c_int, P, DP, DP, C.c_char_p, I64, C.POINTER(I64)]),
    "sgb_bgen_close": (None, [P]),
    "sgb_step2_test_dosages": (C.c_int, [P, DP, I64, I64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                         C.c_double, DP]),
    "sgb_bench_crossprod_device": (C.c_int, [P, C.c_int, C.c_int, C.c_uint64, P, P]),
    "sgb_bench_fetch_result": (C.c_int, [P, C.c_int, DP, DP]),
    "sgb_dense_grm_build": (C.c_int, [P, C.c_int]),
    "sgb_dense_grm_free": (C.c_int, [P]),
    "sgb_dense_grm_get_block": (C.c_int, [P, I64, I64, I64, I64, DP]),
    "sgb_dense_grm_info": (C.c_int, [P, DP]),
    "sgb_set_grm_mode": (C.c_int, [P, C.c_int]),
    "sgb_bench_dense_build": (C.c_int, [P, C.c_int, I64, I64]),
    "sgb_get_counters": (C.c_int, [P, C.POINTER(Counters)]),
    "sgb_reset_counters": (C.c_int, [P]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())


def lib():
    """Loads libsaige_b200.so; raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SaigeB200Error("%s not found: build it with `make -C saige_gpu_b200/csrc` "
                                 "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib
