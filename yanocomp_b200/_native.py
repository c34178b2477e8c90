'''
ctypes binding of libyanocomp_b200.so (C ABI in include/yanocomp_b200.h).

The library is built in-tree by `make -C yanocomp_b200/csrc` (or
`__graft_entry__.build()`); a missing library is a hard error -- there is no
Python / CPU fallback for the kernels.
'''
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('YC_LIB_PATH') or os.path.join(_HERE, 'csrc', 'libyanocomp_b200.so')

YC_MAX_WINDOW = 9
YC_MAX_SAMPLES = 16

ST_NOT_A_WINDOW, ST_KS_FILTERED, ST_TESTED, ST_FIT_FAILED, ST_DEGENERATE, ST_KS_ONLY, ST_GATED = range(7)

YC_EINVAL, YC_ECAPACITY, YC_ENOGPU = -1, -2, -3
YC_COV_FULL, YC_COV_DIAG = 0, 1
# fixed-width result rows of yc_gather_fit_rows (include/yanocomp_b200.h)
ROW_STATUS, ROW_KS_STAT, ROW_KS_PVAL, ROW_N_CNTRL, ROW_N_ALL, ROW_COLUMN, ROW_STATS, ROW_MU = 0, 1, 2, 3, 4, 5, 6, 14


def row_width(w):
    return 14 + 4 * w

YC_OUTLIER_MAD, YC_OUTLIER_DBSCAN = 0, 1


class YcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libyanocomp_b200 error {code}: {msg}')
        self.code = code


class Batch(C.Structure):
    _fields_ = [
        ('n_features', C.c_int32), ('n_samples', C.c_int32),
        ('n_cntrl', C.c_int32), ('max_reads', C.c_int32),
        ('max_cond_reads', C.c_int32), ('flags', C.c_int32),
        ('total_cols', C.c_int64),
        ('events', C.c_void_p), ('x_off', C.c_void_p), ('n_cols', C.c_void_p),
        ('n_reads', C.c_void_p), ('ld', C.c_void_p), ('col_off', C.c_void_p),
        ('samp_start', C.c_void_p), ('feature_key', C.c_void_p),
        ('col_pos', C.c_void_p), ('reverse', C.c_void_p), ('col_feature', C.c_void_p),
    ]


class Params(C.Structure):
    _fields_ = [
        ('window_size', C.c_int32), ('min_read_depth', C.c_int32),
        ('max_fit_depth', C.c_int32), ('gmm', C.c_int32),
        ('add_uniform', C.c_int32), ('covariance_type', C.c_int32),
        ('min_ks', C.c_double), ('fdr_threshold', C.c_double),
        ('outlier_factor', C.c_double), ('seed', C.c_uint64),
        ('dbscan_eps', C.c_double), ('outlier_method', C.c_int32), ('reserved1', C.c_int32),
    ]


class FitOut(C.Structure):
    _fields_ = [
        ('mu', C.c_void_p), ('cov', C.c_void_p), ('box', C.c_void_p),
        ('logw', C.c_void_p), ('cw', C.c_void_p), ('iters', C.c_void_p),
        ('loglik', C.c_void_p), ('preds', C.c_void_p), ('stats', C.c_void_p),
        ('post_off', C.c_void_p),
    ]


_P = C.POINTER
_PROTOTYPES = {
    'yc_version': (C.c_char_p, []),
    'yc_last_error': (C.c_char_p, []),
    'yc_kernel_launches': (C.c_uint64, []),
    'yc_smem_bytes': (C.c_int, [_P(Batch), _P(Params), _P(C.c_int64)]),
    'yc_fill_col_feature': (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'yc_assemble_batch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_void_p]),
    'yc_ks_batch': (C.c_int, [_P(Batch), _P(Params)] + [C.c_void_p] * 8),
    'yc_compact_gated': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    # (fit_list, n_fit, fit_count): fit_count = device int32 written by yc_ks_batch, or None for exactly n_fit slots
    'yc_gmm_init_batch': (C.c_int, [_P(Batch), _P(Params), C.c_void_p, C.c_int32, C.c_void_p, _P(FitOut),
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'yc_gmm_em_batch': (C.c_int, [_P(Batch), _P(Params), C.c_void_p, C.c_int32, C.c_void_p, _P(FitOut),
                                  C.c_void_p, C.c_void_p]),
    'yc_gtest_smpred_batch': (C.c_int, [_P(Batch), _P(Params), C.c_void_p, C.c_int32, C.c_void_p, _P(FitOut),
                                        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'yc_gather_fit_rows': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 5 + [_P(FitOut), C.c_void_p, C.c_void_p]),
    'yc_gmm_fit_batch': (C.c_int, [_P(Batch), _P(Params), C.c_void_p, C.c_int32, C.c_void_p, _P(FitOut), C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'yc_kolmogorov_sf_host': (C.c_double, [C.c_double]),
    'yc_chi2_sf_1dof_host': (C.c_double, [C.c_double]),
    'yc_gtest_fractions_host': (None, [_P(C.c_double), C.c_int32, C.c_int32, C.c_double, _P(C.c_double)]),
    'yc_gtest_fractions2_host': (None, [_P(C.c_double), C.c_int32, C.c_int32, C.c_double, _P(C.c_double)]),
    'yc_position_key_host': (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_int64]),
    'yc_mix64_host': (C.c_uint64, [C.c_uint64]),
    'yc_first_right_singular_vector_device': (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'yc_first_right_singular_vector_host': (C.c_int, [C.c_int32, _P(C.c_double), _P(C.c_double), _P(C.c_double)]),
    'yc_div_int_rn_check_device': (C.c_int, [C.c_int32, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]),
}

EXPORTS = tuple(_PROTOTYPES)
_lib = None


def lib():
    '''load (once) and return the ctypes handle; raises if the library is absent'''
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `make -C yanocomp_b200/csrc` '
                '(or `python -c "import __graft_entry__ as g; g.build()"`). '
                'yanocomp_b200 has no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code):
    if code != 0:
        raise YcError(code, lib().yc_last_error().decode())
