"""ctypes binding of the C ABI declared in include/pbb.h.

The CUDA library is the product: if ``libpbb.so`` is missing or a symbol is
absent this module raises -- there is no CPU fallback anywhere in the package.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBB_LIB selects an alternative build of the same ABI (kernel-variant experiments)
LIB_PATH = os.environ.get('PBB_LIB') or os.path.join(_HERE, 'libpbb.so')

PBB_C64, PBB_C128 = 0, 1
NORM_NONE, NORM_EIGENVALUE, NORM_TRACE = 0, 1, 2
WEIGHT_TIME, WEIGHT_CONST, WEIGHT_TIED_TIME, WEIGHT_TIED = 0, 1, 2, 3


class CacgmmOptions(ctypes.Structure):
    """struct pbb_cacgmm_options (include/pbb.h)."""
    _fields_ = [
        ('iterations', ctypes.c_int),
        ('covariance_norm', ctypes.c_int),
        ('weight_mode', ctypes.c_int),
        ('hermitize', ctypes.c_int),
        ('affiliation_eps', ctypes.c_double),
        ('eigenvalue_floor', ctypes.c_double),
        ('frames_per_block', ctypes.c_int),
        ('reserved', ctypes.c_int),
    ]


_vp, _i, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pbb.h one to one
SIGNATURES = {
    'pbb_last_error': (ctypes.c_char_p, []),
    'pbb_version': (_i, []),
    'pbb_launch_count': (ctypes.c_longlong, []),
    'pbb_profile_enable': (None, [_i]),
    'pbb_profile_reset': (None, []),
    'pbb_profile_dump': (None, []),
    'pbb_profile_dominant': (_i, [ctypes.c_char_p, _i, ctypes.POINTER(_d), ctypes.POINTER(_i)]),
    'pbb_normalize_observation': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'pbb_streamed_task_order': (_i, [_i, _i, _i, _i, ctypes.POINTER(_i)]),
    'pbb_em_dispatch': (_i, [_i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    'pbb_cacgmm_workspace_bytes': (_sz, [_i, _i, _i, _i]),
    'pbb_cacgmm_fit': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                            ctypes.POINTER(CacgmmOptions), _vp, _vp, _vp,
                            _vp, _sz, _vp, _vp]),
    'pbb_cacgmm_predict': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i,
                                _vp, _d, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    'pbb_cacgmm_mstep': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                              ctypes.POINTER(CacgmmOptions), _vp, _vp, _vp,
                              _vp, _sz, _vp, _vp]),
    'pbb_mixture_weight_over_bins': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    'pbb_cwmm_workspace_bytes': (_sz, [_i, _i, _i, _i]),
    'pbb_cwmm_fit': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _d,
                          _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    'pbb_cwmm_predict': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp, _vp]),
    'pbb_heig_batched': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    'pbb_psd_workspace_bytes': (_sz, [_i, _i, _i, _i]),
    'pbb_power_spectral_density': (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    'pbb_gev_batched': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    'pbb_solve_batched': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    'pbb_mvdr': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    'pbb_souden': (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    'pbb_blind_analytic_normalization': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'pbb_dhtv_scratch_doubles': (_sz, [_i, _i, ctypes.POINTER(_i), _i]),
    'pbb_cacg_log_pdf': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    'pbb_gaussian_log_pdf': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'pbb_gaussian_fit_scratch_doubles': (ctypes.c_size_t, [_i, _i, _i]),
    'pbb_gaussian_fit': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'pbb_log_pdf_to_affiliation': (_i, [_vp, _vp, _d, _d, _vp, _i, _vp, _d, _i, _i, _i, _i, _vp, _vp, _vp]),
    'pbb_class_weight': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'pbb_dhtv_mapping': (_i, [_vp, _i, _i, _i, ctypes.POINTER(_i), _i, _vp, _vp, _vp, _vp]),
    'pbb_dhtv_mapping_ex': (_i, [_vp, _i, _i, _i, ctypes.POINTER(_i), _i, _vp, _vp, _vp, _i, _i, _vp]),
    'pbb_apply_mapping': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    'pbb_score_matrix': (_i, [_vp, _vp, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i, _i, _vp, _vp]),
    'pbb_mapping_from_score_matrix': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    'pbb_chain_mapping': (_i, [_vp, _i, _i, _vp, _vp]),
    'pbb_rank_one_estimate': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'pbb_matvec_batched': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'pbb_apply_beamforming_vector': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    'pbb_apply_beamforming_vector_shared': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
}

_lib = None


def load():
    """Loads libpbb.so (once) and attaches the signatures.  Raises ImportError
    with build instructions if the library or any declared symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found: the CUDA library is not built. Run '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`pb_bss_b200/csrc/build.sh`. There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f'{LIB_PATH} does not export {name}; rebuild it') from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PbbError(RuntimeError):
    pass


def check(rc, what):
    """0 -> ok; < 0 -> ValueError (bad argument, LAPACK INFO<0 convention,
    cf. get_gev_vector.pyx:130-147); > 0 -> CUDA runtime failure."""
    if rc == 0:
        return
    msg = load().pbb_last_error().decode()
    if rc < 0:
        raise ValueError(f'{what}: {msg}')
    raise PbbError(f'{what}: {msg}')
