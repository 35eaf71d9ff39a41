"""ctypes binding of ``libmtg_b200.so`` (the C ABI declared in ``include/mtg_b200.h``).

There is NO CPU fallback: if the library is missing, or no CUDA device is visible when a compute
entry point is called, the call raises.  Build the library with
``python -m mind_the_gaps_b200._build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

import numpy as np

from . import _build

MTG_MAX_TERMS = 8
MTG_MAX_PARAMS = 24

FLAG_APPLY_PRIOR = 1
FLAG_DS_MEAN = 2
FLAG_SCAN = 4
FLAG_NO_SCAN = 8

MEAN_CONSTANT, MEAN_LINEAR, MEAN_GAUSSIAN = 0, 1, 2

STATUS_TEXT = {0: "ok", 1: "covariance matrix not positive definite", 2: "non-finite result",
               3: "outside the prior bounds"}


class MtgModel(ctypes.Structure):
    """``mtg_model_t``"""
    _fields_ = [
        ("n_terms", ctypes.c_int32),
        ("kind", ctypes.c_int32 * MTG_MAX_TERMS),
        ("pidx", (ctypes.c_int32 * 4) * MTG_MAX_TERMS),
        ("extra", ctypes.c_double * MTG_MAX_TERMS),
        ("mean_kind", ctypes.c_int32),
        ("mean_pidx", ctypes.c_int32 * 4),
        ("n_full", ctypes.c_int32),
        ("full_default", ctypes.c_double * MTG_MAX_PARAMS),
        ("n_free", ctypes.c_int32),
        ("free_idx", ctypes.c_int32 * MTG_MAX_PARAMS),
        ("lo", ctypes.c_double * MTG_MAX_PARAMS),
        ("hi", ctypes.c_double * MTG_MAX_PARAMS),
    ]


class MtgLbfgsOpts(ctypes.Structure):
    """``mtg_lbfgs_opts_t``"""
    _fields_ = [("max_iter", ctypes.c_int32), ("history", ctypes.c_int32), ("ftol", ctypes.c_double),
                ("gtol", ctypes.c_double), ("max_ls", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class MtgSimConfig(ctypes.Structure):
    """``mtg_sim_config_t``"""
    _fields_ = [("n_obs", ctypes.c_int32), ("n_fft", ctypes.c_int32), ("sim_dt", ctypes.c_double),
                ("sim_duration", ctypes.c_double), ("mean", ctypes.c_double), ("pdf", ctypes.c_int32),
                ("max_iter", ctypes.c_int32), ("noise", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("sigma_noise", ctypes.c_double), ("kraft_counts", ctypes.c_double),
                ("grid", ctypes.c_void_p), ("bin_lo", ctypes.c_void_p), ("bin_hi", ctypes.c_void_p),
                ("exposures", ctypes.c_void_p), ("bkg_counts", ctypes.c_void_p), ("bkg_rate_err", ctypes.c_void_p)]


PDF_GAUSSIAN, PDF_LOGNORMAL, PDF_UNIFORM = 0, 1, 2
SIM_E13_PADDED = 1
NOISE_POISSON, NOISE_KRAFT, NOISE_GAUSSIAN = 0, 1, 2


class MtgError(RuntimeError):
    pass


_lib = None
_vp = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_mp = ctypes.POINTER(MtgModel)

# name -> (restype, argtypes); every symbol include/mtg_b200.h declares
SIGNATURES = {
    "mtg_version": (ctypes.c_int, []),
    "mtg_last_error": (ctypes.c_char_p, []),
    "mtg_device_count": (ctypes.c_int, []),
    "mtg_set_device": (ctypes.c_int, [ctypes.c_int]),
    "mtg_model_width": (ctypes.c_int, [_mp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "mtg_loglike_dev": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32, _vp]),
    "mtg_loglike_host": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32]),
    "mtg_loglike_grad_dev": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32, _vp]),
    "mtg_loglike_grad_host": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32]),
    "mtg_scan_grad_supported": (ctypes.c_int, [_mp, ctypes.c_int]),
    "mtg_lbfgs_dev": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(MtgLbfgsOpts),
                                     _vp, _vp, _vp, _vp, ctypes.c_uint32, _vp]),
    "mtg_lbfgs_host": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(MtgLbfgsOpts),
                                      _vp, _vp, _vp, _vp, ctypes.c_uint32]),
    "mtg_sim_create": (ctypes.c_int, [ctypes.POINTER(MtgSimConfig), ctypes.POINTER(ctypes.c_void_p)]),
    "mtg_sim_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "mtg_sim_trim": (ctypes.c_int, []),
    "mtg_sim_is_fused": (ctypes.c_int, [ctypes.c_void_p]),
    "mtg_power_spectrum_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "mtg_autocorr_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "mtg_lombscargle_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "mtg_sim_generate_host": (ctypes.c_int, [_vp, ctypes.c_int, _mp, _vp, _vp, ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mtg_sim_generate_dev": (ctypes.c_int, [_vp, ctypes.c_int, _mp, _vp, _vp, ctypes.c_uint64, _vp, _vp]),
    "mtg_sim_add_noise_host": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp]),
    "mtg_sim_add_noise_dev": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint64, _vp, _vp, _vp]),
    "mtg_noise_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp]),
    "mtg_sim_regular_host": (ctypes.c_int, [_vp, ctypes.c_int, _mp, _vp, _vp, ctypes.c_uint64, _vp, _vp]),
    "mtg_downsample_host": (ctypes.c_int, [ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mtg_e13_adjust_host": (ctypes.c_int, [ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp]),
    "mtg_prepare_data_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mtg_group_best_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "mtg_predict_host": (ctypes.c_int, [_mp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mtg_ensemble_dev": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                        _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32, _vp]),
    "mtg_ensemble_host": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                         _vp, _vp, _vp, ctypes.c_uint32]),
    "mtg_ensemble_injected_host": (ctypes.c_int, [_mp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                                  ctypes.c_int, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32]),
    "mtg_host_alloc": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_uint64]),
    "mtg_host_free": (ctypes.c_int, [_vp]),
    "mtg_test_math_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mtg_fp64_peak_probe": (ctypes.c_int, [_dp, _dp]),
    "mtg_l2_flush": (ctypes.c_int, [_vp, ctypes.c_uint64, _vp]),
    "mtg_launch_count": (ctypes.c_uint64, []),
}


def library_path():
    """In-tree ``csrc/libmtg_b200.so``; ``MTG_LIB`` overrides it (kernel-variant experiments)."""
    return os.environ.get("MTG_LIB") or _build.LIB


def lib():
    """Load (once) and return the shared library with argtypes set.  Raises if it is not built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise MtgError("libmtg_b200.so is not built (%s missing): run `python -m mind_the_gaps_b200._build`; "
                           "there is no CPU fallback" % path)
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError = ABI mismatch, surface it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MtgError("libmtg_b200 error %d: %s" % (rc, lib().mtg_last_error().decode("utf-8", "replace")))


def require_device():
    n = lib().mtg_device_count()
    if n < 1:
        raise MtgError("no CUDA device visible: mind_the_gaps_b200 has no CPU fallback")
    return n


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %r, got %r" % (tuple(shape), a.shape))
    return a


class _PinnedBlock(object):
    """Owns one cudaMallocHost allocation; numpy views keep it alive through ``.base``."""

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        check(lib().mtg_host_alloc(ctypes.byref(p), max(int(nbytes), 8)))
        self.address = p.value
        self.__array_interface__ = {"shape": (max(int(nbytes), 8),), "typestr": "|u1",
                                    "data": (self.address, False), "version": 3}

    def __del__(self):
        if self.address and _lib is not None:
            _lib.mtg_host_free(ctypes.c_void_p(self.address))
            self.address = None


def pinned_empty(shape, dtype=np.float64):
    """A numpy array backed by page-locked host memory (released when the last view dies)."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    count = int(np.prod(shape))
    raw = np.asarray(_PinnedBlock(count * dtype.itemsize))
    return raw[:count * dtype.itemsize].view(dtype).reshape(shape)
