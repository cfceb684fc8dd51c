"""ctypes binding of libtevo.so (include/tevo.h).  There is no CPU fallback: if the CUDA library is missing
or no B200 is visible, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TEVO_LIB") or os.path.join(_HERE, "libtevo.so")

TEVO_OK, TEVO_EINVAL, TEVO_ECUDA, TEVO_ENOMEM, TEVO_ENOTCONVERGED, TEVO_EINTERNAL = range(6)


class TevoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class tevo_complex(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class tevo_trunc(C.Structure):
    _fields_ = [("has_maxdim", C.c_int), ("maxdim", C.c_int), ("mindim", C.c_int), ("has_cutoff", C.c_int),
                ("cutoff", C.c_double), ("use_absolute_cutoff", C.c_int), ("do_rel_cutoff", C.c_int)]


class tevo_spec(C.Structure):
    _fields_ = [("nkeep", C.c_int), ("nfull", C.c_int), ("truncerr", C.c_double), ("entropy", C.c_double),
                ("sum_all", C.c_double), ("sum_kept", C.c_double)]


class tevo_krylov(C.Structure):
    _fields_ = [("hermitian", C.c_int), ("tol", C.c_double), ("krylovdim", C.c_int), ("maxiter", C.c_int)]


class tevo_krylov_info(C.Structure):
    _fields_ = [("converged", C.c_int), ("numops", C.c_int), ("numiter", C.c_int)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I = C.c_int
_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)
_CP = C.c_void_p  # complex buffers are passed as raw addresses of numpy complex128 arrays

# name -> (restype, argtypes); must list every symbol include/tevo.h declares (tests/test_abi.py checks)
PROTOTYPES = {
    "tevo_version": (C.c_char_p, []),
    "tevo_ctx_create": (_I, [_I, _PP]),
    "tevo_ctx_destroy": (_I, [_P]),
    "tevo_last_error": (C.c_char_p, [_P]),
    "tevo_ctx_synchronize": (_I, [_P]),
    "tevo_ctx_stream": (_I, [_P, _PP]),
    "tevo_ctx_launch_count": (_I, [_P, C.POINTER(C.c_longlong), _I]),
    "tevo_mps_create": (_I, [_P, _I, _I, _PP]),
    "tevo_mps_free": (_I, [_P]),
    "tevo_mps_copy": (_I, [_P, _PP]),
    "tevo_mps_upload_site": (_I, [_P, _I, _I, _I, _CP]),
    "tevo_mps_download_site": (_I, [_P, _I, _CP, C.c_size_t]),
    "tevo_mps_site_dims": (_I, [_P, _I, _IP, _IP]),
    "tevo_mps_bonddims": (_I, [_P, _IP]),
    "tevo_mps_get_limits": (_I, [_P, _IP, _IP]),
    "tevo_mps_set_limits": (_I, [_P, _I, _I]),
    "tevo_mps_random": (_I, [_P, _I, C.c_ulonglong]),
    "tevo_mps_orthogonalize": (_I, [_P, _I]),
    "tevo_mps_reorthogonalize": (_I, [_P]),
    "tevo_mps_inner": (_I, [_P, _P, C.POINTER(tevo_complex)]),
    "tevo_mps_expect_site": (_I, [_P, _I, _CP, C.POINTER(tevo_complex)]),
    "tevo_apply_gate": (_I, [_P, _I, _CP, C.POINTER(tevo_trunc), _I, _I, C.POINTER(tevo_spec), _DP, _I]),
    "tevo_apply_layer": (_I, [_P, _I, _IP, _CP, _I, C.POINTER(tevo_trunc), _I, _I, C.POINTER(tevo_spec), _DP, _I,
                              _I, _CP, _CP]),
    "tevo_measure_bond": (_I, [_P, _I, _I, _CP, _CP]),
    "tevo_gate_expect": (_I, [_P, _I, _CP, C.POINTER(tevo_complex)]),
    "tevo_mpo_create": (_I, [_P, _I, _I, _PP]),
    "tevo_mpo_free": (_I, [_P]),
    "tevo_mpo_upload_site": (_I, [_P, _I, _I, _I, _CP]),
    "tevo_mps_inner_mpo": (_I, [_P, _P, _P, C.POINTER(tevo_complex)]),
    "tevo_env_create": (_I, [_P, _PP]),
    "tevo_env_free": (_I, [_P]),
    "tevo_env_position": (_I, [_P, _P, _I, _I]),
    "tevo_env_apply_host": (_I, [_P, _P, _I, _I, _CP, _CP]),
    "tevo_tdvp_twosite": (_I, [_P, _P, _I, tevo_complex, C.POINTER(tevo_krylov), C.POINTER(tevo_trunc), _I, _I,
                               C.POINTER(tevo_spec), _DP, _I, C.POINTER(tevo_krylov_info)]),
    "tevo_tdvp_onesite": (_I, [_P, _P, _I, tevo_complex, C.POINTER(tevo_krylov), C.POINTER(tevo_krylov_info)]),
    "tevo_mps_normalize_site": (_I, [_P, _I]),
    "tevo_test_zgemm": (_I, [_P, _I, _I, _I, _I, _I, tevo_complex, _CP, _I, _CP, _I, tevo_complex, _CP, _I]),
    "tevo_test_split": (_I, [_P, _I, _I, _CP, C.POINTER(tevo_trunc), _I, _I, C.POINTER(tevo_spec), _DP, _I, _CP, _CP,
                             _I]),
    "tevo_test_heig": (_I, [_P, _I, _CP, _DP, _CP]),
    "tevo_test_heig_ctas": (_I, [_P, _I, _CP, _DP, _CP, _I]),
    "tevo_mps_to_bform": (_I, [_P]),
    "tevo_apply_layer_bform": (_I, [_P, _I, _IP, _CP, C.POINTER(tevo_trunc), _I, C.POINTER(tevo_spec), _DP, _I]),
    "tevo_bform_expect_site": (_I, [_P, _I, _I, _CP, _CP]),
    "tevo_mps_get_lambda": (_I, [_P, _I, _DP, _I, _IP]),
    "tevo_mps_set_lambda": (_I, [_P, _I, _I, _DP]),
    "tevo_mps_mark_bform": (_I, [_P, _I]),
    "tevo_mps_site_devptr": (_I, [_P, _I, _PP, _IP, _IP]),
    "tevo_mps_lambda_devptr": (_I, [_P, _I, _PP, _IP]),
    "tevo_mps_set_site_from_device": (_I, [_P, _I, _I, _I, _P]),
    "tevo_mps_set_lambda_from_device": (_I, [_P, _I, _I, _P]),
    "tevo_test_zgemm_time": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _DP]),
    "tevo_test_sbr": (_I, [_P, _I, _CP, _CP, _CP, _CP, _DP, _DP, _I, _DP, _CP]),
    "tevo_test_potrf_block": (_I, [_P, _I, _I, _CP, _CP, _CP, _DP, _I, _DP, _DP]),
    "tevo_debug_set_heig_mode": (_I, [_I]),
    "tevo_debug_set_zgemm_tile": (_I, [_I]),
    "tevo_debug_set_zgemm_3m": (_I, [_I]),
    "tevo_debug_counters": (_I, [_P, _DP, _I]),
    "tevo_profile_enable": (_I, [_I]),
    "tevo_profile_read": (_I, [C.c_char_p, C.c_size_t, _I]),
}

_lib = None


def load():
    """Load libtevo.so (built by __graft_entry__.build() / make -C csrc).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TevoError(TEVO_EINTERNAL, "libtevo.so not built (%s): run `python -c 'import __graft_entry__ as g; "
                                        "g.build()'` - there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, ctx_handle=None):
    if status == TEVO_OK:
        return
    msg = load().tevo_last_error(ctx_handle)
    msg = msg.decode() if msg else "libtevo error %d" % status
    raise TevoError(status, msg)


class Context:
    """One CUDA context handle = one GPU + one stream (tevo_ctx)."""

    def __init__(self, device: int = 0):
        lib = load()
        h = C.c_void_p()
        st = lib.tevo_ctx_create(device, C.byref(h))
        if st != TEVO_OK:
            msg = lib.tevo_last_error(None)
            raise TevoError(st, msg.decode() if msg else "tevo_ctx_create failed")
        self.h = h
        self.device = device

    def synchronize(self):
        check(load().tevo_ctx_synchronize(self.h), self.h)

    def stream(self) -> int:
        s = C.c_void_p()
        check(load().tevo_ctx_stream(self.h, C.byref(s)), self.h)
        return s.value or 0

    def launch_count(self, reset=False) -> int:
        n = C.c_longlong()
        check(load().tevo_ctx_launch_count(self.h, C.byref(n), 1 if reset else 0), self.h)
        return n.value

    def close(self):
        if self.h:
            load().tevo_ctx_destroy(self.h)
            self.h = None


_default_ctx = None


def default_ctx() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def make_trunc(maxdim=None, mindim=None, cutoff=None, use_absolute_cutoff=False, do_rel_cutoff=True, **_ignored):
    t = tevo_trunc()
    t.has_maxdim = 0 if maxdim is None else 1
    t.maxdim = 0 if maxdim is None else int(maxdim)
    t.mindim = 1 if mindim is None else int(mindim)
    t.has_cutoff = 0 if cutoff is None else 1
    t.cutoff = 0.0 if cutoff is None else float(cutoff)
    t.use_absolute_cutoff = 1 if use_absolute_cutoff else 0
    t.do_rel_cutoff = 1 if do_rel_cutoff else 0
    return t


def profile_enable(on: bool = True):
    load().tevo_profile_enable(1 if on else 0)


def profile_read(reset: bool = True) -> dict:
    import json
    buf = C.create_string_buffer(1 << 16)
    st = load().tevo_profile_read(buf, len(buf), 1 if reset else 0)
    if st != TEVO_OK:
        raise TevoError(st, "tevo_profile_read failed")
    return json.loads(buf.value.decode())
