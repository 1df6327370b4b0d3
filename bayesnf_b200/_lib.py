"""ctypes binding of libbnf_sm100.so (C ABI in include/bnf.h).

The library is the product: there is NO Python/CPU fallback.  If the shared
object is missing this module raises at import time; if it is present but no
sm_100 device is available every compute entry point raises ``BnfError``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbnf_sm100.so')

BNF_ABI_VERSION = 1
NORMAL, NB, ZINB = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_BF16_SIMT, PREC_BF16X3 = 0, 1, 2, 3
WS_FORWARD, WS_GRAD, WS_MAP, WS_VI = 0, 1, 2, 3
ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 1, 2, 3, 4


class BnfError(RuntimeError):
  """A CUDA / workspace / unsupported-configuration failure in libbnf_sm100."""


class Config(C.Structure):
  _fields_ = [
      ('abi_version', C.c_int32), ('input_dim', C.c_int32), ('width', C.c_int32),
      ('depth', C.c_int32), ('likelihood', C.c_int32), ('n_seasonal', C.c_int32),
      ('seasonal_freq', C.POINTER(C.c_float)), ('seasonal_harm', C.POINTER(C.c_float)),
      ('fourier_degrees', C.POINTER(C.c_int32)), ('n_interactions', C.c_int32),
      ('interactions', C.POINTER(C.c_int32)), ('input_scales', C.POINTER(C.c_double)),
  ]


class PlanInfo(C.Structure):
  _fields_ = [('num_params', C.c_int32), ('num_features', C.c_int32),
              ('padded_features', C.c_int32), ('num_leaves', C.c_int32),
              ('num_feature_groups', C.c_int32), ('sm_count', C.c_int32)]


# name -> (restype, argtypes); tests/test_host_logic.py::test_header_symbols_exported_and_bound
# checks this table against the declarations in include/bnf.h.
_P, _I32, _I64, _U64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
SIGNATURES = {
    'bnf_abi_version': (C.c_int, []),
    'bnf_last_error': (C.c_char_p, []),
    'bnf_plan_create': (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    'bnf_plan_destroy': (None, [_P]),
    'bnf_plan_info': (C.c_int, [_P, C.POINTER(PlanInfo)]),
    'bnf_plan_leaf': (C.c_int, [_P, _I32, C.c_char_p, _I32, C.POINTER(_I64),
                                C.POINTER(_I32), C.POINTER(_I32)]),
    'bnf_precision_supported': (C.c_int, [_P, _I32]),
    'bnf_workspace_bytes': (_SZ, [_P, _I32, _I32, _I32, _I32]),
    'bnf_forward': (C.c_int, [_P, _I32, _P, _I32, _P, _P, _I64, _I32, _P, _P, _SZ, _P]),
    'bnf_loglik_grad': (C.c_int, [_P, _I32, _P, _I32, _P, _P, _P, _I64, _I32, _P, _P, _P, _SZ, _P]),
    'bnf_map_steps': (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _P, _P, _P, _I64, _I32, _I32,
                                _I32, _F, _F, _P, _P, _SZ, _P]),
    'bnf_map_epochs': (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _F, _F, _U64, _I64,
                                 _P, _P, _SZ, _P]),
    'bnf_vi_steps': (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _I32, _I32, _U64, _I64, _P, _P, _I32, _I32, _I32,
                               _F, _F, _P, _P, _SZ, _P]),
    'bnf_debug_permutation': (C.c_int, [_U64, _I64, _I32, _I32, C.POINTER(_I32)]),
    'bnf_vi_step': (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _I32, _I32, _P, _U64, _P, _P, _P,
                              _I32, _I32, _F, _F, _P, _P, _SZ, _P]),
    'bnf_vi_sample': (C.c_int, [_P, _P, _P, _I32, _I32, _P, _U64, _P, _P]),
    'bnf_init_params': (C.c_int, [_P, _F, _U64, _I64, _I32, _P, _P]),
    'bnf_mixture_quantiles': (C.c_int, [_P, _P, _I32, _I32, C.POINTER(C.c_double), _I32, _I32,
                                        _P, _P, _SZ, _P]),
    'bnf_quantile_workspace_bytes': (_SZ, [_I32, _I32]),
    'bnf_nb_mixture_quantiles': (C.c_int, [_P, _P, _P, _I32, _I32, C.POINTER(C.c_double), _I32, _P, _P,
                                           _P, _SZ, _P]),
    'bnf_debug_gemm': (C.c_int, [_I32, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    'bnf_debug_philox': (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    'bnf_debug_launch_count': (C.c_uint64, []),
    'bnf_debug_profile': (C.c_int, [_I32]),
    'bnf_debug_profile_report': (C.c_int, [C.c_char_p, _I32]),
}


def _load():
  if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
        'g.build()"` (nvcc, sm_100a).  bayesnf_b200 has no CPU fallback.')
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)  # AttributeError here == ABI mismatch
    fn.restype = res
    fn.argtypes = args
  if lib.bnf_abi_version() != BNF_ABI_VERSION:
    raise ImportError('libbnf_sm100.so ABI version mismatch')
  return lib


lib = _load()


def check(rc: int):
  """Map a C status to the reference's error behaviour."""
  if rc == 0:
    return
  msg = lib.bnf_last_error().decode('utf-8', 'replace')
  if rc == ERR_INVALID:
    raise ValueError(msg)
  raise BnfError(f'libbnf_sm100 error {rc}: {msg}')
