"""tfhe_rs_string_b200 -- B200-native (sm_100a) keyswitch + programmable-bootstrap engine.

Drop-in for the shortint KS+PBS path of M-Bln/tfhe-rs-string (a tfhe-rs 0.5.0 fork) behind a C ABI
(include/b200tfhe.h, libb200tfhe.so).  This Python package is only a thin ctypes binding used by
the tests and the benchmark; the product is the shared library.  There is no CPU fallback: every
entry point raises if the CUDA library is missing or no B200 is visible.
"""
from .engine import (B200TfheError, BooleanEngine, CircuitDesc, Engine, KeyView, Params, Program, lib_path, load_library,  # noqa: F401
                     parse_server_key)

__all__ = ["B200TfheError", "Engine", "Params", "Program", "lib_path", "load_library"]
