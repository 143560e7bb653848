"""ctypes binding of libpcb_host.so: numpy-Generator-compatible sampling draws in C (csrc/pcb_host.cpp)."""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpcb_host.so")
_lib = None
_P, _i64 = C.c_void_p, C.c_int64


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m phyclone_b200.build`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.pcb_smc_draws.restype = C.c_int
        lib.pcb_smc_draws.argtypes = [_P, _i64] + [_P] * 10
        lib.pcb_draw_random.restype = C.c_double
        lib.pcb_draw_random.argtypes = [_P]
        lib.pcb_draw_integers.restype = _i64
        lib.pcb_draw_integers.argtypes = [_P, _i64]
        lib.pcb_draw_choice.restype = None
        lib.pcb_draw_choice.argtypes = [_P, _i64, _i64, _P]
        lib.pcb_draw_multinomial.restype = None
        lib.pcb_draw_multinomial.argtypes = [_P, _i64, _P, _i64, _P]
        _lib = lib
    return _lib


def bitgen_address(rng):
    """Address of numpy's bitgen_t behind a Generator, or None if `rng` is not a real Generator."""
    bg = getattr(rng, "bit_generator", None)
    if bg is None:
        return None
    return bg.ctypes.bit_generator.value
