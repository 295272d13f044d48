"""pirip_b200 — B200-native FSK receive chain (pirip's IQ->bits hot path).

This package is a thin ctypes view of the C-ABI in include/*.h (libpirip_b200.so, built
in-tree by pirip_b200/csrc/Makefile).  The product is the shared library; Python is only
used by the tests and bench.  There is no CPU fallback: if the library is missing, or no
CUDA device is usable, calls raise.
"""
from ._lib import (  # noqa: F401
    FMT_F32,
    FMT_S16,
    FMT_U8,
    FRAME_STATS_DTYPE,
    FskCfg,
    FskPlan,
    FskStream,
    DecimPlan,
    FskLdpcModem,
    LdpcPlan,
    RxPlan,
    ldpc_code,
    ldpc_encode,
    PiripError,
    demod_batch_multi,
    kernel_launches,
    lib,
    lib_path,
    pack_bits_dev,
    version,
)
