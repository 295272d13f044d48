"""Static checks of the compiled kernels (cuobjdump on the objects the Makefile leaves beside the sources; no GPU needed):
the numerics contract forbids contracted multiply-adds, and the TMA / mbarrier / packed-add
instructions the design claims must actually be in the SASS."""
import functools
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pirip_b200", "csrc")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="CUDA toolkit binaries not on PATH")


@functools.lru_cache(maxsize=None)
def _sass(obj):
    path = os.path.join(CSRC, obj)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", CSRC, "-s", obj], stdout=subprocess.DEVNULL)
    return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout


def _count(sass, pattern):
    return len(re.findall(r"\b" + pattern + r"\b", sass))


def test_no_packed_fma_anywhere():
    # ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false: packed multiplies are banned
    for obj in ("fsk_kernel2.o", "fsk_kernel.o", "csdr.o", "ldpc.o"):
        assert _count(_sass(obj), r"FFMA2") == 0, obj


def test_decimator_has_no_fma_at_all_and_uses_tma():
    s = _sass("csdr.o")
    assert _count(s, r"FFMA") == 0            # tap products and accumulations stay separate FMUL / FADD
    assert _count(s, r"UTMALDG[\w.]*") > 0    # cp.async.bulk.tensor
    assert _count(s, r"SYNCS[\w.]*") > 0      # mbarrier
    assert _count(s, r"PRMT") > 1000          # one byte-permute per table lookup


def test_fsk_kernels_fma_only_in_division_and_sqrt_sequences():
    for obj in ("fsk_kernel2.o", "fsk_kernel.o"):
        s = _sass(obj)
        ffma, mufu = _count(s, r"FFMA"), _count(s, r"MUFU\.(RCP|RSQ)")
        # IEEE division = MUFU.RCP + 5 FFMA (+ slow path), square root = MUFU.RSQ + 2 FFMA (+ slow path): a contracted
        # multiply-add of the algorithm itself (tens of thousands of FMUL / FADD pairs) would blow this ratio up
        assert mufu > 0 and ffma <= 12 * mufu, (obj, ffma, mufu)
        assert _count(s, r"UBLKCP[\w.]*") > 0  # cp.async.bulk staging of the raw samples
    assert _count(_sass("fsk_kernel2.o"), r"FADD2") > 1000  # packed integrator adds


def test_fast_kernel_has_no_packed_fma():
    # the side chain's complex multiply packs its multiplications (ocmul_pk) and keeps the additions scalar: a packed add would be
    # contracted with the multiply that feeds it
    s = _sass("fsk_fast.o")
    assert _count(s, r"FFMA2") == 0
    assert _count(s, r"FMUL2") > 0
    assert _count(s, r"FADD2") > 1000  # packed integrator adds
