"""Runs last: the register-streaming decimator kernel (PIRIP_DECIM_NO_TMA=1) and the TMA kernel without its per-step CTA
barrier (PIRIP_DECIM_LOCKSTEP=0) against the default TMA kernel and the oracle."""
import os

import numpy as np
import pytest

import orc
import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,n", [(45, 45 * 16 * 70 + 79 + 45 * 9), (30, 30 * 12 * 33 + 79 + 30 * 7)])
def test_decimator_kernel_variants_agree(D, n):
    import pirip_b200 as pb
    iq = synth.rng(D + 5).integers(0, 256, (4, 2 * n), dtype=np.uint8)
    p = pb.DecimPlan(D)
    f, s = p.run(iq)
    for var, val in (("PIRIP_DECIM_NO_TMA", "1"), ("PIRIP_DECIM_LOCKSTEP", "0")):
        os.environ[var] = val
        try:
            f2, s2 = p.run(iq)
        finally:
            del os.environ[var]
        assert np.array_equal(f.view(np.uint32), f2.view(np.uint32)) and np.array_equal(s, s2), var
    yf, ys = orc.csdr_decimate_stream(iq[2], D)
    assert np.array_equal(f[2].view(np.uint32), yf.view(np.uint32)) and np.array_equal(s[2], ys)
    p.close()
