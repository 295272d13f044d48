"""Golden regression vectors (tests/golden/*.npz, made by tools/make_golden.py from the oracle:
PARITY UNPINNED, see that script).  CPU: the oracle still reproduces them.  GPU: the CUDA path
reproduces them bit for bit, floats included."""
import glob
import os

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
CFGS = {
    "c1_2fsk_p24": dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000)),
    "c3_4fsk_soft": dict(Fs=240000, Rs=10000, M=4, P=8, limits=(500, 45000), soft=True),
    "mask_4fsk": dict(Fs=200000, Rs=10000, M=4, P=20, tone_spacing=10000, mask=True, soft=True),
}


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def test_fixtures_present():
    assert sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))) == sorted(CFGS)


@pytest.mark.parametrize("name", list(CFGS))
def test_oracle_reproduces_golden(name):
    kw, g = CFGS[name], load(name)
    cfg = orc.demod_cfg(kw["Fs"], kw["Rs"], kw["M"], P=kw["P"], limits=kw.get("limits"), mask=kw.get("mask", False),
                        tone_spacing=kw.get("tone_spacing", 400), soft=kw.get("soft", False))
    for c in range(g["iq"].shape[0]):
        r = orc.fsk_demod_stream(cfg, g["iq"][c])
        assert r["nframes"] == g["nframes"][c]
        assert np.array_equal(r["bits"], g["bits_%d" % c])
        assert np.array_equal(r["stats"].view(np.uint8).reshape(r["nframes"], -1), g["stats_%d" % c])
        assert np.array_equal(r["rx_filt"].view(np.uint32), g["rx_filt_%d" % c].view(np.uint32))
        if kw.get("soft"):
            assert np.array_equal(r["llr"].view(np.uint32), g["llr_%d" % c].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CFGS))
def test_cuda_reproduces_golden(name, exact_mode):
    import pirip_b200 as pb
    kw, g = CFGS[name], load(name)
    plan = pb.FskPlan(**kw)
    assert plan.mode == "exact"
    out = plan.demod_batch(g["iq"], want_llr=kw.get("soft", False), want_rx_filt=True, want_stats=True)
    for c in range(g["iq"].shape[0]):
        nf = int(g["nframes"][c])
        assert out["nframes"][c] == nf
        assert np.array_equal(out["bits"][c, :nf * plan.nbits], g["bits_%d" % c])
        assert np.array_equal(out["stats"][c, :nf].view(np.uint8).reshape(nf, -1), g["stats_%d" % c])
        assert np.array_equal(out["rx_filt"][c, :nf * plan.nfilt].view(np.uint32), g["rx_filt_%d" % c].view(np.uint32))
        if kw.get("soft"):
            assert np.array_equal(out["llr"][c, :nf * plan.nbits].view(np.uint32), g["llr_%d" % c].view(np.uint32))
    plan.close()
