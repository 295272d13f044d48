"""GPU tests specific to the batch kernel (fsk_kernel2.cu): more captures than one launch keeps resident (capture groups
inside a CTA), streamed ingest (time-sliced H2D behind a watermark) against copy-then-run, and the no-TMA fallback.
Each is compared bit for bit with the per-capture kernel (fsk_kernel.cu, PIRIP_FSK_K2=0), which the other GPU tests pin
to the oracle, and a sample of captures directly with the oracle."""
import os

import numpy as np
import pytest

import orc
import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_mode")]

C1 = dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000))


def _batch(ncap, nbits, seed0=9000, nbase=16):
    base = [synth.capture_u8(seed0 + c, nbits, 2, 240000, 10000, 10000, 10000, ebno_db=[None, 10.0, 7.0, 4.0][c % 4])[0]
            for c in range(nbase)]
    n = min(len(x) for x in base)
    base = np.stack([x[:n] for x in base])
    return np.stack([np.roll(base[c % nbase], -2 * ((11 * c) % 499)) for c in range(ncap)])


def _run(iq, env, **kw):
    import pirip_b200 as pb
    saved = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        plan = pb.FskPlan(**C1)
        out = plan.demod_batch(iq, want_rx_filt=True, want_stats=True, **kw)
        plan.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out


def _same(a, b):
    assert np.array_equal(a["nframes"], b["nframes"])
    assert np.array_equal(a["bits"], b["bits"])
    assert a["stats"].tobytes() == b["stats"].tobytes()
    assert a["rx_filt"].tobytes() == b["rx_filt"].tobytes()


def _oracle_sample(iq, out, idx):
    cfg = orc.demod_cfg(C1["Fs"], C1["Rs"], 2, P=24, limits=C1["limits"])
    for c in idx:
        ref = orc.fsk_demod_stream(cfg, iq[c])
        nf = ref["nframes"]
        assert out["nframes"][c] == nf
        assert np.array_equal(out["bits"][c, :nf * 50], ref["bits"])
        assert np.array_equal(out["stats"][c, :nf]["norm_rx_timing"].view(np.uint32),
                              ref["stats"]["norm_rx_timing"].view(np.uint32))


def test_more_captures_than_one_launch_keeps_resident():
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ncap = sms * 9 + 61  # one capture group of 9 per CTA, then a second, ragged one
    iq = _batch(ncap, 400)
    new = _run(iq, {"PIRIP_FSK_K2": "1"})
    old = _run(iq, {"PIRIP_FSK_K2": "0"})
    _same(new, old)
    _oracle_sample(iq, new, [0, 1, sms * 9 - 1, sms * 9, ncap - 1])


def test_streamed_ingest_equals_copy_then_run():
    # long enough for sixteen time slices (>= 4 * nin_max samples per slice)
    iq = _batch(40, 3400)
    assert iq.shape[1] // 2 >= 64 * 1206
    # the exact kernel copies first by default (its device-side wait can only fail the launch); the overlap is opt-in
    a = _run(iq, {"PIRIP_FSK_K2": "1", "PIRIP_FSK_NO_STREAM_INGEST": None, "PIRIP_FSK_EXACT_STREAM_INGEST": "1"})
    b = _run(iq, {"PIRIP_FSK_K2": "1", "PIRIP_FSK_NO_STREAM_INGEST": "1", "PIRIP_FSK_EXACT_STREAM_INGEST": None})
    _same(a, b)
    _oracle_sample(iq, a, [0, 7, 39])


def test_batch_kernel_without_tma_staging():
    iq = _batch(20, 900)
    a = _run(iq, {"PIRIP_FSK_K2": "1", "PIRIP_FSK_NO_TMA": "1"})
    b = _run(iq, {"PIRIP_FSK_K2": "1", "PIRIP_FSK_NO_TMA": None})
    _same(a, b)
    _oracle_sample(iq, a, [0, 19])


def test_kernel_description_names_the_batch_kernel():
    import ctypes
    import pirip_b200 as pb
    plan = pb.FskPlan(**C1)
    L = pb.lib()
    L.pirip_fsk_plan_kernel.restype = ctypes.c_char_p
    L.pirip_fsk_plan_kernel.argtypes = [ctypes.c_void_p]
    assert L.pirip_fsk_plan_kernel(plan._h).decode().startswith("fsk2_kernel")
    plan.close()


@pytest.mark.parametrize("per_cta", [9, 8, 3])
def test_post_decision_schedules_give_identical_output(per_cta):
    """static split / task queue (FFT tasks, sums, tone picking from one counter) / queue in mixed order, and FFT tasks of
    the next frame on capture-less worker warps during the streaming phase: same floats, same bits.  9 captures per CTA
    leaves a ragged FFT round (the queue's default case), 8 fills whole rounds (static by default), 3 leaves idle workers."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ncap = sms * per_cta if per_cta < 9 else sms * 8 + 37
    iq = _batch(ncap, 300)
    ref = _run(iq, {"PIRIP_FSK_K2": "1", "PIRIP_FSK_K2_QUEUE": "0", "PIRIP_FSK_K2_FFT_EARLY": None})
    for env in ({"PIRIP_FSK_K2_QUEUE": "1"}, {"PIRIP_FSK_K2_QUEUE": "2"}, {"PIRIP_FSK_K2_QUEUE": None},
                {"PIRIP_FSK_K2_QUEUE": "0", "PIRIP_FSK_K2_FFT_EARLY": "3"}, {"PIRIP_FSK_K2_QUEUE": "2", "PIRIP_FSK_K2_FFT_EARLY": "2"}):
        e = {"PIRIP_FSK_K2": "1", "PIRIP_FSK_K2_QUEUE": None, "PIRIP_FSK_K2_FFT_EARLY": None}
        e.update(env)
        _same(_run(iq, e), ref)
    _oracle_sample(iq, ref, [0, ncap // 2, ncap - 1])
