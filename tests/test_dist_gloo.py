"""world_size-2 CPU test (gloo) of the N>1 path: contiguous capture shards per rank, no
data-path collective, one all-gather of the decoded bit streams.  The per-rank
"demodulator" here is the CPU oracle — the test covers the host-side sharding and gather logic
that bench.py uses with NCCL on GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

N_CAPTURES = 5  # odd on purpose: ranks own 3 and 2 captures


def _make_inputs():
    sys.path.insert(0, HERE)
    import synth
    caps = [synth.capture_u8(900 + i, 1500, 2, 240000, 10000, 10000, 10000, ebno_db=e)[0]
            for i, e in enumerate([None, 9.0, 6.0, 12.0, 3.0])]
    n = min(len(c) for c in caps)
    return np.stack([c[:n] for c in caps])


def _demod(iq_rows):
    sys.path.insert(0, HERE)
    import orc
    cfg = orc.demod_cfg(240000, 10000, 2, P=24, limits=(500, 25000))
    outs = [orc.fsk_demod_stream(cfg, row) for row in iq_rows]
    L = max(o["bits"].size for o in outs) if outs else 0
    L = 32 * 50  # fixed row length: max_frames * Nbits
    bits = np.zeros((len(outs), L), np.uint8)
    for i, o in enumerate(outs):
        bits[i, :o["bits"].size] = o["bits"]
    return bits


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from pirip_b200.shard import gather_bits, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    iq = _make_inputs()
    a, b = shard_range(N_CAPTURES, rank, world)
    local = torch.from_numpy(_demod(iq[a:b]))
    full = gather_bits(dist, local, N_CAPTURES, world)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    sys.path.insert(0, ROOT)
    from pirip_b200.shard import shard_range
    for n in (0, 1, 5, 8, 1250, 10000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=60)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _demod(_make_inputs())
    assert np.array_equal(got, want)
