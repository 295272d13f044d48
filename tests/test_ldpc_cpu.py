"""CPU tests of the host side of the FEC/framing stage (no GPU compute calls): code tables,
RA encoder, CRC16 / pack / unpack, framer layout — product functions against the oracle."""
import ctypes as C

import numpy as np
import pytest

import orc

CODES = ["HRA_112_112", "H_256_512_4", "H_128_256_5"]


def H_matrix(v):
    N, NPB = v["CodeLength"], v["NumberParityBits"]
    K = N - NPB
    H = np.zeros((NPB, N), np.uint8)
    for p in range(NPB):
        for i in range(v["max_row_weight"]):
            c = int(v["H_rows"][p + i * NPB])
            if c:
                H[p, c - 1] = 1
        H[p, K + p] = 1
        if p:
            H[p, K + p - 1] = 1
    return H


@pytest.mark.parametrize("name", CODES)
def test_tables_consistent_and_encoder_satisfies_all_checks(name):
    import pirip_b200 as pb
    l, v = pb.ldpc_code(name)
    N, NPB = v["CodeLength"], v["NumberParityBits"]
    K = N - NPB
    assert v["NumberRowsHcols"] == K  # repeat-accumulate layout: implicit dual-diagonal parity part
    H = H_matrix(v)
    # H_cols is the transpose view of H_rows
    for c in range(K):
        rows = sorted(int(r) - 1 for r in v["H_cols"][c::K] if r)
        assert rows == list(np.nonzero(H[:, c])[0])
    g = np.random.default_rng(3)
    for _ in range(5):
        ib = g.integers(0, 2, K, dtype=np.uint8)
        p = pb.ldpc_encode(l, ib)
        assert np.array_equal(p, orc.ldpc_encode(v, ib))
        cw = np.concatenate([ib, p])
        assert not (H.astype(np.int64) @ cw % 2).any()


def test_code_dimensions_match_the_reference_call_sites():
    import pirip_b200 as pb
    # [R README.md:200] pcc: 256 for H_256_512_4; [R script/ping:8-9] 256 data bits / 512 coded
    _, v = pb.ldpc_code("H_256_512_4")
    assert (v["CodeLength"], v["NumberParityBits"]) == (512, 256)
    _, v = pb.ldpc_code("HRA_112_112")
    assert (v["CodeLength"], v["NumberParityBits"]) == (224, 112)
    with pytest.raises(pb.PiripError):
        pb.ldpc_code("no_such_code")


def test_crc_pack_unpack_payload_bits_match_oracle():
    import pirip_b200 as pb
    L = pb._lib._ldpc_protos()
    data = np.frombuffer(b"123456789", np.uint8).copy()
    assert L.freedv_gen_crc16(data.ctypes.data, data.size) == 0x29B1
    g = np.random.default_rng(5)
    bits = g.integers(0, 2, 256, dtype=np.uint8)
    by = np.zeros(32, np.uint8)
    L.freedv_pack(by.ctypes.data, bits.ctypes.data, 256)
    ref = np.zeros(32, np.uint8)
    orc.lib().orc_pack(orc._ptr(ref), orc._ptr(bits), 256)
    assert np.array_equal(by, ref)
    back = np.zeros(256, np.uint8)
    L.freedv_unpack(back.ctypes.data, by.ctypes.data, 256)
    assert np.array_equal(back, bits)
    pl = np.zeros(256, np.uint8)
    L.ofdm_generate_payload_data_bits(pl.ctypes.data, 256)
    assert np.array_equal(pl, orc.payload_data_bits(256))
    withcrc = orc.insert_crc(pl)
    assert L.freedv_check_crc16_unpacked(withcrc.ctypes.data, 256) == 1
    withcrc[100] ^= 1
    assert L.freedv_check_crc16_unpacked(withcrc.ctypes.data, 256) == 0


@pytest.mark.parametrize("name", CODES[:2])
def test_framer_layout_uw_data_parity(name):
    """[R tx/rpitx_fsk.cpp:394-395]: frame = UW(32) | data (CRC16 in the last 16 bits) | parity"""
    import pirip_b200 as pb
    m = pb.FskLdpcModem(name)
    _, v = pb.ldpc_code(name)
    assert m.bits_per_frame == 32 + v["CodeLength"]
    payload = orc.insert_crc(orc.payload_data_bits(m.data_bits))
    fr = m.frame(payload)
    assert np.array_equal(fr, orc.fsk_ldpc_frame(v, payload))
    assert np.array_equal(fr[:32], orc.uw_bits())
    assert np.array_equal(fr[32:32 + m.data_bits], payload)
    m.close()
