"""The synthetic-signal generator restates the TRANSMIT format (reference tx/PacketTX.py:123-137, tx/ldpc_encoder.py,
tx/radio_wrappers.py:385-417); these checks tie it to constants quoted from the reference."""
import numpy as np

from wenet_b200 import siggen


def test_crc_and_frame_layout():
    assert siggen.crc16_ccitt_false(b"123456789") == 0x29B1            # CRC-16/CCITT-FALSE check value
    body = siggen.encode_body(b"\x55" * 256)
    assert len(body) == 256 + 2 + 65
    crc = siggen.crc16_ccitt_false(body[:256])
    assert body[256] == crc & 0xFF and body[257] == crc >> 8           # '<H', tx/PacketTX.py:131
    assert body[-1] & 0x0F == 0                                        # 516 parity bits, 4 pad bits
    v1 = siggen.frame_bits(b"\x00" * 256, "v1")
    assert v1.size == (16 + 4 + 323) * 10
    assert v1[:10].tolist() == [0, 1, 0, 1, 0, 1, 0, 1, 0, 1]          # 0x55 LSB first inside start 0 / stop 1
    uw = v1[160:200].tolist()                                          # AB CD EF 01, reference src/drs232_ldpc.c:77-86
    assert uw[:10] == [0, 1, 1, 0, 1, 0, 1, 0, 1, 1]
    v2 = siggen.frame_bits(b"\x00" * 256, "v2")
    assert v2.size == (16 + 4 + 323) * 8
    assert np.packbits(v2[128:160]).tolist() == [0xAB, 0xCD, 0xEF, 0x01]
    assert siggen.scramble_bytes()[0] == 0xB9                          # tx/radio_wrappers.py:385 first table byte


def test_parity_matches_oracle_encoder(oracle_port):
    rng = np.random.default_rng(0)
    for _ in range(5):
        bits = rng.integers(0, 2, 2064, dtype=np.uint8)
        assert np.array_equal(siggen.ldpc_parity_bits(bits), oracle_port.ldpc_encode(bits))


def test_streams_are_seeded_and_shaped():
    a, pa = siggen.make_stream(3, n_samples=50000, ebno_db=10.0, fmt="cf32")
    b, pb = siggen.make_stream(3, n_samples=50000, ebno_db=10.0, fmt="cf32")
    assert a.size == 100000 and np.array_equal(a, b) and pa == pb
    assert pa[0][0] == 0x55
    c, _ = siggen.make_stream(4, n_samples=50000, ebno_db=10.0, fmt="cu8")
    assert c.dtype == np.uint8 and c.size == 100000
    assert np.max(np.abs(a)) <= 1.0 + 1e-6
