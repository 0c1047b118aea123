"""The CPU oracle (oracle/liboracle.so, our C restatement) against the committed golden vectors, which were
produced by the UNMODIFIED reference (tools/gen_golden.py: the reference CLIs over pipes and its library
functions) and by the reference's own known-answer vector (src/H2064_516_sparse.h:27-33).
This is what pins the oracle on machines where /root/reference does not exist."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def g(name):
    return np.load(os.path.join(GOLD, name))


def test_ldpc_known_answer(oracle_port):
    z = g("ldpc_kat.npz")
    bits, it, pcc = oracle_port.ldpc_decode(z["llr"], max_iter=int(z["max_iter"]), pcc_init=-1)
    assert it == int(z["iters"]) == 8 and pcc == int(z["parity_ok"]) == 516
    assert np.array_equal(bits, z["detected"].astype(np.uint8))
    # encode() consistency: parity of the decoded data bits == decoded parity bits
    assert np.array_equal(oracle_port.ldpc_encode(bits[:2064]), bits[2064:])


def test_phi0(oracle_port):
    z = g("phi0.npz")
    assert np.array_equal(oracle_port.phi0(z["x"]).view(np.uint32), z["y"].view(np.uint32))
    assert oracle_port.phi0(np.float32([32768.0]))[0] == 10.0      # cvttss2si overflow quirk
    assert oracle_port.phi0(np.float32([32767.0]))[0] == 0.0


def test_sd_to_llr_and_decode(oracle_port):
    z = g("ldpc_llr.npz")
    for k in range(z["sd"].shape[0]):
        llr = oracle_port.sd_to_llr(z["sd"][k].astype(np.float64))
        assert np.array_equal(llr.view(np.uint32), z["llr"][k].view(np.uint32)), k
        for mi in (10, 100):
            bits, it, pcc = oracle_port.ldpc_decode(llr, max_iter=mi, pcc_init=-1)
            assert it == z["iters%d" % mi][k] and pcc == z["pcc%d" % mi][k], (k, mi)
            assert np.array_equal(np.packbits(bits), z["bits%d" % mi][k]), (k, mi)
    assert len(set(z["iters10"].tolist())) >= 4


def test_fsk_v1_stream(oracle_port):
    z = g("fsk_v1.npz")
    sd, log, _ = oracle_port.fsk(921416, 115177).run(z["raw"], "cu8")
    assert np.array_equal(sd.view(np.uint32), z["sd"].view(np.uint32))
    assert np.array_equal(log[:, 0].astype(np.int16), z["nin"])
    assert set(z["nin"].tolist()) == {380, 384, 388}               # the clock offset exercises nin adaptation
    res = oracle_port.deframer("v1", 10).feed(sd)
    assert res["packets"] == z["packets"].tobytes() == z["payloads"].tobytes()


def test_fsk_v2_stream(oracle_port):
    z = g("fsk_v2.npz")
    sd, _, _ = oracle_port.fsk(960000, 96000).run(z["raw"], "cs16")
    assert np.array_equal(sd.view(np.uint32), z["sd"].view(np.uint32))
    res = oracle_port.deframer("v2", 10).feed(sd)
    assert res["packets"] == z["packets"].tobytes() == z["payloads"].tobytes()


def test_fsk_4fsk_stream(oracle_port):
    z = g("fsk_4fsk.npz")
    sd, _, _ = oracle_port.fsk(921416, 115177, M=4).run(z["raw"], "cu8")
    assert np.array_equal(sd.view(np.uint32), z["sd"].view(np.uint32))
    # polarity: positive soft value = bit 1, MSB first
    n = min(len(sd) // 2, len(z["symbols"]))
    hard = (sd[0:2 * n:2] > 0).astype(int) * 2 + (sd[1:2 * n:2] > 0).astype(int)
    # the demodulator output lags the transmitted symbols by a constant few symbols: find it
    best = max(np.mean(hard[lag:n] == z["symbols"][:n - lag]) for lag in range(0, 6))
    assert best > 0.99


def test_fsk_hard_bits(oracle_port):
    """fsk_demod without -s: the arg-max tone decisions of src/fsk.c:936-959, the reference CLI's own bytes"""
    z = g("fsk_hard.npz")
    b4 = oracle_port.fsk(921416, 115177, M=4).run_bits(z["raw4"], "cu8")
    assert np.array_equal(b4, z["bits4"])
    sd4, _, _ = oracle_port.fsk(921416, 115177, M=4).run(z["raw4"], "cu8")
    assert np.mean((sd4 > 0) != z["bits4"]) > 0.01      # they are NOT the signs of the 4-FSK soft decisions
    b2 = oracle_port.fsk(921416, 115177, M=2).run_bits(z["raw2"], "cs16")
    assert np.array_equal(b2, z["bits2"])


def test_testframe_counter():
    """`fsk_demod -f` bookkeeping (src/fsk_demod.c:226-243, :304-343; host side, wenet_b200/cli/_testframes.py) on the
    reference CLI's own hard bits: the same "errs: ..." lines, whatever the block size the bits arrive in"""
    from wenet_b200.cli._testframes import TestFrames
    z = g("testframes.npz")
    want = z["stderr"].tobytes().decode()
    assert want.count("errs:") == 40
    for block in (48, 48 * 7, 48 * 64, 10 ** 6):
        tf, got = TestFrames(), []
        for k in range(0, z["bits"].size, block):
            got += [tf.line(h) for h in tf.feed(z["bits"][k:k + block])]
        assert "".join(got) == want, block
        assert (tf.frames, tf.bits) == (40, 4000)


def test_tx_side(oracle_port):
    """transmit side: the modulator restatement against the reference's fsk_mod_c output, the frame
    builder against frames the reference receiver accepted, the scramble table against tx/radio_wrappers.py's"""
    g = np.load(os.path.join(GOLD, "tx.npz"))
    assert np.array_equal(oracle_port.fsk_mod(g["bits2"], 921416, 115177, 129763, 143594, M=2).view(np.uint32),
                          g["mod2"].view(np.uint32))
    assert np.array_equal(oracle_port.fsk_mod(g["bits4"], 921416, 115177, 46071, 115177, M=4).view(np.uint32),
                          g["mod4"].view(np.uint32))
    for k, pl in enumerate(g["payloads"]):
        assert np.array_equal(oracle_port.tx_frame_bits(pl.tobytes(), "v1"), g["frames_v1"][k])
        assert np.array_equal(oracle_port.tx_frame_bits(pl.tobytes(), "v2"), g["frames_v2"][k])
    from wenet_b200 import siggen
    assert np.array_equal(siggen.scramble_bytes(), g["scramble"])
