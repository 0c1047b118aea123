#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref.

Runs only in the build container (needs /root/reference to have been compiled by `make -C oracle ref`).
The outputs are committed; tests never read /root/reference.

  fsk_v1.npz    cu8 IQ of one seeded v1 stream -> reference `fsk_demod --cu8 -s 2 921416 115177` soft decisions
                and `| drs232_ldpc` packet bytes (the real CLIs over pipes), plus the harness' nin sequence
  fsk_v2.npz    cs16 IQ of one v2 stream (Fs 960000 / Rs 96000) -> fsk_demod --cs16 -s | wenet_ldpc
  fsk_4fsk.npz  cu8 IQ of an unframed 4-FSK stream -> fsk_demod --cu8 -s 4 921416 115177 soft decisions
  ldpc_llr.npz  soft-decision blocks -> reference sd_to_llr() LLRs -> run_ldpc_decoder() bits / iterations / parity
                counts for max_iter 10 and 100
  phi0.npz      arguments around every breakpoint (+ specials) -> reference phi0()
  fsk_hard.npz  hard-bit output (fsk_demod WITHOUT -s, one byte per bit): a noisy 4-FSK cu8 stream and a noisy v1 2-FSK
                cs16 stream -> the reference CLI's bytes; `python tools/gen_golden.py hard` writes only this file
  testframes.npz  `fsk_demod -f`: the reference's known 100-bit frame sent 40 times at 7 dB -> the reference CLI's hard
                bits (stdout) and its "errs: ..." lines (stderr); `python tools/gen_golden.py hard` writes this file too
  tx.npz        transmit side: bit patterns -> the reference's fsk_mod_c samples (2-FSK at the v1
                tones, 4-FSK); three payloads and their v1 / v2 on-air frame bits, accepted by the reference receiver
                (this script checks that fsk_mod_c of those frames | fsk_demod | drs232_ldpc / wenet_ldpc returns
                the payloads before it writes them); the v2 scramble table as tx/radio_wrappers.py:386-398 lists it
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from wenet_b200 import siggen           # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def cli(args, data):
    return subprocess.run(args, input=data, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout


def gen_hard():
    fsk_demod = O.ref_cli("fsk_demod")
    raw4, _ = siggen.make_4fsk_stream(21, 2400, ebno_db=5.0, fmt="cu8")
    b4 = cli([fsk_demod, "--cu8", "4", "921416", "115177", "-", "-"], raw4.tobytes())
    s4 = cli([fsk_demod, "--cu8", "-s", "4", "921416", "115177", "-", "-"], raw4.tobytes())
    b4 = np.frombuffer(b4, np.uint8)
    # the point of the vector: arg-max decisions are not the signs of the 4-FSK soft decisions
    assert np.mean(b4 != (np.frombuffer(s4, np.float32) > 0)) > 0.01
    raw2, _ = siggen.make_stream(22, n_packets=1, ebno_db=4.0, fmt="cs16", clock_ppm=-2000.0)
    b2 = cli([fsk_demod, "--cs16", "2", "921416", "115177", "-", "-"], raw2.tobytes())
    np.savez_compressed(os.path.join(GOLD, "fsk_hard.npz"), raw4=raw4, bits4=b4, raw2=raw2,
                        bits2=np.frombuffer(b2, np.uint8))
    raw = testframe_signal()
    r = subprocess.run([fsk_demod, "--cs16", "-f", "2", "921416", "115177", "-", "-"], input=raw.tobytes(),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    assert r.stderr.count(b"errs:") >= 30
    np.savez_compressed(os.path.join(GOLD, "testframes.npz"), raw=raw, bits=np.frombuffer(r.stdout, np.uint8),
                        stderr=np.frombuffer(r.stderr, np.uint8))


def testframe_signal(n_frames=40, ebno_db=7.0, seed=5):
    """the known frame of src/fsk_demod.c:236-240 back to back through the oracle's restatement of fsk_mod_c, AWGN, cs16"""
    from wenet_b200.cli._testframes import tx_frame
    port = O.Oracle("port")
    bits = np.concatenate([np.ones(960, np.uint8), np.tile(tx_frame(), n_frames), np.ones(480, np.uint8)])
    bits = np.concatenate([bits, np.ones((-bits.size) % 48, np.uint8)])
    x = port.fsk_mod(bits, 921416, 115177, 129763, 143594, M=2).astype(np.float64) / 2.0
    x = x[0::2] + 1j * x[1::2]
    y = siggen.add_noise(x, ebno_db, 921416, 115177, np.random.default_rng(seed))
    return siggen.to_format(y / np.max(np.abs(y)), "cs16")


def main():
    if sys.argv[1:] == ["hard"]:
        gen_hard()
        return
    gen_hard()
    ref = O.Oracle("reference")
    fsk_demod, drs, wen = O.ref_cli("fsk_demod"), O.ref_cli("drs232_ldpc"), O.ref_cli("wenet_ldpc")
    os.makedirs(GOLD, exist_ok=True)

    raw, payloads = siggen.make_stream(7, n_packets=2, ebno_db=9.0, fmt="cu8", clock_ppm=2000.0)
    sd = cli([fsk_demod, "--cu8", "-s", "2", "921416", "115177", "-", "-"], raw.tobytes())
    pk = cli([drs, "-", "-"], sd)
    _, log, _ = ref.fsk(921416, 115177).run(raw, "cu8")
    np.savez_compressed(os.path.join(GOLD, "fsk_v1.npz"), raw=raw, sd=np.frombuffer(sd, np.float32),
                        packets=np.frombuffer(pk, np.uint8), nin=log[:, 0].astype(np.int16),
                        payloads=np.frombuffer(b"".join(payloads), np.uint8))

    raw, payloads = siggen.make_stream(8, n_packets=2, ebno_db=10.0, framing="v2", fmt="cs16", clock_ppm=-1500.0)
    sd = cli([fsk_demod, "--cs16", "-s", "2", "960000", "96000", "-", "-"], raw.tobytes())
    pk = cli([wen, "-", "-"], sd)
    np.savez_compressed(os.path.join(GOLD, "fsk_v2.npz"), raw=raw, sd=np.frombuffer(sd, np.float32),
                        packets=np.frombuffer(pk, np.uint8), payloads=np.frombuffer(b"".join(payloads), np.uint8))

    raw, sym = siggen.make_4fsk_stream(9, 2000, ebno_db=10.0, fmt="cu8")
    sd = cli([fsk_demod, "--cu8", "-s", "4", "921416", "115177", "-", "-"], raw.tobytes())
    np.savez_compressed(os.path.join(GOLD, "fsk_4fsk.npz"), raw=raw, sd=np.frombuffer(sd, np.float32),
                        symbols=sym.astype(np.uint8))

    rng = np.random.default_rng(2024)
    sds, llrs, res = [], [], {10: [], 100: []}
    for k, snr in enumerate([0.5, 1.5, 2.0, 2.5, 3.0, 4.0, 6.0, 9.0]):
        data = rng.integers(0, 2, 2064, dtype=np.uint8)
        if k == 5:
            data[:] = 0
        cw = np.concatenate([data, ref.ldpc_encode(data)]).astype(np.float64)
        s = ((1 - 2 * cw) * rng.uniform(0.3, 3.0) + 10 ** (-snr / 20) * rng.standard_normal(2580)).astype(np.float32)
        llr = ref.sd_to_llr(s.astype(np.float64))
        sds.append(s); llrs.append(llr)
        for mi in (10, 100):
            bits, it, pcc = ref.ldpc_decode(llr, max_iter=mi, pcc_init=-1)
            res[mi].append((np.packbits(bits), it, pcc))
    np.savez_compressed(os.path.join(GOLD, "ldpc_llr.npz"), sd=np.stack(sds), llr=np.stack(llrs),
                        bits10=np.stack([r[0] for r in res[10]]), iters10=np.array([r[1] for r in res[10]], np.int32),
                        pcc10=np.array([r[2] for r in res[10]], np.int32),
                        bits100=np.stack([r[0] for r in res[100]]), iters100=np.array([r[1] for r in res[100]], np.int32),
                        pcc100=np.array([r[2] for r in res[100]], np.int32))

    z = np.load(os.path.join(ROOT, "wenet_b200", "data", "h2064_516.npz"))
    xs = [0.0, -0.0, -1.0, 1e-6, 8.6e-5, 0.5, 1.0, 5.0, 9.999, 10.0, 16.0, 32767.9, 32768.0, 1e9, np.inf, np.nan]
    brk = z["phi0_brk"] if "phi0_brk" in z.files else None
    if brk is None:
        import re
        txt = open(os.path.join(ROOT, "wenet_b200", "csrc", "wb_tables.h")).read()
        body = txt[txt.index("wb_phi0_brk[103] = {") + 20:]
        brk = np.array([int(v) for v in re.findall(r"-?\d+", body[:body.index("}")])])
    for b in brk:
        v = np.float32(b / 65536.0)
        xs += [np.nextafter(v, np.float32(-np.inf)), v, np.nextafter(v, np.float32(np.inf))]
    x = np.array(xs, dtype=np.float32)
    np.savez_compressed(os.path.join(GOLD, "phi0.npz"), x=x, y=ref.phi0(x))

    # ---- transmit side ----
    import re
    port = O.Oracle("port")
    rng = np.random.default_rng(77)
    bits2 = rng.integers(0, 2, 48 * 6).astype(np.uint8)
    bits4 = rng.integers(0, 2, 96 * 4).astype(np.uint8)
    mod2 = ref.fsk_mod(bits2, 921416, 115177, 129763, 143594, M=2)
    mod4 = ref.fsk_mod(bits4, 921416, 115177, 46071, 115177, M=4)
    payloads = siggen.random_payloads(rng, 3)
    frames = {}
    for framing, cfg in (("v1", siggen.V1), ("v2", siggen.V2)):
        fb = [port.tx_frame_bits(p, framing) for p in payloads]
        allb = np.concatenate([np.ones(2000, np.uint8)] + fb + [np.ones(400, np.uint8)])
        allb = np.concatenate([allb, np.ones((-allb.size) % 48, np.uint8)])
        x = ref.fsk_mod(allb, cfg["Fs"], cfg["Rs"], int(cfg["f_lo"]), int(cfg["f_hi"] - cfg["f_lo"]))
        out = O.run_ref_pipe(np.round(x.astype(np.float64) * 500.0).astype(np.int16).tobytes(), "cs16", Fs=cfg["Fs"], Rs=cfg["Rs"],
                             framing=framing)
        assert out == b"".join(payloads), "the reference receiver does not accept the %s frames" % framing
        frames[framing] = np.stack(fb)
    src = open("/root/reference/tx/radio_wrappers.py").read()
    m = re.search(r"class RFM98W_I2S.*?scramble_code = \[(.*?)\]", src, re.S)
    scramble = np.array([int(t, 16) for t in re.findall(r"0x[0-9a-fA-F]+", m.group(1))], dtype=np.uint8)
    assert scramble.size == 125
    np.savez_compressed(os.path.join(GOLD, "tx.npz"), bits2=bits2, mod2=mod2, bits4=bits4, mod4=mod4,
                        payloads=np.stack([np.frombuffer(p, np.uint8) for p in payloads]), frames_v1=frames["v1"],
                        frames_v2=frames["v2"], scramble=scramble)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
