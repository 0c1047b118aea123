"""Seeded synthetic Wenet downlink signals (numpy; test and benchmark input only).

Restates the transmit-side FORMAT so the receive path has something to chew on:

  payload(256) -> CRC16-CCITT-FALSE little endian        (reference tx/PacketTX.py:131)
  -> RA-LDPC parity, 516 bits -> 65 bytes zero padded      (reference tx/ldpc_encoder.py:42-52,
                                                            src/mpdecode_core.c:72-91)
  -> preamble 16 x 0x55 + unique word AB CD EF 01 + body   (reference tx/PacketTX.py:65-66,123-137)
  v1: every byte sent UART style: start 0, 8 data bits LSB first, stop 1
                                                           (reference tx/radio_wrappers.py:553-559)
  v2: body XOR-scrambled with the 125-byte table, bits MSB first, no UART framing
                                                           (reference tx/radio_wrappers.py:385-417)
  -> continuous-phase 2-FSK, bit 1 = upper tone            (reference src/fsk.c:954)
  -> AWGN at a calibrated Eb/N0, peak-normalised           (reference benchmarking/generate_lowsnr.py:70-89)

SURVEY.md section 8(d) fixes the conventions: stream s is seeded with default_rng(1000 + s), payload
byte 0 is 0x55, tones 129763 / 273357 Hz at Fs 921416 / Rs 115177.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "h2064_516.npz")
_T = None

PREAMBLE = b"\x55" * 16
UNIQUE_WORD = b"\xab\xcd\xef\x01"

V1 = dict(Fs=921416, Rs=115177, f_lo=129763.0, f_hi=273357.0)      # start_rx.sh:35-42,103-108
V2 = dict(Fs=960000, Rs=96000, f_lo=168000.0, f_hi=264000.0)       # v2 probe config


def _tables():
    global _T
    if _T is None:
        z = np.load(_DATA)
        _T = dict(hrows=z["hrows"].astype(np.int64), scramble_neg=z["scramble_neg"].astype(np.uint8))
    return _T


def crc16_ccitt_false(data: bytes) -> int:
    crc = 0xFFFF
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def ldpc_parity_bits(ibits: np.ndarray) -> np.ndarray:
    """516 accumulate-parity bits of 2064 systematic bits (repeat-accumulate code)."""
    hr = _tables()["hrows"]                      # [516][12]
    row = ibits.astype(np.int64)[hr].sum(axis=1)
    return (np.cumsum(row) & 1).astype(np.uint8)


def scramble_bytes() -> np.ndarray:
    """The 125-byte XOR table of the v2 transmitter, rebuilt from the 1000 scramble signs
    (bit k of byte n is sign index 8n + 7 - k, MSB first on air)."""
    neg = _tables()["scramble_neg"].reshape(125, 8)
    return np.packbits(neg, axis=1).reshape(125)


def encode_body(payload: bytes) -> bytes:
    """payload (<=256 B, padded with 0x55) -> 323 bytes payload+crc+parity."""
    payload = bytes(payload[:256]) + b"\x55" * (256 - len(payload[:256]))
    crc = crc16_ccitt_false(payload)
    msg = payload + bytes([crc & 0xFF, crc >> 8])
    ibits = np.unpackbits(np.frombuffer(msg, dtype=np.uint8))
    par = np.packbits(np.concatenate([ldpc_parity_bits(ibits), np.zeros(4, dtype=np.uint8)]))
    return msg + par.tobytes()


def frame_bits(payload: bytes, framing: str = "v1") -> np.ndarray:
    """One on-air frame (preamble + UW + body) as a 0/1 uint8 array in transmit order."""
    body = np.frombuffer(encode_body(payload), dtype=np.uint8)
    if framing == "v2":
        sc = scramble_bytes()
        body = body ^ sc[np.arange(body.size) % sc.size]
    raw = np.concatenate([np.frombuffer(PREAMBLE + UNIQUE_WORD, dtype=np.uint8), body])
    if framing == "v1":
        bits = np.unpackbits(raw[:, None], axis=1, bitorder="little")          # LSB first
        out = np.concatenate([np.zeros((raw.size, 1), np.uint8), bits, np.ones((raw.size, 1), np.uint8)], axis=1)
        return out.reshape(-1)
    return np.unpackbits(raw)                                                  # MSB first


def random_payloads(rng, n):
    p = rng.integers(0, 256, size=(n, 256), dtype=np.uint8)
    p[:, 0] = 0x55
    return [bytes(r) for r in p]


def stream_bits(payloads, framing="v1", lead_in=2000, gap=0, tail=400):
    """Back-to-back frames after `lead_in` idle '1' bits; `gap` idle bits between frames."""
    parts = [np.ones(lead_in, np.uint8)]
    for p in payloads:
        parts.append(frame_bits(p, framing))
        if gap:
            parts.append(np.ones(gap, np.uint8))
    parts.append(np.ones(tail, np.uint8))
    return np.concatenate(parts)


def modulate(symbols, Fs, Rs, tones, clock_ppm=0.0, phase0=0.0):
    """Continuous-phase M-FSK, unit amplitude.  `symbols` index into `tones` (Hz).
    clock_ppm stretches the transmitter's symbol clock to exercise the nin adaptation."""
    tones = np.asarray(tones, dtype=np.float64)
    ts = Fs / Rs * (1.0 + clock_ppm * 1e-6)
    n = int(np.floor(len(symbols) * ts))
    idx = np.minimum((np.arange(n) / ts).astype(np.int64), len(symbols) - 1)
    f = tones[np.asarray(symbols)[idx]]
    ph = phase0 + 2 * np.pi * np.cumsum(f) / Fs
    return np.exp(1j * ph)


def add_noise(x, ebno_db, Fs, Rs, rng, bits_per_symbol=1.0):
    """generate_lowsnr.py:70-89 with unit signal variance, then peak normalisation."""
    if ebno_db is None:
        return x / np.max(np.abs(x))
    ebno = 10.0 ** (ebno_db / 10.0)
    nvar = 1.0 * Fs / (Rs * ebno * bits_per_symbol)
    n = np.sqrt(nvar / 2.0) * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    y = x + n
    return y / np.max(np.abs(y))


def to_format(x, fmt):
    """complex -> the byte layouts fsk_demod reads (reference src/fsk_demod.c:273-296)."""
    if fmt == "cf32":
        return np.ascontiguousarray(x.astype(np.complex64)).view(np.float32)
    iq = np.empty(2 * len(x), dtype=np.float64)
    iq[0::2], iq[1::2] = x.real, x.imag
    if fmt == "cs16":
        # x1000 undoes the reference's /FDMDV_SCALE, so every format presents the same amplitude to the
        # demodulator.  That matters: the reference's LLRs scale with the raw amplitude
        # (src/mpdecode_core.c:592-594) and its phi0 saturates at 10, so the FEC only works for soft-decision
        # magnitudes of roughly 0.3..3 -- peak-normalised float/cu8 input lands there, "x30" does not.
        return np.round(iq * 1000.0).astype(np.int16)
    if fmt == "cu8":
        return np.clip(np.round(iq * 127.0 + 127.0), 0, 255).astype(np.uint8)
    if fmt == "s16":
        return np.round(x.real * 1000.0).astype(np.int16)
    raise ValueError(fmt)


def make_stream(stream_id, n_packets=None, n_samples=None, ebno_db=10.0, framing="v1", fmt="cf32",
                clock_ppm=0.0, gap=0, lead_in=2000, cfg=None):
    """One synthetic 2-FSK stream.  Give n_packets or n_samples (then enough packets are framed to
    cover n_samples and the result is cut to exactly n_samples).
    Returns (raw array in `fmt`, list of payload bytes)."""
    cfg = cfg or (V1 if framing == "v1" else V2)
    rng = np.random.default_rng(1000 + stream_id)
    ts = cfg["Fs"] // cfg["Rs"]
    bits_per_frame = (16 + 4 + 323) * (10 if framing == "v1" else 8) + gap
    if n_packets is None:
        n_packets = max(1, int(np.ceil((n_samples / ts - lead_in) / bits_per_frame)) + 1)
    payloads = random_payloads(rng, n_packets)
    bits = stream_bits(payloads, framing, lead_in=lead_in, gap=gap)
    x = modulate(bits, cfg["Fs"], cfg["Rs"], [cfg["f_lo"], cfg["f_hi"]], clock_ppm=clock_ppm)
    if n_samples is not None:
        if len(x) < n_samples:
            x = np.concatenate([x, np.exp(1j * 2 * np.pi * cfg["f_hi"] / cfg["Fs"] * np.arange(n_samples - len(x)))])
        x = x[:n_samples]
    y = add_noise(x, ebno_db, cfg["Fs"], cfg["Rs"], rng)
    return to_format(y, fmt), payloads


def make_4fsk_stream(stream_id, n_symbols, ebno_db=12.0, fmt="cf32", Fs=921416, Rs=115177):
    """Unframed random 4-FSK (the reference has no 4-FSK deframer): tones 46071 + k*115177 Hz."""
    rng = np.random.default_rng(1000 + stream_id)
    sym = rng.integers(0, 4, size=n_symbols)
    tones = [46071.0 + k * 115177.0 for k in range(4)]
    x = modulate(sym, Fs, Rs, tones)
    y = add_noise(x, ebno_db, Fs, Rs, rng, bits_per_symbol=2.0)
    return to_format(y, fmt), sym
