"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Two interchangeable back ends with the same Python surface:

  * ``Oracle("port")``      -> oracle/liboracle.so, our own C restatement (oracle/wenet_oracle.c)
  * ``Oracle("reference")`` -> oracle/_ref/libwenet_ref.so, the unmodified reference sources
                               compiled by oracle/Makefile plus oracle/ref_harness.c
                               (FSK / LDPC / sd_to_llr / phi0 taps; no deframer, that lives in the
                               reference's main() and is reached through the _ref binaries)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package wenet_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libwenet_ref.so")

NCODE, NDATA, NPAR = 2580, 2064, 516
FMT = {"cf32": 0, "cu8": 1, "cs16": 2, "s16": 3}


def build(force=False):
    """Compile liboracle.so (always possible) and oracle/_ref (only where /root/reference exists)."""
    if force or not os.path.exists(PORT_SO) or \
            os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "wenet_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.exists("/root/reference/src/fsk.c"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_reference():
    return os.path.exists(REF_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_SO):
                build()
            self.lib = C.CDLL(PORT_SO)
            self.pre = "wo_"
        elif kind == "reference":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (build it with `make -C oracle ref` where /root/reference exists)")
            self.lib = C.CDLL(REF_SO)
            self.pre = "ref_"
        else:
            raise ValueError(kind)
        L, pre = self.lib, self.pre
        f = getattr(L, pre + "phi0_array"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_long]; f.restype = None
        f = getattr(L, pre + "sd_to_llr"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]; f.restype = None
        f = getattr(L, pre + ("ldpc_decode")); f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        f.restype = C.c_int
        f = getattr(L, pre + ("ldpc_encode" if kind == "port" else "encode")); f.argtypes = [C.c_void_p, C.c_void_p]
        f.restype = None
        f = getattr(L, pre + "fsk_create"); f.argtypes = [C.c_int] * 4; f.restype = C.c_void_p
        f = getattr(L, pre + "fsk_destroy"); f.argtypes = [C.c_void_p]; f.restype = None
        f = getattr(L, pre + "fsk_set_est_limits"); f.argtypes = [C.c_void_p, C.c_int, C.c_int]; f.restype = None
        f = getattr(L, pre + "fsk_nin"); f.argtypes = [C.c_void_p]; f.restype = C.c_int
        f = getattr(L, pre + "fsk_nbits"); f.argtypes = [C.c_void_p]; f.restype = C.c_int
        f = getattr(L, pre + "fsk_state"); f.argtypes = [C.c_void_p, C.c_void_p]; f.restype = None
        f = getattr(L, pre + "fsk_fft_est"); f.argtypes = [C.c_void_p, C.c_void_p]; f.restype = None
        f = getattr(L, pre + "fsk_eye"); f.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        f.restype = None
        f = getattr(L, pre + "fsk_run")
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_long),
                      C.c_void_p, C.c_long, C.POINTER(C.c_long)]
        f.restype = C.c_long
        f = getattr(L, pre + "fsk_run_bits")
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
        f.restype = C.c_long
        if kind == "port":
            L.wo_fsk_demod.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]; L.wo_fsk_demod.restype = None
            L.wo_crc16.argtypes = [C.c_void_p, C.c_int]; L.wo_crc16.restype = C.c_uint16
            L.wo_deframer_create.argtypes = [C.c_int, C.c_int]; L.wo_deframer_create.restype = C.c_void_p
            L.wo_deframer_destroy.argtypes = [C.c_void_p]; L.wo_deframer_destroy.restype = None
            L.wo_deframer_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long] + \
                [C.c_void_p] * 6 + [C.c_long, C.POINTER(C.c_long)]
            L.wo_deframer_feed.restype = C.c_long
            L.wo_deframer_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        else:
            L.ref_fsk_demod_sd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; L.ref_fsk_demod_sd.restype = None
            L.ref_fsk_demod_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; L.ref_fsk_demod_bits.restype = None

    # ---- scalar stages ----
    def phi0(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        getattr(self.lib, self.pre + "phi0_array")(_p(x), _p(y), x.size)
        return y

    def sd_to_llr(self, sd, n=NCODE):
        sd = np.ascontiguousarray(sd, dtype=np.float64)
        llr = np.empty(n, dtype=np.float32)
        getattr(self.lib, self.pre + "sd_to_llr")(_p(sd), _p(llr), n) if False else \
            getattr(self.lib, self.pre + "sd_to_llr")(_p(llr), _p(sd), n)
        return llr

    def ldpc_decode(self, llr, max_iter=10, pcc_init=-1):
        """-> (bits uint8[2580], iterations, parityCheckCount)."""
        llr = np.ascontiguousarray(llr, dtype=np.float32)
        assert llr.size == NCODE
        bits = np.zeros(NCODE, dtype=np.uint8)
        pcc = C.c_int(pcc_init)
        it = getattr(self.lib, self.pre + "ldpc_decode")(_p(llr), _p(bits), max_iter, C.byref(pcc))
        return bits, it, pcc.value

    def ldpc_encode(self, ibits):
        ibits = np.ascontiguousarray(ibits, dtype=np.uint8)
        assert ibits.size == NDATA
        p = np.zeros(NPAR, dtype=np.uint8)
        getattr(self.lib, self.pre + ("ldpc_encode" if self.kind == "port" else "encode"))(_p(ibits), _p(p))
        return p

    def crc16(self, data):
        assert self.kind == "port"
        d = np.frombuffer(bytes(data), dtype=np.uint8)
        return int(self.lib.wo_crc16(_p(d), d.size))

    # ---- transmit side ----
    def tx_frame_bits(self, payload, framing="v1"):
        """one on-air frame as 0/1 bytes (reference tx/PacketTX.py:123-137 + UART / scramble)"""
        assert self.kind == "port"
        bits = np.zeros(4096, dtype=np.uint8)
        pl = np.frombuffer(bytes(payload), dtype=np.uint8)
        self.lib.wo_tx_frame_bits.restype = C.c_int
        n = self.lib.wo_tx_frame_bits(_p(pl), C.c_int(pl.size), C.c_int(1 if framing == "v1" else 2), _p(bits))
        return bits[:n].copy()

    def fsk_mod(self, bits, Fs, Rs, f1_tx, fs_tx, M=2):
        """fsk_mod_c (reference src/fsk.c:1162-1204) call after call over `bits` (a multiple of Nbits) -> cf32"""
        pre, L = self.pre, self.lib
        h = C.c_void_p(getattr(L, pre + "fsk_create")(Fs, Rs, Fs // Rs, M))
        getattr(L, pre + "fsk_set_tx")(h, C.c_int(f1_tx), C.c_int(fs_tx))
        nbits = getattr(L, pre + "fsk_nbits")(h)
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        assert bits.size % nbits == 0
        ncall = bits.size // nbits
        ns = 48 * (Fs // Rs)
        out = np.zeros((ncall, 2 * ns), dtype=np.float32)
        for k in range(ncall):
            getattr(L, pre + "fsk_mod_c")(h, _p(out[k]), _p(bits[k * nbits:(k + 1) * nbits]))
        getattr(L, pre + "fsk_destroy")(h)
        return out.reshape(-1)

    # ---- FSK ----
    def fsk(self, Fs, Rs, M=2, P=None):
        return _Fsk(self, Fs, Rs, M, P if P else Fs // Rs)

    # ---- deframe + decode ----
    def deframer(self, mode, max_iter=10):
        assert self.kind == "port"
        return _Deframer(self, mode, max_iter)


class _Fsk:
    def __init__(self, o, Fs, Rs, M, P):
        self.o, self.L, self.pre = o, o.lib, o.pre
        self.h = getattr(self.L, self.pre + "fsk_create")(Fs, Rs, P, M)
        if not self.h:
            raise ValueError("invalid FSK configuration")
        self.M = M
        self.nbits = getattr(self.L, self.pre + "fsk_nbits")(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            getattr(self.L, self.pre + "fsk_destroy")(self.h)
            self.h = None

    def set_est_limits(self, lo, hi):
        getattr(self.L, self.pre + "fsk_set_est_limits")(self.h, lo, hi)

    @property
    def nin(self):
        return getattr(self.L, self.pre + "fsk_nin")(self.h)

    @staticmethod
    def _samples(raw, fmt):
        raw = np.ascontiguousarray(raw)
        if fmt == "cf32":
            raw = raw.view(np.float32) if raw.dtype == np.complex64 else raw.astype(np.float32)
            nsamp = raw.size // 2
        elif fmt == "cu8":
            assert raw.dtype == np.uint8; nsamp = raw.size // 2
        elif fmt == "cs16":
            assert raw.dtype == np.int16; nsamp = raw.size // 2
        else:
            assert raw.dtype == np.int16; nsamp = raw.size
        return raw, nsamp

    def run(self, raw, fmt="cf32"):
        """Frame loop over a whole stream -> (sd float32[], frame_log float32[frames][8], consumed)."""
        raw, nsamp = self._samples(raw, fmt)
        max_frames = nsamp // 300 + 2
        sd = np.zeros(max_frames * self.nbits, dtype=np.float32)
        log = np.zeros((max_frames, 8), dtype=np.float32)
        n_sd, cons = C.c_long(0), C.c_long(0)
        nf = getattr(self.L, self.pre + "fsk_run")(self.h, FMT[fmt], _p(raw), nsamp, _p(sd), sd.size,
                                                  C.byref(n_sd), _p(log), max_frames, C.byref(cons))
        return sd[:n_sd.value].copy(), log[:nf].copy(), cons.value

    def run_bits(self, raw, fmt="cf32"):
        """The same frame loop through fsk_demod(): the hard bits `fsk_demod` writes without -s, one byte each."""
        raw, nsamp = self._samples(raw, fmt)
        bits = np.zeros((nsamp // 300 + 2) * self.nbits, dtype=np.uint8)
        n = C.c_long(0)
        getattr(self.L, self.pre + "fsk_run_bits")(self.h, FMT[fmt], _p(raw), nsamp, _p(bits), bits.size, C.byref(n))
        return bits[:n.value].copy()

    def state(self):
        s = np.zeros(19, dtype=np.float32)
        getattr(self.L, self.pre + "fsk_state")(self.h, _p(s))
        return s

    def fft_est(self, ndft=256):
        s = np.zeros(ndft // 2, dtype=np.float32)
        getattr(self.L, self.pre + "fsk_fft_est")(self.h, _p(s))
        return s

    def eye(self):
        out = np.zeros(8 * 160, dtype=np.float32)
        a, b = C.c_int(0), C.c_int(0)
        getattr(self.L, self.pre + "fsk_eye")(self.h, C.byref(a), C.byref(b), _p(out))
        return out[:a.value * b.value].reshape(a.value, b.value).copy()


class _Deframer:
    def __init__(self, o, mode, max_iter):
        self.L = o.lib
        self.h = self.L.wo_deframer_create({"v1": 1, "v2": 2, 1: 1, 2: 2}[mode], max_iter)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.wo_deframer_destroy(self.h)
            self.h = None

    def feed(self, sd, taps=True):
        """-> dict(packets=bytes, llr, iters, pcc, crc_ok, pos, bytes258)."""
        sd = np.ascontiguousarray(sd, dtype=np.float32)
        cap = sd.size // 2584 + 2
        out = np.zeros(cap * 256, dtype=np.uint8)
        llr = np.zeros((cap, NCODE), dtype=np.float32)
        iters = np.zeros(cap, dtype=np.int32)
        pcc = np.zeros(cap, dtype=np.int32)
        ok = np.zeros(cap, dtype=np.uint8)
        pos = np.zeros(cap, dtype=np.int64)
        b258 = np.zeros((cap, 258), dtype=np.uint8)
        ncw = C.c_long(0)
        nb = self.L.wo_deframer_feed(self.h, _p(sd), sd.size, _p(out), out.size, _p(llr), _p(iters), _p(pcc),
                                     _p(ok), _p(pos), _p(b258), cap, C.byref(ncw))
        n = ncw.value
        return dict(packets=out[:nb].tobytes(), llr=llr[:n], iters=iters[:n], pcc=pcc[:n],
                    crc_ok=ok[:n], pos=pos[:n], bytes258=b258[:n])

    def counts(self):
        a, b = C.c_int(0), C.c_int(0)
        self.L.wo_deframer_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


# ---- the reference CLIs over pipes (packet-level ground truth) ----

def ref_cli(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def run_ref_pipe(raw_bytes, fmt="cs16", M=2, Fs=921416, Rs=115177, framing="v1", extra_fsk_args=()):
    """raw bytes -> fsk_demod | drs232_ldpc (or wenet_ldpc) -> packet bytes.  Needs oracle/_ref binaries."""
    fsk, dec = ref_cli("fsk_demod"), ref_cli("drs232_ldpc" if framing == "v1" else "wenet_ldpc")
    if not fsk or not dec:
        raise FileNotFoundError("oracle/_ref binaries missing")
    flag = {"cs16": ["--cs16"], "cu8": ["--cu8"], "s16": []}[fmt]
    p1 = subprocess.Popen([fsk] + flag + ["-s"] + list(extra_fsk_args) + [str(M), str(Fs), str(Rs), "-", "-"],
                          stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    p2 = subprocess.Popen([dec, "-", "-"], stdin=p1.stdout, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    p1.stdout.close()
    import threading
    t = threading.Thread(target=lambda: (p1.stdin.write(raw_bytes), p1.stdin.close()))
    t.start()
    out = p2.stdout.read()
    t.join(); p1.wait(); p2.wait()
    return out


def run_ref_deframer(sd_float32, framing="v1"):
    dec = ref_cli("drs232_ldpc" if framing == "v1" else "wenet_ldpc")
    if not dec:
        raise FileNotFoundError("oracle/_ref binaries missing")
    r = subprocess.run([dec, "-", "-"], input=np.ascontiguousarray(sd_float32, dtype=np.float32).tobytes(),
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    return r.stdout
