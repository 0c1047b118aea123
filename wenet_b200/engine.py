"""ctypes host for libwenet_b200.so -- the Python side of the drop-in boundary (no PyTorch).

The C ABI is declared in include/wenet_b200.h; this module binds it the way the reference binds its own
FFI (tx/ldpc_encoder.py:23-31: ctypes.CDLL + explicit argtypes) and exposes

  * ``Engine``             -- one GPU, n_streams independent IQ streams: feed / process / drain, the batched form
                              of the reference's ``fsk_demod | drs232_ldpc`` (or ``| wenet_ldpc``) pipe;
  * ``Engine.ldpc_decode_batch`` / ``Engine.sd_to_llr_batch`` -- the stage-level calls that stand in for
                              run_ldpc_decoder() / sd_to_llr() (reference src/mpdecode_core.h:35-39).

There is no CPU fallback: if the CUDA library is missing or no GPU is visible this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (WB_LIBRARY: another build of the same library, e.g. an experiment beside the shipped one for an A/B timing on one box)
LIB_PATH = os.environ.get("WB_LIBRARY") or os.path.join(_HERE, "libwenet_b200.so")

FMT = {"cf32": 0, "cu8": 1, "cs16": 2, "s16": 3}
FMT_DTYPE = {"cf32": np.float32, "cu8": np.uint8, "cs16": np.int16, "s16": np.int16}
FMT_ELEMS = {"cf32": 2, "cu8": 2, "cs16": 2, "s16": 1}          # array elements per sample
FMT_BPS = {"cf32": 8, "cu8": 2, "cs16": 4, "s16": 2}
FRAMING = {"none": 0, None: 0, "v1": 1, "v2": 2}
FLAG_KEEP_LLR = 1
FLAG_STATS = 2
FLAG_HARD_BITS = 4

WB_OK, WB_EINVAL, WB_ENOMEM, WB_ECUDA, WB_ENODEV, WB_ERANGE = 0, -1, -2, -3, -4, -5
NCODE = 2580
PACKET_BYTES = 256


class WbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libwenet_b200: %s (code %d)" % (msg, code))
        self.code = code


class WbConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("n_streams", C.c_int32),
                ("Fs", C.c_int32), ("Rs", C.c_int32), ("M", C.c_int32), ("P", C.c_int32),
                ("est_lo", C.c_int32), ("est_hi", C.c_int32), ("in_fmt", C.c_int32), ("framing", C.c_int32),
                ("ldpc_max_iter", C.c_int32), ("flags", C.c_uint32), ("chunk_samples", C.c_uint64)]


class WbStats(C.Structure):
    _fields_ = [("EbNodB", C.c_float), ("ppm", C.c_float), ("f_est", C.c_float * 4),
                ("rx_timing", C.c_float), ("foff", C.c_float), ("norm_rx_timing", C.c_float),
                ("nin", C.c_int32), ("neyetr", C.c_int32), ("neyesamp", C.c_int32),
                ("rx_eye", (C.c_float * 160) * 8), ("nfft", C.c_int32), ("samp_fft", C.c_float * 512),
                ("frames", C.c_uint64), ("packets", C.c_uint32), ("packet_errors", C.c_uint32)]


CODEWORD_DTYPE = np.dtype([("stream", "<i4"), ("seq", "<u4"), ("iters", "<i4"), ("parity_ok", "<i4"),
                           ("crc_ok", "<i4"), ("bytes", "u1", (258,)), ("pad", "u1", (2,))])
assert CODEWORD_DTYPE.itemsize == 280

# every symbol include/wenet_b200.h declares: (name, restype, argtypes)
_VP, _SZ, _U64 = C.c_void_p, C.c_size_t, C.c_uint64
class WbTxConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_packets", C.c_int32), ("lead_in_bits", C.c_int32), ("gap_bits", C.c_int32),
                ("tail_bits", C.c_int32), ("f1_tx", C.c_int32), ("fs_tx", C.c_int32), ("ebno_db", C.c_float),
                ("seed", C.c_uint64), ("ebno_db_per_stream", C.c_void_p)]


ABI = [
    ("wb_create", C.c_int, [C.POINTER(WbConfig), C.POINTER(_VP)]),
    ("wb_destroy", None, [_VP]),
    ("wb_last_error", C.c_char_p, []),
    ("wb_abi_version", C.c_int, []),
    ("wb_nin", C.c_int, [_VP, _VP]),
    ("wb_feed", C.c_int, [_VP, _VP, _VP]),
    ("wb_feed_strided", C.c_int, [_VP, _VP, _U64, _U64]),
    ("wb_process", C.c_int, [_VP]),
    ("wb_process_soft", C.c_int, [_VP, _VP, _VP]),
    ("wb_sync", C.c_int, [_VP]),
    ("wb_drain_packets", C.c_int, [_VP, C.c_int, _VP, _SZ, C.POINTER(_SZ)]),
    ("wb_drain_all_packets", C.c_int, [_VP, _VP, _SZ, C.POINTER(_SZ), C.POINTER(_U64)]),
    ("wb_drain_soft", C.c_int, [_VP, C.c_int, _VP, _SZ, C.POINTER(_SZ)]),
    ("wb_drain_hard", C.c_int, [_VP, C.c_int, _VP, _SZ, C.POINTER(_SZ)]),
    ("wb_drain_codewords", C.c_int, [_VP, _VP, _VP, _SZ, C.POINTER(_SZ)]),
    ("wb_get_stats", C.c_int, [_VP, C.c_int, C.POINTER(WbStats)]),
    ("wb_clear_estimators", C.c_int, [_VP]),
    ("wb_ldpc_decode_batch", C.c_int, [_VP, _VP, _SZ, C.c_int, _VP, _VP, _VP]),
    ("wb_sd_to_llr_batch", C.c_int, [_VP, _VP, _SZ, _VP]),
    ("wb_tx_synthesize", C.c_int, [_VP, _VP, C.POINTER(WbTxConfig), C.POINTER(_U64)]),
    ("wb_tx_read_bits", C.c_int, [_VP, C.c_int, _VP, _SZ, C.POINTER(_SZ)]),
    ("wb_dev_read_input", C.c_int, [_VP, C.c_int, _U64, _U64, _VP]),
    ("wb_dev_input", C.c_int, [_VP, C.POINTER(_VP), C.POINTER(_U64), C.POINTER(_U64)]),
    ("wb_dev_set_fill", C.c_int, [_VP, _U64]),
    ("wb_dev_replicate", C.c_int, [_VP, C.c_int, _U64, _U64]),
    ("wb_dev_ldpc_setup", C.c_int, [_VP, _VP, _SZ, _SZ]),
    ("wb_dev_ldpc_run", C.c_int, [_VP, C.c_int]),
    ("wb_dev_ldpc_result", C.c_int, [_VP, _SZ, _SZ, _VP, _VP, _VP]),
    ("wb_timer_start", C.c_int, [_VP]),
    ("wb_timer_stop", C.c_int, [_VP, C.POINTER(C.c_float)]),
    ("wb_last_kernel_ms", C.c_int, [_VP, _VP]),
    ("wb_launch_count", _U64, [_VP]),
    ("wb_last_codewords", _U64, [_VP]),
    ("wb_last_samples", _U64, [_VP]),
    ("wb_copy_probe_create", C.c_int, [C.c_int, _SZ, C.POINTER(_VP)]),
    ("wb_copy_probe_run", C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    ("wb_copy_probe_destroy", None, [_VP]),
    ("wb_host_alloc", _VP, [_SZ]),
    ("wb_host_free", None, [_VP]),
    ("wb_geometry", C.c_int, [_VP, _VP, C.c_int]),
    ("wb_enable_frame_log", C.c_int, [_VP, C.c_int]),
    ("wb_read_frame_log", C.c_int, [_VP, C.c_int, _VP, _SZ]),
]

_lib = None


def load_library(path=LIB_PATH):
    """dlopen libwenet_b200.so and bind every ABI symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C wenet_b200/csrc`); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, res, args in ABI:
        f = getattr(lib, name)          # AttributeError if the library does not export it
        f.restype = res
        f.argtypes = args
    if path == LIB_PATH:
        _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PinnedBuffer:
    """cudaHostAlloc'd staging memory exposed as a numpy array (freed with the object)."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = self._lib.wb_host_alloc(self.nbytes)
        if not self._p:
            raise WbError(WB_ENOMEM, self._lib.wb_last_error().decode())
        buf = (C.c_uint8 * self.nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        p, self._p = getattr(self, "_p", None), None
        if p:
            self.array = None
            self._lib.wb_host_free(p)


class CopyProbe:
    """Pinned host <-> device copy rate of one GPU (wb_copy_probe_*): the ceiling of the host-buffer (e2e) path.  Run on
    every GPU of a box at the same time it gives the box's concurrent-copy ceiling (bench.py, tools/micro/pcie_multi.cu)."""

    def __init__(self, device=0, nbytes=1 << 30):
        self.lib = load_library()
        self.nbytes = nbytes
        self.h = _VP()
        rc = self.lib.wb_copy_probe_create(device, nbytes, C.byref(self.h))
        if rc != 0:
            raise WbError(rc, self.lib.wb_last_error().decode())

    def run(self, iters=4, d2h=False):
        """-> GB/s of `iters` back-to-back copies of nbytes (CUDA events on the copy stream)"""
        ms = C.c_float(0)
        rc = self.lib.wb_copy_probe_run(self.h, iters, 1 if d2h else 0, C.byref(ms))
        if rc != 0:
            raise WbError(rc, self.lib.wb_last_error().decode())
        return self.nbytes * iters / (ms.value * 1e-3) / 1e9

    def close(self):
        if getattr(self, "h", None):
            self.lib.wb_copy_probe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """n_streams FSK demodulators + deframers + LDPC decoders resident on one GPU.

    Arguments follow the reference CLIs: ``fsk_demod [-p P] [--cu8|--cs16] [-b lo -u hi] M Fs Rs``
    (src/fsk_demod.c:164) and the choice of ``drs232_ldpc`` ("v1") or ``wenet_ldpc`` ("v2") behind it.
    """

    def __init__(self, n_streams, Fs=921416, Rs=115177, M=2, P=0, in_fmt="cf32", framing="v1", max_iter=0,
                 chunk_samples=1 << 20, device=0, est_limits=None, keep_llr=False, stats=False, hard_bits=False):
        self.lib = load_library()
        self.fmt = in_fmt
        cfg = WbConfig()
        cfg.struct_size = C.sizeof(WbConfig)
        cfg.device = device
        cfg.n_streams = n_streams
        cfg.Fs, cfg.Rs, cfg.M, cfg.P = Fs, Rs, M, P
        cfg.est_lo, cfg.est_hi = est_limits if est_limits else (0, 0)
        cfg.in_fmt = FMT[in_fmt]
        cfg.framing = FRAMING[framing]
        cfg.ldpc_max_iter = max_iter
        cfg.flags = ((FLAG_KEEP_LLR if keep_llr else 0) | (FLAG_STATS if stats else 0)
                     | (FLAG_HARD_BITS if hard_bits else 0))
        cfg.chunk_samples = chunk_samples
        self.h = _VP()
        self._check(self.lib.wb_create(C.byref(cfg), C.byref(self.h)))
        self.n_streams = n_streams
        self.device = device
        self.keep_llr = keep_llr
        self.chunk_samples = chunk_samples
        g = np.zeros(14, dtype=np.int32)
        self._check(self.lib.wb_geometry(self.h, _ptr(g), 14))
        (self.N, self.Nbits, self.Ts, self.P, self.Ndft, self.nmax, self.job_cap, self.sd_cap, self.Nsym, self.M,
         self.max_iter, self.packet_syms, self.streams_per_cta, self.fsk_smem_bytes) = [int(v) for v in g]
        self.Fs, self.Rs = Fs, Rs

    def _check(self, rc):
        if rc != 0:
            raise WbError(rc, self.lib.wb_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.wb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- streaming path ----
    def nin(self):
        """fsk_nin() of every stream (reference src/fsk.h:150)."""
        out = np.zeros(self.n_streams, dtype=np.uint32)
        self._check(self.lib.wb_nin(self.h, _ptr(out)))
        return out

    def _as_raw(self, a):
        a = np.ascontiguousarray(a)
        if self.fmt == "cf32" and a.dtype == np.complex64:
            a = a.view(np.float32)
        if a.dtype != FMT_DTYPE[self.fmt]:
            raise TypeError("stream data must be %s for in_fmt=%s" % (FMT_DTYPE[self.fmt], self.fmt))
        return a

    def _hold(self, arrays):
        """Host arrays handed to an asynchronous copy stay referenced until the copies have run: until sync(), or until
        the first feed after a process() -- wb_feed / wb_process_soft wait for a pending wb_process (and with it for every
        copy queued before it) before they queue anything new.  Several feeds may be queued before one process()."""
        if getattr(self, "_keep", None) is None or getattr(self, "_processed", False):
            self._keep = []
        self._processed = False
        self._keep.extend(arrays)

    def feed(self, streams):
        """Append samples: ``streams[s]`` is the raw array of stream s (None/empty to skip)."""
        assert len(streams) == self.n_streams
        keep, ptrs, ns = [], (C.c_void_p * self.n_streams)(), np.zeros(self.n_streams, dtype=np.uint64)
        for s, a in enumerate(streams):
            if a is None or len(a) == 0:
                continue
            a = self._as_raw(a)
            keep.append(a)
            ptrs[s] = a.ctypes.data
            ns[s] = a.size // FMT_ELEMS[self.fmt]
        self._check(self.lib.wb_feed(self.h, ptrs, _ptr(ns)))
        self._hold(keep)            # keep the host arrays alive until the copies have run

    def feed_strided(self, block):
        """Append the same number of samples to every stream from one [n_streams, elems] array."""
        block = self._as_raw(block) if block.flags["C_CONTIGUOUS"] else block
        assert block.shape[0] == self.n_streams and block.dtype == FMT_DTYPE[self.fmt]
        assert block.strides[-1] == block.itemsize
        nsamp = block.shape[1] // FMT_ELEMS[self.fmt]
        self._check(self.lib.wb_feed_strided(self.h, C.c_void_p(block.ctypes.data), block.strides[0], nsamp))
        self._hold([block])

    def process(self):
        self._check(self.lib.wb_process(self.h))
        self._processed = True

    def process_soft(self, streams):
        """Deframe + decode float32 soft symbols directly (what `drs232_ldpc` / `wenet_ldpc` read on stdin)."""
        assert len(streams) == self.n_streams
        keep, ptrs, ns = [], (C.c_void_p * self.n_streams)(), np.zeros(self.n_streams, dtype=np.uint64)
        for s, a in enumerate(streams):
            if a is None or len(a) == 0:
                continue
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            ptrs[s] = a.ctypes.data
            ns[s] = a.size
        self._check(self.lib.wb_process_soft(self.h, ptrs, _ptr(ns)))
        self._hold(keep)
        self._processed = True

    def sync(self):
        self._check(self.lib.wb_sync(self.h))
        self._keep = None

    def drain_packets(self, stream, max_packets=1 << 16):
        """CRC-valid 256-byte payloads of `stream`, decode order (what drs232_ldpc writes to stdout)."""
        buf = np.empty(max_packets * PACKET_BYTES, dtype=np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.wb_drain_packets(self.h, stream, _ptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].tobytes()

    def drain_all_packets(self, cap_bytes=None):
        """-> structured array of (stream, seq, payload[256]) sorted by (stream, seq)."""
        dt = np.dtype([("stream", "<i4"), ("seq", "<u4"), ("payload", "u1", (PACKET_BYTES,))])
        cap = cap_bytes or max(1 << 20, self.n_streams * self.job_cap * dt.itemsize)
        buf = np.empty(cap, dtype=np.uint8)
        n, npk = C.c_size_t(0), C.c_uint64(0)
        rc = self.lib.wb_drain_all_packets(self.h, _ptr(buf), buf.size, C.byref(n), C.byref(npk))
        if rc == WB_ERANGE:
            return self.drain_all_packets(cap_bytes=n.value)
        self._check(rc)
        return buf[:n.value].view(dt).copy()

    def drain_soft(self, stream):
        """Soft decisions `stream` produced in the last process() (what fsk_demod -s writes to stdout)."""
        buf = np.empty(self.sd_cap, dtype=np.float32)
        n = C.c_size_t(0)
        self._check(self.lib.wb_drain_soft(self.h, stream, _ptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def drain_hard(self, stream):
        """Hard bits (one byte each) `stream` produced in the last process(): what fsk_demod without -s writes,
        the rx_bits of fsk_demod() (reference src/fsk.c:936-959).  Needs hard_bits=True."""
        buf = np.empty(self.sd_cap, dtype=np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.wb_drain_hard(self.h, stream, _ptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def drain_codewords(self, with_llr=False):
        n = C.c_size_t(0)
        cap = int(self.lib.wb_last_codewords(self.h))
        cw = np.zeros(max(cap, 1), dtype=CODEWORD_DTYPE)
        llr = np.zeros((max(cap, 1), NCODE), dtype=np.float32) if with_llr else None
        self._check(self.lib.wb_drain_codewords(self.h, _ptr(cw), _ptr(llr) if with_llr else None, cap, C.byref(n)))
        return (cw[:n.value], llr[:n.value]) if with_llr else cw[:n.value]

    def stats(self, stream):
        st = WbStats()
        self._check(self.lib.wb_get_stats(self.h, stream, C.byref(st)))
        return st

    def clear_estimators(self):
        self._check(self.lib.wb_clear_estimators(self.h))

    def enable_frame_log(self, frames_per_stream):
        self._check(self.lib.wb_enable_frame_log(self.h, frames_per_stream))
        self._log_cap = frames_per_stream

    def read_frame_log(self, stream, frames):
        buf = np.zeros((min(frames, self._log_cap), 8), dtype=np.float32)
        self._check(self.lib.wb_read_frame_log(self.h, stream, _ptr(buf), buf.shape[0]))
        return buf

    # ---- stage-level ----
    def ldpc_decode_batch(self, llr, max_iter=0):
        """run_ldpc_decoder() over [n, 2580] LLRs -> (bits uint8 [n, 2580], iters [n], parityCheckCount [n])."""
        llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, NCODE)
        n = llr.shape[0]
        packed = np.zeros((n, 323), dtype=np.uint8)
        iters = np.zeros(n, dtype=np.int32)
        pcc = np.zeros(n, dtype=np.int32)
        self._check(self.lib.wb_ldpc_decode_batch(self.h, _ptr(llr), n, max_iter, _ptr(packed), _ptr(iters), _ptr(pcc)))
        bits = np.unpackbits(packed, axis=1)[:, :NCODE]
        return bits, iters, pcc

    def sd_to_llr_batch(self, sd):
        sd = np.ascontiguousarray(sd, dtype=np.float32).reshape(-1, NCODE)
        llr = np.zeros_like(sd)
        self._check(self.lib.wb_sd_to_llr_batch(self.h, _ptr(sd), sd.shape[0], _ptr(llr)))
        return llr

    # ---- HBM-resident benchmarking ----
    def dev_input(self):
        p, stride, cap = _VP(), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.wb_dev_input(self.h, C.byref(p), C.byref(stride), C.byref(cap)))
        return p.value, stride.value, cap.value

    # ---- transmit side on the device ----
    def tx_synthesize(self, payloads, f1_tx, fs_tx, ebno_db=None, lead_in=2000, gap=0, tail=400, seed=0):
        """Build every stream's input on the device, as a wb_feed would have delivered it: payloads[n_streams][n_packets]
        [256] -> framed (the engine's framing), modulated by the reference's fsk_mod_c arithmetic, optionally AWGN +
        peak normalisation (benchmarking/generate_lowsnr.py).  Returns the samples appended to every stream."""
        pl = np.ascontiguousarray(payloads, dtype=np.uint8)
        if pl.ndim != 3 or pl.shape[0] != self.n_streams or pl.shape[2] != 256:
            raise ValueError("payloads must be [n_streams][n_packets][256] bytes")
        per = None
        if ebno_db is not None and np.ndim(ebno_db) == 1:       # one Eb/N0 per stream
            per = np.ascontiguousarray(ebno_db, dtype=np.float32)
            if per.size != self.n_streams:
                raise ValueError("ebno_db: one value or one per stream")
        cfg = WbTxConfig(C.sizeof(WbTxConfig), pl.shape[1], lead_in, gap, tail, int(f1_tx), int(fs_tx),
                         float("nan") if (ebno_db is None or per is not None) else float(ebno_db), int(seed),
                         per.ctypes.data if per is not None else None)
        ns = _U64(0)
        self._check(self.lib.wb_tx_synthesize(self.h, _ptr(pl), C.byref(cfg), C.byref(ns)))
        return int(ns.value)

    def tx_read_bits(self, stream, cap=1 << 24):
        bits = np.zeros(cap, dtype=np.uint8)
        n = _SZ(0)
        self._check(self.lib.wb_tx_read_bits(self.h, stream, _ptr(bits), cap, C.byref(n)))
        return bits[:n.value].copy()

    def dev_read_input(self, stream, nsamp, first=0):
        """samples [first, first + nsamp) of a stream's resident input row, in the engine's input format (test tap)"""
        out = np.zeros(nsamp * FMT_ELEMS[self.fmt], dtype=FMT_DTYPE[self.fmt])
        self._check(self.lib.wb_dev_read_input(self.h, stream, first, nsamp, _ptr(out)))
        return out

    def dev_set_fill(self, nsamp):
        self._check(self.lib.wb_dev_set_fill(self.h, nsamp))

    def dev_replicate(self, n_src, nsamp, rot):
        self._check(self.lib.wb_dev_replicate(self.h, n_src, nsamp, rot))

    def dev_ldpc_setup(self, llr, n):
        llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, NCODE)
        self._check(self.lib.wb_dev_ldpc_setup(self.h, _ptr(llr), llr.shape[0], n))

    def dev_ldpc_run(self, max_iter=0):
        self._check(self.lib.wb_dev_ldpc_run(self.h, max_iter))

    def dev_ldpc_result(self, first, n):
        packed = np.zeros((n, 323), dtype=np.uint8)
        iters = np.zeros(n, dtype=np.int32)
        pcc = np.zeros(n, dtype=np.int32)
        self._check(self.lib.wb_dev_ldpc_result(self.h, first, n, _ptr(packed), _ptr(iters), _ptr(pcc)))
        return np.unpackbits(packed, axis=1)[:, :NCODE], iters, pcc

    def timer_start(self):
        self._check(self.lib.wb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self.lib.wb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        ms = np.zeros(4, dtype=np.float32)
        self._check(self.lib.wb_last_kernel_ms(self.h, _ptr(ms)))
        return ms

    @property
    def launch_count(self):
        return int(self.lib.wb_launch_count(self.h))

    @property
    def last_codewords(self):
        return int(self.lib.wb_last_codewords(self.h))

    @property
    def last_samples(self):
        return int(self.lib.wb_last_samples(self.h))
