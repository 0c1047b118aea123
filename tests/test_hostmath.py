"""Scalar numeric building blocks of the CUDA kernels (wenet_b200/csrc/wb_math.h, wb_phi0.h), compiled for the
host as libwb_hostmath.so: they must equal what the reference's x86-64 build computes -- glibc atan2f, the x87
long-double expressions of sd_to_llr (src/mpdecode_core.c:593-595) and the phi0 compare tree (src/phi0.c)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "wenet_b200", "libwb_hostmath.so")


@pytest.fixture(scope="module")
def hm():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "wenet_b200", "csrc"), "../libwb_hostmath.so"])
    return C.CDLL(SO)


def _call(fn, *arrays_in, out_dtype):
    n = arrays_in[0].size
    out = np.empty(n, dtype=out_dtype)
    fn(*[a.ctypes.data_as(C.c_void_p) for a in arrays_in], out.ctypes.data_as(C.c_void_p), C.c_long(n))
    return out


def test_atan2f_equals_glibc(hm):
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.standard_normal(200000) * 10 ** rng.uniform(-6, 6, 200000),
                        [0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-38, 3e38]]).astype(np.float32)
    x = np.concatenate([rng.standard_normal(200000) * 10 ** rng.uniform(-6, 6, 200000),
                        [1.0, -1.0, 0.0, -0.0, np.inf, 1.0, 1.0, -3e38, 1e-38]]).astype(np.float32)
    got = _call(hm.wbh_atan2f, y, x, out_dtype=np.float32)
    ref = np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)
    same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
    assert same.all(), np.nonzero(~same)[0][:5]


def test_x87_emulation(hm):
    rng = np.random.default_rng(1)
    v = np.concatenate([10 ** rng.uniform(-9, 3, 300000), [0.0, 1e-3, 0.5, 2.0 ** -20]])
    a = _call(hm.wbh_esn0_from_var, v, out_dtype=np.float64)
    b = _call(hm.wbh_esn0_from_var_x87, v, out_dtype=np.float64)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    c = 4.0 * 10 ** rng.uniform(-3, 3, 300000)
    sd = (rng.standard_normal(300000) * 10 ** rng.uniform(-4, 2, 300000)).astype(np.float32)
    a = _call(hm.wbh_llr_scale, c, sd, out_dtype=np.float32)
    b = _call(hm.wbh_llr_scale_x87, c, sd, out_dtype=np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # the one-multiplication form the decoder uses (exact emulation only next to a float rounding midpoint)
    f = _call(hm.wbh_llr_scale_fast, c, sd, out_dtype=np.float32)
    assert np.array_equal(f.view(np.uint32), b.view(np.uint32))


def test_llr_scale_fast_next_to_float_midpoints(hm):
    """products engineered to land on, and within a few double ulps of, the midpoint between two floats -- where rounding
    the exact product to 64 bits (the reference's x87 fmul) or to 53 bits first can decide the float differently -- and
    where the 53-bit product is itself a tie: the fast form must agree with genuine long double arithmetic on all"""
    rng = np.random.default_rng(7)
    n = 400000
    sd = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 1, n)).astype(np.float32)
    sd[sd == 0] = np.float32(1.0)
    # target: a float midpoint t = (k + 0.5) ulp; c = t / sd rounded, then nudged by -3..3 double ulps
    k = rng.integers(1 << 23, 1 << 24, n).astype(np.float64)
    t = (k + 0.5) * 2.0 ** rng.integers(-30, 4, n)
    c = np.abs(t / sd.astype(np.float64))
    c = (c.view(np.uint64) + rng.integers(-3, 4, n).astype(np.uint64)).view(np.float64)
    ref = _call(hm.wbh_llr_scale_x87, c, sd, out_dtype=np.float32)
    got = _call(hm.wbh_llr_scale_fast, c, sd, out_dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # how many of them actually sat in the guarded window (the test is vacuous if none did)
    p = c * sd.astype(np.float64)
    low = p.view(np.uint64) & np.uint64(0x1fffffff)
    assert np.count_nonzero((low >= 0x0fffffff) & (low <= 0x10000001)) > 1000


def test_phi0_tables(hm, oracle_port):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-1, 40, 500000), 2.0 ** rng.uniform(-20, 17, 500000),
                        [0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, 32768, 32767.99, 1e9, 10, 9.9999, 16, 15.999999]]).astype(np.float32)
    # every breakpoint of the step function and its float neighbours
    brk = np.float32(2.0) ** np.arange(-15, 5, dtype=np.float32)
    grid = np.concatenate([brk, brk * np.float32(2 ** -0.5), np.arange(1, 10.5, 1 / 16, dtype=np.float32)])
    near = np.concatenate([np.nextafter(grid, np.float32(0)), grid, np.nextafter(grid, np.float32(100))])
    q = (np.arange(0, 70000, dtype=np.float64) / 65536.0).astype(np.float32)          # every Q16 integer below 1.07
    x = np.concatenate([x, near, q, np.nextafter(q, np.float32(0)), np.nextafter(q, np.float32(100))]).astype(np.float32)
    ref = oracle_port.phi0(x)
    for name in ("wbh_phi0", "wbh_phi0_pairs", "wbh_phi0_flag"):
        fn = getattr(hm, name)
        fn.restype = C.c_long
        got = np.full(x.size, -1.0, dtype=np.float32)
        rc = fn(x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_long(x.size))
        assert name == "wbh_phi0" or rc == 0, (name, rc)      # the table builders report a layout they cannot express
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), name
    assert math.isclose(float(oracle_port.phi0(np.float32([0.5]))[0]), 1.5735153, rel_tol=1e-7)


@pytest.mark.parametrize("P,Rs", [(8, 115177), (10, 96000)])
def test_fine_timing_oscillator_is_periodic(hm, P, Rs):
    """the reference's float recurrence for phi_ft (src/fsk.c:858-873) falls into an exactly periodic orbit of period P
    within two symbols: wb_fsk_kernel keeps one period of multipliers in registers from block WB_PFT_NT = 2 on
    (wb_create re-checks this on the table it builds and falls back to the table walk otherwise)"""
    n = 49 * P
    re, im = np.empty(n, np.float32), np.empty(n, np.float32)
    hm.wbh_pft_table(C.c_int(P), C.c_int(Rs), C.c_int(n), re.ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p))
    assert re[0] == 1.0 and im[0] == 0.0
    nt = 2
    assert np.array_equal(re[(nt + 1) * P:].view(np.uint32), re[nt * P:-P].view(np.uint32))
    assert np.array_equal(im[(nt + 1) * P:].view(np.uint32), im[nt * P:-P].view(np.uint32))
    assert not np.array_equal(re[P:2 * P].view(np.uint32), re[:P].view(np.uint32))      # the transient is real


def test_division_shortcuts(hm):
    """the kernel multiplies by 1/(2 pi) and 1/48 where the reference divides (src/fsk.c:883,888: norm_rx_timing and the
    ppm update): identical floats for every argument those divisions can ever see, checked exhaustively (2 x 10^9 floats)"""
    hm.wbh_check_div_2pi.restype = C.c_long
    hm.wbh_check_div_48.restype = C.c_long
    assert hm.wbh_check_div_2pi() == 0
    assert hm.wbh_check_div_48() == 0

