#!/usr/bin/env python3
"""Generate wenet_b200/csrc/wb_tables.h and tests/golden/*.npz from the compiled reference.

Runs ONLY in the build container (needs oracle/_ref/libwenet_ref.so, i.e. /root/reference).
The outputs are committed, so nothing at test/bench time reads /root/reference.

What is extracted (data, not code):
  * the H2064_516 parity-check tables  (reference src/H2064_516_sparse.h:17-25), re-laid out
    0-based and row-major: hrows[516][12] (check -> data columns, reference H_rows[p + i*516]),
    hcols[2064][3] (data column -> checks, reference H_cols[c + j*2064]);
  * the v2 scramble code (reference src/wenet_scramble.h:22) as 1000 sign bits (1 = -1);
  * the phi0 step function (reference src/phi0.c:13-218), MEASURED by sweeping the compiled
    reference over every Q16 argument 0 .. 10*65536+64 and recording each breakpoint
    (first Q16 integer of a step, step value); 102 steps;
  * the LDPC known-answer vector (reference src/H2064_516_sparse.h:27-33) -> tests/golden/ldpc_kat.npz
    together with what the compiled reference returns for it (iterations, parity count).
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFSO = os.path.join(ROOT, "oracle", "_ref", "libwenet_ref.so")


def main():
    ref = ctypes.CDLL(REFSO)
    ref.ref_H_rows.restype = ctypes.POINTER(ctypes.c_uint16)
    ref.ref_H_cols.restype = ctypes.POINTER(ctypes.c_uint16)
    ref.ref_kat_input.restype = ctypes.POINTER(ctypes.c_float)
    ref.ref_kat_detected.restype = ctypes.POINTER(ctypes.c_char)
    ref.ref_scramble_code.restype = ctypes.POINTER(ctypes.c_double)
    ref.ref_phi0.restype = ctypes.c_float
    ref.ref_phi0.argtypes = [ctypes.c_float]

    NP, NW, ND, CW = 516, 12, 2064, 3
    assert ref.ref_H_rows_len() == NP * NW and ref.ref_H_cols_len() == ND * CW
    hr = np.ctypeslib.as_array(ref.ref_H_rows(), shape=(NP * NW,)).astype(np.int32)
    hc = np.ctypeslib.as_array(ref.ref_H_cols(), shape=(ND * CW,)).astype(np.int32)
    hrows = hr.reshape(NW, NP).T - 1          # [516][12], 0-based data column
    hcols = hc.reshape(CW, ND).T - 1          # [2064][3], 0-based check
    assert hrows.min() >= 0 and hrows.max() < ND and hcols.min() >= 0 and hcols.max() < NP
    # consistency: the two tables describe the same matrix
    s1 = {(p, int(c)) for p in range(NP) for c in hrows[p]}
    s2 = {(int(p), c) for c in range(ND) for p in hcols[c]}
    assert s1 == s2 and len(s1) == NP * NW

    ns = ref.ref_scramble_len()
    scr = np.ctypeslib.as_array(ref.ref_scramble_code(), shape=(ns,)).copy()
    assert ns == 1000 and set(np.unique(scr)) == {-1.0, 1.0}
    scr_neg = (scr < 0).astype(np.uint8)

    # ---- phi0 sweep ----
    n_in = 10 * 65536 + 64
    xs = (np.arange(n_in, dtype=np.float64) / 65536.0).astype(np.float32)   # exact in float32
    ys = np.empty(n_in, dtype=np.float32)
    ref.ref_phi0_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    ref.ref_phi0_array(xs.ctypes.data, ys.ctypes.data, n_in)
    brk = [0] + [int(i) for i in (np.nonzero(ys[1:] != ys[:-1])[0] + 1)]
    vals = [ys[i] for i in brk]
    assert vals[0] == np.float32(10.0) and vals[-1] == np.float32(0.0) and brk[-1] == 10 * 65536
    # spot properties: overflow quirk and negatives (reference src/phi0.c:11,15)
    assert ref.ref_phi0(ctypes.c_float(32768.0)) == 10.0
    assert ref.ref_phi0(ctypes.c_float(32767.0)) == 0.0
    assert ref.ref_phi0(ctypes.c_float(-1.0)) == 10.0
    assert ref.ref_phi0(ctypes.c_float(float("nan"))) == 10.0

    # ---- KAT ----
    assert ref.ref_kat_input_len() == 2580 and ref.ref_kat_detected_len() == 2580
    assert ref.ref_kat_input_elsize() == 4 and ref.ref_kat_detected_elsize() == 1
    kin = np.ctypeslib.as_array(ref.ref_kat_input(), shape=(2580,)).copy()
    kdet = np.frombuffer(ctypes.string_at(ref.ref_kat_detected(), 2580), dtype=np.uint8).copy()
    out = np.zeros(2580, dtype=np.uint8)
    pcc = ctypes.c_int(-1)
    ref.ref_ldpc_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    it = ref.ref_ldpc_decode(kin.ctypes.data, out.ctypes.data, ref.ref_max_iter(), ctypes.byref(pcc))
    assert np.array_equal(out, kdet), "reference does not reproduce its own KAT?"
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ldpc_kat.npz"),
                        llr=kin, detected=kdet, iters=np.int32(it), parity_ok=np.int32(pcc.value),
                        max_iter=np.int32(ref.ref_max_iter()))
    print(f"KAT: iters={it} parityCheckCount={pcc.value} max_iter={ref.ref_max_iter()}")

    # ---- python-side copy of the code tables (used by wenet_b200/siggen.py) ----
    os.makedirs(os.path.join(ROOT, "wenet_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "wenet_b200", "data", "h2064_516.npz"),
                        hrows=hrows.astype(np.uint16), hcols=hcols.astype(np.uint16), scramble_neg=scr_neg)

    # ---- header ----
    def arr(name, ctype, data, per_line=16, fmt="{}"):
        lines = [f"static const {ctype} {name}[{len(data)}] = {{"]
        for i in range(0, len(data), per_line):
            lines.append("  " + ", ".join(fmt.format(v) for v in data[i:i + per_line]) + ",")
        lines.append("};")
        return "\n".join(lines)

    h = []
    h.append("/* GENERATED by tools/gen_tables.py from the compiled reference -- do not edit.\n"
             " * Data only: the H2064_516 code tables (reference src/H2064_516_sparse.h:17-25),\n"
             " * the v2 scramble signs (reference src/wenet_scramble.h:22) and the measured steps\n"
             " * of the phi0 function (reference src/phi0.c:13-218), re-laid out for this engine. */")
    h.append("#ifndef WB_TABLES_H\n#define WB_TABLES_H\n#include <stdint.h>")
    h.append(f"#define WB_NPAR {NP}      /* parity checks (reference NUMBERPARITYBITS) */")
    h.append(f"#define WB_NDATA {ND}    /* systematic bits */")
    h.append(f"#define WB_NCODE {NP + ND}    /* code length (reference CODELENGTH) */")
    h.append(f"#define WB_ROWW {NW}       /* H1 row weight (reference MAX_ROW_WEIGHT) */")
    h.append(f"#define WB_COLW {CW}        /* H1 column weight (reference MAX_COL_WEIGHT) */")
    h.append(f"#define WB_LDPC_MAX_ITER {ref.ref_max_iter()} /* reference MAX_ITER */")
    h.append("/* wb_hrows[p*12+i] = 0-based data column of the i-th H1 entry of check p */")
    h.append(arr("wb_hrows", "uint16_t", hrows.reshape(-1).tolist(), 24))
    h.append("/* wb_hcols[c*3+j] = 0-based check of the j-th H1 entry of data column c */")
    h.append(arr("wb_hcols", "uint16_t", hcols.reshape(-1).tolist(), 24))
    h.append("#define WB_SCRAMBLE_LEN 1000")
    h.append("/* 1 = multiply the soft symbol by -1 */")
    h.append(arr("wb_scramble_neg", "uint8_t", scr_neg.tolist(), 50))
    h.append(f"#define WB_PHI0_NSTEPS {len(brk)}")
    h.append("/* phi0(x) for Q16 argument q = trunc(x*65536): value of the last step with brk <= q;\n"
             "   q < 0 (negative, NaN, or x >= 32768 through the cvttss2si overflow) -> 10.0 */")
    h.append(arr("wb_phi0_brk", "int32_t", brk, 12))
    h.append(arr("wb_phi0_val", "float", [np.format_float_scientific(v, unique=True) for v in vals], 6, "{}f"))
    h.append("#endif")
    path = os.path.join(ROOT, "wenet_b200", "csrc", "wb_tables.h")
    with open(path, "w") as f:
        f.write("\n".join(h) + "\n")
    print("wrote", path, "phi0 steps:", len(brk))


if __name__ == "__main__":
    sys.exit(main())
