#!/usr/bin/env python3
"""Join an ncu SASS source page with nvdisasm's line table: instructions executed per source line, per phase of
wb_fsk_kernel and per opcode class.

    ncu -i REP --page source --csv --kernel-name regex:wb_fsk > fsk_sass.csv
    cuobjdump -xelf all wenet_b200/libwenet_b200.so ; nvdisasm -gi -c wb_engine.sm_100a.cubin > all.sass
    python tools/ncu_lines.py fsk_sass.csv all.sass _Z13wb_fsk_kernelILi2ELi8ELb1ELb1ELb0EE [samples_per_launch]

The two listings are matched by instruction order (the report must come from the same build of the kernel: the tool
checks the instruction count and every opcode).  Phases are line ranges of wenet_b200/csrc/wb_fsk_kernel.cuh found from
the phase banner comments, so the split follows the source as it is edited.
"""
import csv
import re
import sys
from collections import defaultdict


def read_ncu(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = {n: i for i, n in enumerate(rows[hi])}
    out = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":       # a second launch of the kernel in the same report: the first one is enough
            break
        if len(r) < len(h):
            continue
        out.append(dict(sass=r[h["Source"]].strip(), inst=int(r[h["Instructions Executed"]]),
                        thr=int(r[h["Thread Instructions Executed"]]),
                        samples=int(r[h["# Samples"]] or 0),
                        wave=int(r[h["L1 Wavefronts Shared"]] or 0), wave_ideal=int(r[h["L1 Wavefronts Shared Ideal"]] or 0),
                        stalls={k[6:]: int(r[i] or 0) for k, i in h.items() if k.startswith("stall_") and "Not Issued" not in k}))
    return out


def read_disasm(path, func):
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + func))
    # with -gi an instruction of an inlined helper carries a chain of comments, innermost first and the call site in the
    # kernel body last: keep (innermost, outermost)
    out, inner, outer, fresh = [], ("?", 0), ("?", 0), True
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            if fresh:
                inner, fresh = cur, False
            outer = cur
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            out.append((inner, outer, m.group(2).strip()))
            fresh = True
    return out


def opclass(sass):
    t = sass.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    op = t[0].split(".")[0] if t else "?"
    if op in ("FADD", "FMUL", "FFMA", "FSETP", "FMNMX", "FSEL", "FCHK", "FADD2", "FMUL2", "FFMA2"):
        return "fp32"
    if op in ("MUFU",):
        return "sfu"
    if op in ("LDS", "STS", "LDSM"):
        return "smem"
    if op in ("LDG", "STG", "LDL", "STL", "LDGSTS", "LD", "ST", "ATOM", "RED", "CCTL", "UBLKCP", "UTMALDG", "LDGDEPBAR", "DEPBAR"):
        return "gmem/local"
    if op in ("LDC", "ULDC", "LDCU"):
        return "const"
    if op in ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "BAR", "NANOSLEEP", "YIELD", "BREAK", "JMP", "B2R", "SYNCS"):
        return "control"
    if op in ("SHFL", "VOTE", "REDUX", "MATCH"):
        return "warp"
    if op in ("LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "FENCE"):
        return "tmem"
    if op in ("DADD", "DMUL", "DFMA", "DSETP", "F2F", "I2F", "F2I", "I2FP", "F2FP", "FRND"):
        return "cvt/fp64"
    return "int/move"


def main():
    ncu, dis, func = sys.argv[1], sys.argv[2], sys.argv[3]
    samples = float(sys.argv[4]) if len(sys.argv) > 4 else None
    a, b = read_ncu(ncu), read_disasm(dis, func)
    if len(a) != len(b):
        sys.exit("instruction counts differ: ncu %d, nvdisasm %d (different build?)" % (len(a), len(b)))
    for x, (_, _, s) in zip(a, b):
        if x["sass"].split()[0:1] != s.split()[0:1] and x["sass"].split()[1:2] != s.split()[1:2]:
            sys.exit("listings diverge at %r vs %r" % (x["sass"], s))
    # phase boundaries from the banner comments of the kernel source
    src = open("wenet_b200/csrc/wb_fsk_kernel.cuh").read().split("\n")
    marks = []
    for i, l in enumerate(src, 1):
        m = re.search(r"/\* =+ (A|B1|B2|B3|C)\b", l)
        if m:
            marks.append((i, m.group(1)))
        if "---- write the state back ----" in l:
            marks.append((i, "epilogue"))
        if re.search(r"^\s*for \(;;\) \{", l):
            marks.append((i, "loop-head"))
    marks.sort()

    def phase_of(f, ln):
        if f != "wb_fsk_kernel.cuh":
            return None
        ph = "prologue/helpers"
        for i, n in marks:
            if ln >= i:
                ph = n
        return ph

    tot = sum(x["inst"] for x in a)
    per_line, per_phase, per_class, per_phase_class = defaultdict(int), defaultdict(int), defaultdict(int), defaultdict(lambda: defaultdict(int))
    smp_phase, wave_phase = defaultdict(int), defaultdict(lambda: [0, 0])
    for x, ((f, ln), (fo, lno), s) in zip(a, b):
        ph = phase_of(fo, lno) or "prologue/helpers"      # the phase of the call site in the kernel body
        per_line[(f, ln)] += x["inst"]
        per_phase[ph] += x["inst"]
        c = opclass(x["sass"])
        per_class[c] += x["inst"]
        per_phase_class[ph][c] += x["inst"]
        smp_phase[ph] += x["samples"]
        wave_phase[ph][0] += x["wave"]
        wave_phase[ph][1] += x["wave_ideal"]
    print("kernel %s: %d SASS instructions, %.3f G warp-instructions executed" % (func, len(a), tot / 1e9))
    if samples:
        print("  = %.2f warp-instructions per IQ sample (%.0f samples per launch)" % (tot / samples, samples))
    print("\nper opcode class:")
    for c, n in sorted(per_class.items(), key=lambda t: -t[1]):
        print("  %-12s %8.3f G  %5.1f %%" % (c, n / 1e9, 100.0 * n / tot))
    print("\nper phase (instructions, share, warp-inst per sample, stall samples share, smem wavefronts actual/ideal):")
    ts = sum(smp_phase.values()) or 1
    for ph, n in sorted(per_phase.items(), key=lambda t: -t[1]):
        w = wave_phase[ph]
        print("  %-18s %8.3f G  %5.1f %%  %s  samples %5.1f %%  waves %.2f/%.2f G" % (
            ph, n / 1e9, 100.0 * n / tot, ("%.2f/sample" % (n / samples)) if samples else "", 100.0 * smp_phase[ph] / ts, w[0] / 1e9, w[1] / 1e9))
        cl = per_phase_class[ph]
        print("      " + "  ".join("%s %.1f%%" % (c, 100.0 * v / n) for c, v in sorted(cl.items(), key=lambda t: -t[1]) if v * 50 > n))
    print("\ntop 25 source lines by instructions executed:")
    for (f, ln), n in sorted(per_line.items(), key=lambda t: -t[1])[:25]:
        text = src[ln - 1].strip()[:100] if f == "wb_fsk_kernel.cuh" and ln <= len(src) else ""
        print("  %5.2f %%  %s:%d  %s" % (100.0 * n / tot, f, ln, text))


if __name__ == "__main__":
    main()
