#!/usr/bin/env python3
"""Turn the outputs of tools/run_gpu_check.sh + tools/run_gpu_profile.sh (gpurun_out/<tag>_*) into the tracked summaries
under profiles/:

    python tools/profile_digest.py r02i r02

writes profiles/<out>_launches.csv (copy), <out>_launches_summary.txt, <out>_traffic.json (what bench.py's roofline reads),
<out>_ncu_summary.txt (tools/ncu_summary.py), <out>_fsk_instruction_mix.txt (tools/ncu_lines.py), <out>_sanitizer.txt,
<out>_bench_final.json and <out>_gpu_pytest_final.log.  The library the report was captured from must still be the one in
the tree (ncu_lines.py matches the SASS instruction by instruction)."""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
G = lambda name: os.path.join(ROOT, "gpurun_out", "%s_%s" % (tag, name))
P = lambda name: os.path.join(ROOT, "profiles", "%s_%s" % (out, name))
HOT = ("wb_fsk_kernel", "wb_deframe_kernel", "wb_llr_stats_kernel", "wb_ldpc_kernel", "wb_carry_kernel")


def short(name):
    name = name.replace("(int)", "").replace("(bool)", "")
    return name.split("(")[0]


def launches():
    rows = [r for r in csv.reader(open(G("launches.csv"))) if len(r) > 14 and r[0].isdigit()]
    seq = [(short(r[4]), r[8], float(r[14]) / 1e6) for r in rows]
    shutil.copy(G("launches.csv"), P("launches.csv"))
    per = defaultdict(lambda: [0, 0.0])
    for k, _, ms in seq:
        per[k][0] += 1
        per[k][1] += ms
    # steps of the headline workload: fsk<2,8,1,1,0> launches with the full grid, each followed by deframe/llr/ldpc(/carry)
    grid = max((g for k, g, _ in seq if k.startswith("void wb_fsk_kernel<2, 8, 1, 1, 0>")), key=lambda g: int(g.strip("()").split(",")[0]))
    steps = []
    for i, (k, g, ms) in enumerate(seq):
        if k.startswith("void wb_fsk_kernel<2, 8, 1, 1, 0>") and g == grid:
            st = OrderedDict(fsk=ms)
            for k2, _, ms2 in seq[i + 1:i + 8]:
                for nm, key in (("wb_deframe_kernel", "deframe"), ("wb_llr_stats_kernel", "llr_stats"), ("wb_ldpc_kernel", "ldpc"), ("wb_carry_kernel", "carry")):
                    if k2 == nm and key not in st:
                        st[key] = ms2
                if k2.startswith("void wb_fsk_kernel"):
                    break
            if "ldpc" in st:
                steps.append(st)
    with open(P("launches_summary.txt"), "w") as f:
        f.write("ncu launch list of `python bench.py --steps 2 --warmup 1` (the default bench command, 1 x B200, %s, final kernels):\n" % out)
        f.write("gpu__time_duration per launch, --clock-control none; times under ncu are serialised and cold-cache, shares are what counts.\n\n")
        f.write("headline workload (4096 streams x 1 Mi samples, cf32, v1 framing; FSK grid %s): %d steps seen\n" % (grid, len(steps)))
        for st in steps:
            tot = sum(st.values())
            f.write("  " + "  ".join("%s %.3f" % kv for kv in st.items()) + "  | step %.3f ms, fsk share %.1f %%\n" % (tot, 100 * st["fsk"] / tot))
        f.write("\nall launches by kernel (the command also runs the parity pass, the e2e legs with cu8 / cs16 / cf32 host buffers and the `extra` configurations):\n")
        for k, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("  %-44s launches %4d  total %10.3f ms\n" % (k, n, ms))
        foreign = [k for k in per if not (k.startswith("wb_") or k.startswith("void wb_"))]
        f.write("\n" + ("no library kernels: every launch above is one of libwenet_b200.so's own (torch is not imported at N = 1).\n" if not foreign
                        else "launches that are not libwenet_b200.so's: %s\n" % foreign))
    return steps


def traffic():
    outp = subprocess.run(["ncu", "-i", G("full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(outp.splitlines()))
    hdr, units = rows[0], rows[1]
    bench = json.loads(open(G("bench.json")).read().strip().splitlines()[-1])
    wl = {"streams": bench["config"]["streams_per_gpu"], "chunk_samples": bench["config"]["chunk_samples"], "in_fmt": bench["config"]["in_fmt"]}
    res = OrderedDict(source="ncu --set full --clock-control none --import-source on, second launch of each kernel of `python bench.py --steps 1 "
                      "--warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity` (4096 streams x 1 Mi samples, 1 x B200), round 2, final kernels "
                      "(tools/run_gpu_profile.sh + tools/profile_digest.py; summary profiles/%s_ncu_summary.txt, per-phase instruction mix "
                      "profiles/%s_fsk_instruction_mix.txt)" % (out, out))

    def num(d, k, scale_by_unit=True):
        v = d.get(k, "")
        if v in ("", "n/a"):
            return None
        x = float(v.replace(",", ""))
        u = units[hdr.index(k)]
        if scale_by_unit:
            x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        return x

    seen = set()
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        nm = short(d["Kernel Name"]).replace("void ", "").split("<")[0]
        if nm in seen or nm not in HOT:
            continue
        seen.add(nm)
        e = OrderedDict()
        e["dram_bytes_read"] = num(d, "dram__bytes_read.sum")
        e["dram_bytes_write"] = num(d, "dram__bytes_write.sum")
        e["workload"] = wl
        e["issue_active_pct"] = round(num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active", False), 1)
        e["inst_executed"] = int(num(d, "smsp__inst_executed.sum", False))
        sw = num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False)
        bc = num(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False)
        if sw:
            e["smem_wavefronts"] = int(sw)
        if bc is not None:
            e["smem_bank_conflicts"] = int(bc)
        e["gpu_time_ms"] = round(num(d, "gpu__time_duration.sum"), 3)
        if nm == "wb_fsk_kernel":
            e["warp_inst_per_sample"] = round(e["inst_executed"] / float(bench["work_per_step_per_gpu"]["samples"]), 2)
        if nm == "wb_ldpc_kernel":
            e["codewords"] = bench["work_per_step_per_gpu"]["codewords"]
        res[nm] = e
    json.dump(res, open(P("traffic.json"), "w"), indent=1)
    return res


def text_tools():
    with open(P("ncu_summary.txt"), "w") as f:
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), G("full.ncu-rep")], capture_output=True, text=True).stdout)
    tmp = "/tmp/profile_digest"
    os.makedirs(tmp, exist_ok=True)
    sass_csv = os.path.join(tmp, "fsk_sass.csv")
    with open(sass_csv, "w") as f:
        f.write(subprocess.run(["ncu", "-i", G("full.ncu-rep"), "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:wb_fsk"],
                               capture_output=True, text=True).stdout)
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "wenet_b200", "libwenet_b200.so")], cwd=tmp, check=True, capture_output=True)
    cubin = [x for x in os.listdir(tmp) if x.endswith(".cubin")][0]
    with open(os.path.join(tmp, "all.sass"), "w") as f:
        f.write(subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), sass_csv, os.path.join(tmp, "all.sass"), "_Z13wb_fsk_kernelILi2ELi8ELb1ELb1ELb0EE",
                        str(json.loads(open(G("bench.json")).read().strip().splitlines()[-1])["work_per_step_per_gpu"]["samples"])],
                       capture_output=True, text=True)
    with open(P("fsk_instruction_mix.txt"), "w") as f:
        f.write(r.stdout)
    if r.returncode:
        print("ncu_lines.py failed:", r.stderr[-400:])


def small_files():
    shutil.copy(G("bench.json"), P("bench_final.json"))
    shutil.copy(G("pytest.log"), P("gpu_pytest_final.log"))
    # the sanitizer note keeps its prose; only the two result lines are refreshed
    res = {}
    for tool in ("memcheck", "racecheck"):
        lines = [l.strip() for l in open(G("san_%s.log" % tool)) if "SUMMARY" in l]
        res[tool] = lines[-1] if lines else "no summary line (see gpurun_out/%s_san_%s.log)" % (tag, tool)
    if os.path.exists(P("sanitizer.txt")):
        txt = open(P("sanitizer.txt")).read().split("\n")
        txt = [("%s:%s%s" % (l.split(":")[0], " " * (11 - len(l.split(":")[0]) - 1), res[l.split(":")[0]]) if l.split(":")[0] in res else l) for l in txt]
    else:
        txt = ["compute-sanitizer over tools/sanitize_smoke.py, 1 x B200 (tools/run_gpu_profile.sh):"] + ["%s: %s" % kv for kv in res.items()]
    open(P("sanitizer.txt"), "w").write("\n".join(txt))


if __name__ == "__main__":
    st = launches()
    tr = traffic()
    text_tools()
    small_files()
    print("steps in launch list:", len(st), "| fsk full-capture ms", tr["wb_fsk_kernel"]["gpu_time_ms"], "issue", tr["wb_fsk_kernel"]["issue_active_pct"],
          "inst/sample", tr["wb_fsk_kernel"]["warp_inst_per_sample"])
