#!/bin/bash
# quick 1-GPU check while iterating on a kernel: GPU tests, then the headline timing only (parity gate on)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "parity", {k: v for k, v in d["parity"].items() if k.endswith("equal")})
except Exception as e:
    print("no bench line:", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
