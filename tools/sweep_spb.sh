#!/bin/bash
# Streams per CTA of wb_fsk_kernel for stream counts that do not fill the GPU (bench.py --mode fskonly; WB_FSK_SPB
# overrides wb_create's choice, "auto" = no override).  Output: Msamples/s, ms per step, per-kernel ms.
mkdir -p gpurun_out
out=gpurun_out/${SPB_LOG:-spb_sweep.log}
cases=("1024 auto" "1024 14" "1024 8" "1024 7" "1024 6" "512 auto" "512 14" "512 4" "512 3" "256 auto" "256 2" "2048 auto")
for cfg in "${cases[@]}"; do
  set -- $cfg
  echo "streams=$1 spb=$2" >> $out
  if [ "$2" = auto ]; then unset WB_FSK_SPB; else export WB_FSK_SPB=$2; fi
  timeout 120 python bench.py --mode fskonly --streams $1 --sources 8 --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" >> $out 2>&1
done
cat $out
