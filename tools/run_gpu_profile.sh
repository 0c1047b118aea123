#!/bin/bash
# one 1-GPU box call: ncu launch list of the bench command, one --set full capture of the hot kernels, sanitizer passes
TAG=${1:-r02}
mkdir -p gpurun_out
# (1) launch list of the default bench command (durations only; numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launch list rc=$?"
# (2) full capture: second launch of each hot kernel of the headline workload
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'wb_fsk_kernel|wb_ldpc_kernel|wb_llr_stats_kernel|wb_deframe_kernel' \
    --launch-skip 4 -c 4 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity \
    > gpurun_out/${TAG}_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/${TAG}_full.ncu-rep
# (3) sanitizer
for tool in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/${TAG}_san_$tool.log 2>&1
    echo "$tool rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_san_$tool.log | tail -1)"
done
