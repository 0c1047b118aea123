#!/bin/bash
# A/B timing of two builds of the library on one box: the shipped libwenet_b200.so against $1 (default libwenet_b200_exp.so)
EXP=${1:-wenet_b200/libwenet_b200_exp.so}
run() { python bench.py --steps 6 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['kernel_ms'])"; }
for i in 1 2; do
  run shipped
  WB_LIBRARY=$PWD/$EXP run experiment
done
