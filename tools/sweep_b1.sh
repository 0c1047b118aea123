set -e
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "8 0.82" "8 0.7" "8 0.9" "4 0.8" "14 0.85" "14 0.75" "1 0.8"; do
  set -- $cfg
  WB_FSK_B1W=$1 WB_FSK_B1R=$2 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['roofline']['kernel_ms'])"
done
