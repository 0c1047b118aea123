# sweep of the mixer phase's redundancy (warps W sharing the oscillator chains, geometric segment ratio R): bench.py headline timing only
for cfg in "4 0.7" "3 0.7" "5 0.7" "6 0.7" "4 0.6" "4 0.8" "5 0.8" "6 0.8" "5 0.6" "8 0.8"; do
  set -- $cfg
  WB_FSK_B1W=$1 WB_FSK_B1R=$2 python bench.py --no-cpu-baseline --no-e2e --no-extra --no-parity --steps 4 --warmup 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('W R = $cfg', d['value'], d['roofline']['kernel_ms'])"
done
