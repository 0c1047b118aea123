# start offset between the CTAs that share an SM (WB_FSK_STAGGER=mode,cycles; mode 1: upper half of the grid late, mode 2: odd CTAs late)
LIB=${1:-wenet_b200/libwenet_b200.so}
run() { WB_LIBRARY=$PWD/$LIB python bench.py --steps 6 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['kernel_ms']['fsk'])"; }
unset WB_FSK_STAGGER; run none
for st in ${STAGGERS:-1,8000 1,16000 1,24000 2,8000 2,16000 2,24000 1,4000 1,12000}; do
  WB_FSK_STAGGER=$st run $st
done
