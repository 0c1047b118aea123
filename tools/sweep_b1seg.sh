# sweep of explicit segment ends of the mixer phase (WB_FSK_B1SEG), headline timing only; $1 = library to test
LIB=${1:-wenet_b200/libwenet_b200.so}
run() { WB_LIBRARY=$PWD/$LIB python bench.py --steps 6 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['kernel_ms']['fsk'])"; }
unset WB_FSK_B1SEG; run default
for seg in ${SEGS:-160,272,352 160,280,360 168,288,360 160,280,352 152,272,352 168,280,360 160,288,368 176,296,368}; do
  WB_FSK_B1SEG=$seg run $seg
done
