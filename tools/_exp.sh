python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-e2e --no-cpu-baseline --steps 3 --warmup 2 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['roofline']['kernel_ms'])"
python tools/phase_clk.py 2>&1 | tail -1
