python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s3_bench2.log 2>&1; tail -1 gpurun_out/s3_bench2.log
