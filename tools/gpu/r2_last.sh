set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_test_last.log 2>&1; tail -2 gpurun_out/r2_test_last.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err; tail -c 200 gpurun_out/r2_bench_last.json
