set -x
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2_bench_plain.json; tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/r2_launches_bench.csv
