set -x
python bench.py --workload config1 --steps 5 --warmup 3 > gpurun_out/r2_config1.json 2> gpurun_out/r2_config1.err
python bench.py --workload config3 --steps 3 --warmup 3 > gpurun_out/r2_config3.json 2> gpurun_out/r2_config3.err
python tools/profile_ops.py 1 gpurun_out/r2_ops_b1.json > gpurun_out/r2_ops_b1.txt 2>&1
tail -c 500 gpurun_out/r2_config1.json; tail -c 500 gpurun_out/r2_config3.json; head -1 gpurun_out/r2_ops_b1.txt
