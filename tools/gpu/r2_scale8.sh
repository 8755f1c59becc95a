set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
$TR --master-port 29522 bench.py --gpus 8 --workload config4 --steps 1 --warmup 0 > gpurun_out/r2_bench_8gpu_config4.json 2> gpurun_out/r2_bench_8gpu_config4.err
$TR --master-port 29523 bench.py --gpus 8 --workload config5 --steps 1 --warmup 0 > gpurun_out/r2_bench_8gpu_config5.json 2> gpurun_out/r2_bench_8gpu_config5.err
for f in r2_bench_8gpu r2_bench_8gpu_config4 r2_bench_8gpu_config5; do tail -c 700 gpurun_out/$f.json; echo; tail -3 gpurun_out/$f.err; done
