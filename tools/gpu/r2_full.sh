set -x
python -m pytest tests -m gpu -q -rP > gpurun_out/r2_test_full.log 2>&1; tail -3 gpurun_out/r2_test_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python tools/ncu_target.py > gpurun_out/ncu_target_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 2 -c 10 -o gpurun_out/r2_prof_pair python tools/ncu_target.py > gpurun_out/ncu_pair.log 2>&1
ls -la gpurun_out/r2_prof_pair.ncu-rep
