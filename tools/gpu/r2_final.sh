set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_test_final.log 2>&1; tail -3 gpurun_out/r2_test_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -1 gpurun_out/r2_smoke_final.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref_full.json 2> gpurun_out/r2_bench_ref_full.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 400 gpurun_out/r2_bench_ref_full.json; tail -c 300 gpurun_out/r2_bench_final.json
