set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q > gpurun_out/r2_test_epi_ops.log 2>&1; tail -2 gpurun_out/r2_test_epi_ops.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "unet or generate_tiny or full_unet or ragged or batch4_unet or vae or 192 or boundary" > gpurun_out/r2_test_epi_model.log 2>&1; tail -2 gpurun_out/r2_test_epi_model.log
for i in 1 2; do
(cd .ab_prev && python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > ../gpurun_out/ab_epi4_$i.json 2>/dev/null)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/ab_epi8_$i.json 2>/dev/null
done
python tools/profile_ops.py 4 gpurun_out/r2_ops_b4_epi8.json > gpurun_out/r2_ops_b4_epi8.txt 2>&1
(cd .ab_prev && python tools/profile_ops.py 4 ../gpurun_out/r2_ops_b4_epi4.json > ../gpurun_out/r2_ops_b4_epi4.txt 2>&1)
for f in ab_epi4_1 ab_epi8_1 ab_epi4_2 ab_epi8_2; do python -c "import json,sys; d=json.loads(open(\"gpurun_out/$f.json\").read().strip().splitlines()[-1]); print(\"$f\", d[\"ms_per_step\"], d[\"config\"][\"unet_step_ms_batch4\"], d[\"clocks\"][\"sm_mhz\"], d[\"roofline\"][\"ms\"])"; done
