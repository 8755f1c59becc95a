set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python tools/profile_ops.py 4 gpurun_out/r2_ops_b4.json > gpurun_out/r2_ops_b4.txt 2>&1
python tools/ncu_generate.py > gpurun_out/ncu_gen_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/ncu_generate.py > gpurun_out/ncu_gen.log 2>&1
python tools/ncu_target.py > gpurun_out/ncu_target_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_t_kernel -s 1 -c 2 -o gpurun_out/r2_prof_t0_unet python tools/ncu_target.py > gpurun_out/ncu_t0a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_t_kernel -s 27 -c 2 -o gpurun_out/r2_prof_t0_dec python tools/ncu_target.py > gpurun_out/ncu_t0b.log 2>&1
ncu --set full --clock-control none -k regex:"gn_apply_kernel|gn_res_tsum|attn_gemm|add_bcast" -s 4 -c 24 -o gpurun_out/r2_prof_hbm python tools/ncu_target.py > gpurun_out/ncu_hbm.log 2>&1
tail -2 gpurun_out/ncu_gen.log gpurun_out/ncu_t0a.log gpurun_out/ncu_t0b.log gpurun_out/ncu_hbm.log
ls -la gpurun_out/*.ncu-rep
