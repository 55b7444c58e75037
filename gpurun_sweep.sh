timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
export VIO_B200_GROUP_LM=100 VIO_B200_GROUP_WARPS=10
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c5_final.json 2> gpurun_out/bench_c5_final.err; tail -2 gpurun_out/bench_c5_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu --pcg-max-iter 200 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_linearize_grouped -s 4 -c 1 -o gpurun_out/prof_lin_grouped_c5 python bench.py --steps 1 --warmup 3 --no-cpu --pcg-max-iter 3 > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none -k regex:k_bpcg_persistent -s 2 -c 1 -o gpurun_out/prof_pcg_c5 python bench.py --steps 1 --warmup 3 --no-cpu --pcg-max-iter 300 > gpurun_out/ncu_p.log 2>&1
ncu --set full --clock-control none -k regex:"k_chi2_lm|k_backsub" -s 4 -c 2 -o gpurun_out/prof_chi2_backsub_c5 python bench.py --steps 1 --warmup 3 --no-cpu --pcg-max-iter 3 > gpurun_out/ncu_c.log 2>&1
cat gpurun_out/bench_c5_final.json | cut -c1-600
