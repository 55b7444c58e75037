export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
python bench.py > gpurun_out/r02_bench_config5.json 2> gpurun_out/r02_bench_config5.err; tail -2 gpurun_out/r02_bench_config5.err
for w in config4_ba_1k_cams_100k_landmarks_1m_obs config1_monoba_20x300_v15 config1_monoba_20x300_v17 config2_vins_window config3_batched_4096_windows; do
python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config5.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_b.log 2>&1
for k in k_bcr_run k_backsub; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r02_$k python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_$k.log 2>&1
ncu -i gpurun_out/r02_$k.ncu-rep --page details > gpurun_out/r02_${k}_ncu_details.txt 2>&1
ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>&1
rm -f gpurun_out/r02_$k.ncu-rep
done
python __graft_entry__.py --smoke 2>&1 | tail -1
