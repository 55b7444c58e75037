export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
( time python bench.py ) > gpurun_out/r02_bench_config5.json 2> gpurun_out/r02_bench_config5.err; tail -4 gpurun_out/r02_bench_config5.err
( time python bench.py --impl reference ) > gpurun_out/r02_bench_config5_reference.json 2> gpurun_out/r02_bench_config5_reference.err; tail -4 gpurun_out/r02_bench_config5_reference.err
for w in config4_ba_1k_cams_100k_landmarks_1m_obs config1_monoba_20x300_v15 config1_monoba_20x300_v17 config2_vins_window config3_batched_4096_windows; do
python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"
done
python __graft_entry__.py --smoke 2>&1 | tail -2
