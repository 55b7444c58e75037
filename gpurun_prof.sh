export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "graph_replay" > gpurun_out/t_graph.log 2>&1; echo "graph tests rc=$?"; tail -3 gpurun_out/t_graph.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config5.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_b.log 2>&1
for k in k_lin_edges k_schur_groups k_bcr_run k_backsub k_chi2_lm; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r02_$k python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
