export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config5.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_b.log 2>&1
for k in k_lin_edges k_schur_groups k_bcr_run k_backsub k_chi2_lm; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r02_$k python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_$k.log 2>&1
ncu -i gpurun_out/r02_$k.ncu-rep --page details > gpurun_out/r02_${k}_details.txt 2>&1
ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>&1
if [ $k != k_lin_edges ] && [ $k != k_schur_groups ]; then rm -f gpurun_out/r02_$k.ncu-rep; fi
done
ls -la gpurun_out/ | tail -20
