export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
run() { python bench.py --workload $2 --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 $2', 'ms/step', round(d['ms_per_step'],4), d.get('lm'))"; }
for w in config2_vins_window config1_monoba_20x300_v17; do
run t512 $w
VIO_B200_DCH_THREADS=256 run t256 $w
VIO_B200_DCH_THREADS=384 run t384 $w
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l_c2.csv python bench.py --workload config2_vins_window --steps 10 --warmup 3 --no-cpu > /dev/null 2>&1
