export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
run() { python bench.py --workload $2 --steps $3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 $2', 'ms/step', round(d['ms_per_step'],4), d.get('lm'))"; }
run x config2_vins_window 20
run x config1_monoba_20x300_v17 20
run x config5_ba_10k_cams_1m_landmarks_10m_obs 10
run x config4_ba_1k_cams_100k_landmarks_1m_obs 10
