export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
python bench.py --workload config3_batched_4096_windows --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c3 ms/step', round(d['ms_per_step'],5), d['value'], d['e2e']['value'])"
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],3), 'lin ms', round(d['roofline_linearize']['kernel_ms'],3), 'solve ms', round(d['roofline_reduced_solve']['kernel_ms'],3), 'chi2', d['lm']['chi2_final'])"
