export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/t_all.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json
d=json.load(open('gpurun_out/final_bench.json'))
print('ms/step', round(d['ms_per_step'],3), d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
