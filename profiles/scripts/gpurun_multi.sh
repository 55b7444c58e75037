export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-p2p}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c5_n${N}_$TAG.json 2> gpurun_out/bench_c5_n${N}_$TAG.err
echo "rc=$?"; tail -3 gpurun_out/bench_c5_n${N}_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c5_n${N}_$TAG.json'))
print('N=$N $TAG ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'lin ms', d['roofline_linearize']['kernel_ms'], 'solve ms', d['roofline_reduced_solve']['kernel_ms'], 'parity', d.get('parity_vs_n1',{}).get('ok'), d['config']['collective'])
PY
