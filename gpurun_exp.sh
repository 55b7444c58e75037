export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests -x -q -m gpu -k "block_cyclic_reduction_is_exact and 97" > gpurun_out/r02_sanitizer_racecheck_bcr.txt 2>&1; echo "racecheck bcr rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck_bcr.txt
python -m pytest tests -x -q -m gpu -k "block_cyclic or config5 or ring or graph_replay" > gpurun_out/t_bcr.log 2>&1; echo "bcr tests rc=$?"; tail -3 gpurun_out/t_bcr.log
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],3), 'lin ms', round(d['roofline_linearize']['kernel_ms'],3), 'solve ms', round(d['roofline_reduced_solve']['kernel_ms'],3), 'chi2', d['lm']['chi2_final'])"
