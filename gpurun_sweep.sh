timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for lm in 100 50; do for w in 5 10; do
  echo "GROUP_LM=$lm WARPS=$w"
  VIO_B200_GROUP_LM=$lm VIO_B200_GROUP_WARPS=$w python bench.py --steps 4 --warmup 3 --no-cpu --pcg-max-iter 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   kernel_ms', d['roofline']['kernel_ms'], 'fp64 frac', d['roofline']['fp64']['frac'])"
done; done
VIO_B200_GROUP_LM=100 VIO_B200_GROUP_WARPS=10 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('full bench: ms/step', d['ms_per_step'], 'pcg iters', d['lm']['pcg_iterations'], 'value', d['value'], 'chi', d['lm']['chi2_final'])"
