for cfg in "50 10" "50 5" "34 10" "25 10"; do set -- $cfg
  echo "GROUP_LM=$1 WARPS=$2"
  VIO_B200_GROUP_LM=$1 VIO_B200_GROUP_WARPS=$2 python bench.py --steps 4 --warmup 3 --no-cpu --pcg-max-iter 3 2> gpurun_out/p.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   kernel_ms', d['roofline']['kernel_ms'], 'fp64 frac', d['roofline']['fp64']['frac'])"
done
