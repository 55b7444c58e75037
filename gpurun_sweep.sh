timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
VIO_B200_PROFILE=1 python bench.py --steps 4 --warmup 3 --no-cpu --pcg-max-iter 3 2> gpurun_out/p.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   kernel_ms', d['roofline']['kernel_ms'], 'fp64 frac', d['roofline']['fp64']['frac'])"
grep "linearise phases" gpurun_out/p.err | tail -1
