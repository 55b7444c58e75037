timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lockstep or batched" 2>&1 | tail -3
VIO_B200_PROFILE=1 timeout 300 python tests/bench_batched.py --n 1024 --lockstep --cpu-n 1 2>&1 | grep -v "linearise phases" | tail -9
timeout 300 python tests/bench_batched.py --n 4096 --lockstep --cpu-n 1 --chunk 2048 2>&1 | tail -1
