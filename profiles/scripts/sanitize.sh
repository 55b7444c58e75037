export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -x -q -m gpu -k "graph_replay or block_cyclic_reduction_is_exact or schur_and_step or window_solve or fixed_landmarks or xyz_solve" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck.txt
timeout 280 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests -x -q -m gpu -k "schur_and_step or window_solve_vs_golden or fixed_landmarks or (graph_replay and monoba) or (block_cyclic_reduction_is_exact and 97)" > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_racecheck.txt
