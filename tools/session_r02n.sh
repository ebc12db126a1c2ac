python -m pytest tests/test_gpu_r02.py -m gpu -q -x -k "cluster" 2>&1 | tail -12 | tee gpurun_out/r02n_pytest_cluster.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_r02.py -m gpu -q -x -k "k_steps_ring and 2-256-16-64-3-3" > gpurun_out/r02n_sanitizer.log 2>&1
grep -v "^$" gpurun_out/r02n_sanitizer.log | head -60
tools/deck_times.sh 2>&1 | tee gpurun_out/r02n_decks.log
