python -m pytest tests/test_gpu_r02.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r02o_pytest.log
tools/deck_times.sh 2>&1 | tee gpurun_out/r02o_decks.log
