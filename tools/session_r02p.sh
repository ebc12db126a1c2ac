python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r02p_pytest.log
tools/deck_times.sh 2>&1 | tee gpurun_out/r02p_decks.log
tools/deck_times.sh LBM_B200_CLUSTER_ROWS=0 2>&1 | head -2 | tee -a gpurun_out/r02p_decks.log
