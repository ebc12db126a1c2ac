for d in 0 1 2 4 8 3 7 15; do tools/deck_times.sh LBM_B200_CLUSTER_DEBUG=$d 2>&1 | head -1; done | tee gpurun_out/r02q_cluster_debug.log
