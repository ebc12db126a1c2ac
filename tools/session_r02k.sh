python -m pytest tests/test_gpu_r02.py -m gpu -q -x -k "k_steps_per_pass" 2>&1 | tail -8 | tee gpurun_out/r02k_pytest.log
tools/sweep_r02.sh gpurun_out/r02k "--fused-steps 3 --fused-deep 1" "--fused-steps 3 --fused-deep 0" "--fused-steps 4 --fused-deep 1" "--fused-steps 4 --fused-deep 0" \
  "--fused-steps 3 --fused-deep 0 --band-rows 64" "--fused-steps 3 --fused-deep 0 --band-rows 128" "--fused-steps 3 --fused-deep 0 --band-rows 256" \
  "--fused-steps 4 --fused-deep 0 --band-rows 128" "--fused-steps 4 --fused-deep 0 --band-rows 256"
