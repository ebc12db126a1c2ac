python -m pytest tests/test_gpu_r02.py -m gpu -q -x -k "k_steps_per_pass" 2>&1 | tail -8 | tee gpurun_out/r02j_pytest.log
tools/sweep_r02.sh gpurun_out/r02j "--fused-steps 2" "--fused-steps 3" "--fused-steps 4" "--fused-steps 3 --band-rows 32" "--fused-steps 3 --band-rows 128" "--fused-steps 3 --band-rows 256" "--fused-steps 4 --band-rows 32" "--fused-steps 4 --band-rows 128" "--fused-steps 4 --band-rows 256" \
  "--nx 4096 --ny 4096 --band-rows 12" "--nx 4096 --ny 4096 --band-rows 24" "--nx 4096 --ny 4096 --band-rows 48" "--nx 4096 --ny 4096 --band-rows 64" "--nx 4096 --ny 4096 --band-rows 96" \
  "--nx 2048 --ny 2048 --band-rows 12" "--nx 2048 --ny 2048 --band-rows 16" "--nx 2048 --ny 2048 --band-rows 24" "--nx 2048 --ny 2048 --band-rows 48" \
  "--nx 16384 --ny 2048 --band-rows 16" "--nx 16384 --ny 2048 --band-rows 32" "--nx 16384 --ny 2048 --band-rows 64" "--nx 16384 --ny 2048 --band-rows 128" \
  "--nx 4096 --ny 4096 --fused-steps 3" "--nx 4096 --ny 4096 --fused-steps 4"
