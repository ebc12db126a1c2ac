A=()
for shape in "--nx 16384 --ny 16384" "--nx 8192 --ny 8192" "--nx 4096 --ny 4096" "--nx 2048 --ny 2048" "--nx 16384 --ny 2048" "--nx 16384 --ny 4096"; do
  for k in 2 3 4; do A+=("$shape --fused-steps $k"); done
done
tools/sweep_r02.sh gpurun_out/r02r "${A[@]}"
