A=()
for shape in "--nx 2048 --ny 2048" "--nx 4096 --ny 4096" "--nx 8192 --ny 8192" "--nx 16384 --ny 2048" "--nx 16384 --ny 4096"; do
  A+=("$shape --fused-steps 2" "$shape --fused-steps 2 --fused-k7 1 --fused-ctas 12" "$shape --fused-steps 2 --fused-k7 1 --fused-ctas 12 --fused-deep 1" "$shape --fused-steps 3 --fused-ctas 11" "$shape --fused-steps 4 --fused-ctas 8")
done
A+=("--fused-steps 2 --fused-k7 1 --fused-ctas 12" "--fused-steps 2 --fused-k7 1 --fused-ctas 12 --fused-deep 1")
tools/sweep_r02.sh gpurun_out/r02y "${A[@]}"
