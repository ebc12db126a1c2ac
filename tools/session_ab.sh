# A/B of two builds of the library on the same box: lib/liblbm_b200.so against lib/liblbm_b200_old.so
L=mpilattice-boltzmann_b200/lib
cp $L/liblbm_b200.so $L/liblbm_b200_new.so
tools/sweep_r02.sh gpurun_out/ab_new "--fused-steps 4 --fused-deep 0" "--fused-steps 2"
cp $L/liblbm_b200_old.so $L/liblbm_b200.so
tools/sweep_r02.sh gpurun_out/ab_old "--fused-steps 4 --fused-deep 0" "--fused-steps 2"
cp $L/liblbm_b200_new.so $L/liblbm_b200.so
tools/sweep_r02.sh gpurun_out/ab_new "--fused-steps 4 --fused-deep 0" "--fused-steps 4 --fused-deep 0 --band-rows 192"
