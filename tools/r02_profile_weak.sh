# ncu --set full of the weak-propagation kernels on the C3-shaped workload (one full-resolution APD pass), after a plain run of the same command
CMD="python bench.py --config C3 --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
$CMD > gpurun_out/r02_c3_plain.json 2> gpurun_out/r02_c3_plain.err &&
APDE_PROFILE_PASS=5:2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_weak_anchor1|k_weak_center1|k_weak_anchor3|k_weak_center3" -c 4 -o gpurun_out/r02_prof_weak $CMD > gpurun_out/r02_ncu_weak.log 2>&1
tail -3 gpurun_out/r02_ncu_weak.log; ls -la gpurun_out/*.ncu-rep
