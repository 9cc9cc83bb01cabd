CMD="python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
APDE_PROFILE_PASS=9:2 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_sweep_classify" -c 1 -o gpurun_out/r02_prof_classify $CMD > gpurun_out/r02_ncu_classify.log 2>&1
tail -2 gpurun_out/r02_ncu_classify.log | cut -c1-200
