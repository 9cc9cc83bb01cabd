# round-2 evidence run: reference arm, then the ncu launch list and one full capture of the two top kernels for ONE
# full-resolution geometric pass of the default bench (APDE_PROFILE_PASS brackets it with cudaProfilerStart/Stop)
set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
CMD="python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
$CMD > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err &&
APDE_PROFILE_PASS=9:2 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_pass9.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
APDE_PROFILE_PASS=9:2 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_sweep_columns|k_prop_strong" -c 3 -o gpurun_out/r02_prof $CMD > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out
