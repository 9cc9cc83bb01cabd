# ncu launch list of ONE full-resolution geometric pass (pass 9, third occurrence) of the default bench, after a plain run of the same command
CMD="python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
$CMD > gpurun_out/r02_list_plain.json 2> gpurun_out/r02_list_plain.err &&
APDE_PROFILE_PASS=9:2 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_pass9_final.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_launches_pass9_final.csv
