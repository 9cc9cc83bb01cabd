CMD="python bench.py --config C3 --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
$CMD > gpurun_out/r02_c3_plain.json 2> gpurun_out/r02_c3_plain.err &&
APDE_PROFILE_PASS=5:2 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3_pass5.csv $CMD > gpurun_out/r02_c3_ncu.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_launches_c3_pass5.csv
