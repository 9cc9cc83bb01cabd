# round-2 final evidence on one B200: GPU suite, both arms at the driver's settings, C2 / C3 lines, ncu launch list + full captures
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gputests.log 2>&1; tail -3 gpurun_out/r02_final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ours.json 2> gpurun_out/r02_final_bench_ours.err
timeout 600 python bench.py --config C2 --steps 2 --warmup 1 > gpurun_out/r02_final_bench_c2_ours.json 2> gpurun_out/r02_final_bench_c2_ours.err
timeout 900 python bench.py --config C2 --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_c2_reference.json 2> gpurun_out/r02_final_bench_c2_reference.err
timeout 600 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_final_bench_c3_ours.json 2> gpurun_out/r02_final_bench_c3_ours.err
CMD="python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/r02_final_prof_plain.json 2> gpurun_out/r02_final_prof_plain.err &&
APDE_PROFILE_PASS=9:2 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_pass9.csv $CMD > gpurun_out/r02_final_ncu_list.log 2>&1
APDE_PROFILE_PASS=9:2 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_sweep_columns" -c 1 -o gpurun_out/r02_final_prof_sweep $CMD > gpurun_out/r02_final_ncu_sweep.log 2>&1
APDE_PROFILE_PASS=9:2 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_prop_strong|k_sweep_classify" -c 3 -o gpurun_out/r02_final_prof_strong $CMD > gpurun_out/r02_final_ncu_strong.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_final_launches_pass9.csv | head -12
for f in gpurun_out/r02_final_bench_*.json; do echo $f; head -c 400 $f; echo; done
