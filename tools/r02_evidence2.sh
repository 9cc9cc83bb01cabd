# round-2 closing evidence with the final build: GPU suite, smoke, headline at the driver's settings, C2, C3, launch list
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gputests.log 2>&1; tail -3 gpurun_out/r02_final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ours.json 2> gpurun_out/r02_final_bench_ours.err
timeout 600 python bench.py --config C2 --steps 2 --warmup 1 > gpurun_out/r02_final_bench_c2_ours.json 2> gpurun_out/r02_final_bench_c2_ours.err
timeout 600 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_final_bench_c3_ours.json 2> gpurun_out/r02_final_bench_c3_ours.err
CMD="python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/r02_final_prof_plain.json 2> gpurun_out/r02_final_prof_plain.err &&
APDE_PROFILE_PASS=9:2 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_pass9.csv $CMD > gpurun_out/r02_final_ncu_list.log 2>&1
APDE_PROFILE_PASS=9:2 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_sweep_classify" -c 1 -o gpurun_out/r02_final_prof_classify $CMD > gpurun_out/r02_final_ncu_classify.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_final_launches_pass9.csv > gpurun_out/r02_final_launch_summary.txt; head -8 gpurun_out/r02_final_launch_summary.txt
for f in gpurun_out/r02_final_bench_*ours.json; do echo $f; head -c 300 $f; echo; done
