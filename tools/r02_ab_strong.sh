# A/B: k_prop_strong with 3 vs 4 resident blocks per SM (packed reference patch in shared memory makes 4 fit), same box, alternating
timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q -k "strong_propagation or half_sweep or strong" > gpurun_out/r02_ab_strong_tests.log 2>&1; tail -1 gpurun_out/r02_ab_strong_tests.log
for v in base st4 base st4; do
  if [ $v = st4 ]; then export APDE_LIB=$PWD/ab/st4/libapde.so; else unset APDE_LIB; fi
  timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_ab_strong_$v.json 2> gpurun_out/r02_ab_strong_$v.err
  python - << PY
import json
for ln in open("gpurun_out/r02_ab_strong_$v.json"):
    if ln.startswith("{"):
        j = json.loads(ln); r = j["roofline"]
        print("$v value %.4f sweep %.1f ms/step strong %.1f" % (j["value"], r["stage_ms"]["depth_to_weak"] / j["steps"], r["stage_ms"]["prop_strong"] / j["steps"]))
PY
done
unset APDE_LIB
APDE_LIB=$PWD/ab/st4/libapde.so timeout 200 python -m pytest tests/test_gpu_edge.py -m gpu -x -q -k "strong_propagation" > gpurun_out/r02_ab_strong_tests4.log 2>&1; tail -1 gpurun_out/r02_ab_strong_tests4.log
