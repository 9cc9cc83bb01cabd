# A/B: k_sweep_columns resident blocks per SM (register caps) and one vs two instances of the evaluation code, same box
for v in sw_u4 sw_m4 sw_u5 sw_m5 sw_m6; do
  export APDE_LIB=$PWD/ab/$v/libapde.so
  timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_ab_$v.json 2> gpurun_out/r02_ab_$v.err
  python - << PY
import json
for ln in open("gpurun_out/r02_ab_$v.json"):
    if ln.startswith("{"):
        j = json.loads(ln); r = j["roofline"]
        print("$v value %.4f sweep %.1f ms/step strong %.1f" % (j["value"], r["stage_ms"]["depth_to_weak"] / j["steps"], r["stage_ms"]["prop_strong"] / j["steps"]))
PY
done
