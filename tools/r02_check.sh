# full GPU suite + a short headline bench line
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; tail -5 gpurun_out/r02_gputests.log
python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_check_bench.json 2> gpurun_out/r02_check_bench.err
python - << 'PY'
import json
for ln in open("gpurun_out/r02_check_bench.json"):
    if ln.startswith("{"):
        j = json.loads(ln); r = j["roofline"]
        print("value %.4f e2e %.4f" % (j["value"], j["e2e"]["value"]), {k: round(v / j["steps"], 1) for k, v in r["stage_ms"].items()}, r["stage_tex_frac"])
PY
