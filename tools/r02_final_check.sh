# last check of the round: full GPU suite, smoke, one headline line
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_last_gputests.log 2>&1; tail -2 gpurun_out/r02_last_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_last_smoke.log 2>&1; tail -1 gpurun_out/r02_last_smoke.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err
python - << 'PY'
import json
for ln in open("gpurun_out/r02_last_bench.json"):
    if ln.startswith("{"):
        j = json.loads(ln); r = j["roofline"]
        print("value %.4f e2e %.4f" % (j["value"], j["e2e"]["value"]), r["kernel"], "frac %.3f" % r["frac"], {k: round(v / j["steps"], 1) for k, v in r["stage_ms"].items()}, r["stage_tex_frac"])
PY
