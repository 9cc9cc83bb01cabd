# quick A/B: strong-propagation identity tests + a short headline bench line
timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_reference_pins.py -m gpu -x -q -k "strong_propagation or pins or reference or banded" > gpurun_out/r02_quick_tests.log 2>&1; tail -2 gpurun_out/r02_quick_tests.log
timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_quick_bench.json 2> gpurun_out/r02_quick_bench.err
python - << 'PY'
import json
for ln in open("gpurun_out/r02_quick_bench.json"):
    if ln.startswith("{"):
        j = json.loads(ln); r = j["roofline"]
        print("value %.4f e2e %.4f" % (j["value"], j["e2e"]["value"]), {k: round(v / j["steps"], 1) for k, v in r["stage_ms"].items()}, r["stage_tex_frac"])
PY
