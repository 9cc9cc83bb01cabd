#!/bin/bash
# round-end style check on a GPU box: smoke, GPU tests, default bench (both arms); results under gpurun_out/
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print("ours: value %.4f  ms/step %.1f  e2e %.4f  frac %.4f  launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], j["gpu_launches"]))
print({k: round(v / j["steps"], 1) for k, v in j["roofline"]["stage_ms"].items()})
print("fusion:", j.get("fusion"))
PY
if [ "$1" == "ref" ]; then
  python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  python -c "
import json
j = json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1])
print('reference: value %.4f  e2e %.4f' % (j['value'], j['e2e']['value']))"
fi
