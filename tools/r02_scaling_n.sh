# one strong-scaling line (fixed 88-view scene) and, with a second argument "weak", the weak-scaling line at N GPUs
N=${1:-8}
if [ "$2" = "weak" ]; then
  timeout 600 python bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_weak_${N}gpu.json 2> gpurun_out/r02_bench_weak_${N}gpu.err
fi
timeout 900 python bench.py --gpus $N --scaling strong --views-total 88 --steps 1 --warmup 1 --no-cpu-baseline --no-fusion --no-job-check > gpurun_out/r02_bench_strong88_${N}gpu.json 2> gpurun_out/r02_bench_strong88_${N}gpu.err
python - << 'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*88_*gpu.json") + glob.glob("gpurun_out/r02_bench_weak_*gpu.json")):
    for ln in open(f):
        if ln.startswith("{"):
            j = json.loads(ln)
            print(f, "n=%d value %.3f e2e %.3f ms/step %.0f views %d" % (j["n_gpus"], j["value"], j["e2e"]["value"], j["ms_per_step"], j["config"]["views_total"]), j.get("exchange", {}))
PY
