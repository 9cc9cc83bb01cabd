# round-2 multi-GPU evidence on N GPUs of one box (N = the box size): job tests, weak-scaling bench line, strong scaling at 44 views
N=${1:-4}
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_multi_${N}gpu.log 2>&1; tail -3 gpurun_out/r02_multi_${N}gpu.log
python bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_weak_${N}gpu.json 2> gpurun_out/r02_bench_weak_${N}gpu.err
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    python bench.py --gpus $n --scaling strong --views-total 44 --steps 1 --warmup 1 --no-cpu-baseline --no-fusion --no-job-check > gpurun_out/r02_bench_strong44_${n}gpu.json 2> gpurun_out/r02_bench_strong44_${n}gpu.err
  fi
done
python - << 'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*gpu.json")):
    for ln in open(f):
        if ln.startswith("{"):
            j = json.loads(ln)
            print(f, "n=%d value %.3f e2e %.3f ms/step %.0f views %d" % (j["n_gpus"], j["value"], j["e2e"]["value"], j["ms_per_step"], j["config"]["views_total"]), j.get("exchange", {}))
PY
