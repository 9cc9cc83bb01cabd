# A/B of the strong propagation's candidate-cost memo on the headline workload (same box, back to back) + identity test
python -m pytest tests/test_gpu_edge.py -m gpu -q -k "memo or strong_propagation" > gpurun_out/r02_memo_test.log 2>&1; tail -3 gpurun_out/r02_memo_test.log
for m in 1 0 1 0; do
  APDE_MEMO=$m python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_memo_${m}.json 2> gpurun_out/r02_memo_${m}.err
  python - << PY
import json
for ln in open("gpurun_out/r02_memo_${m}.json"):
    if ln.startswith("{"):
        j = json.loads(ln)
        r = j["roofline"]
        print("memo=${m} value %.4f e2e %.4f evals/step %.4g memo/step %.4g prop_strong %.0f ms sweep %.0f ms" % (j["value"], j["e2e"]["value"], j["cost_evals_per_step"], j.get("memo_evals_per_step", 0), r["stage_ms"]["prop_strong"], r["stage_ms"]["depth_to_weak"]))
PY
done
