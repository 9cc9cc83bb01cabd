timeout 200 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_last_bench_c2_ours.json 2> gpurun_out/r02_last_bench_c2_ours.err
timeout 100 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_last_bench_c3_ours.json 2> gpurun_out/r02_last_bench_c3_ours.err
python - << 'PY'
import json
for f in ("c2", "c3"):
    for ln in open("gpurun_out/r02_last_bench_%s_ours.json" % f):
        if ln.startswith("{"):
            j = json.loads(ln); r = j["roofline"]
            print(f, "value %.4f e2e %.4f" % (j["value"], j["e2e"]["value"]), r["kernel"], "frac %.3f" % r["frac"], r["stage_tex_frac"], r.get("prop_weak_tex_frac_counted_anchors"))
PY
