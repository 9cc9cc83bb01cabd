timeout 240 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_quick_c3.json 2> gpurun_out/r02_quick_c3.err
timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fusion > gpurun_out/r02_quick_bench.json 2> gpurun_out/r02_quick_bench.err
python - << 'PY'
import json
for f in ("gpurun_out/r02_quick_c3.json", "gpurun_out/r02_quick_bench.json"):
    for ln in open(f):
        if ln.startswith("{"):
            j = json.loads(ln); r = j["roofline"]
            print(f, "value %.4f" % j["value"], {k: r.get(k) for k in ("prop_weak_tex_frac_counted_anchors", "anchor_patches_per_deformable_eval", "prop_weak_tex_frac_with_8_anchors", "traffic")}, r["stage_tex_frac"])
PY
