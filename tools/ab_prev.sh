#!/bin/bash
# maps of the current build vs a previous build of the library (apde_mvs_b200/_build/libapde_prev.so), then stage times of
# the headline and of a weak-texture-heavy workload
python tools/dump_maps.py /tmp/m_now.npz 2>&1 | tail -1
if [ -f apde_mvs_b200/_build/libapde_prev.so ]; then
  APDE_LIB=$PWD/apde_mvs_b200/_build/libapde_prev.so python tools/dump_maps.py /tmp/m_prev.npz 2>&1 | tail -1
  python tools/compare_maps.py /tmp/m_now.npz /tmp/m_prev.npz
fi
for extra in "" "--weak 0.4 --width 1600 --height 1200 --views-per-gpu 6 --src 5"; do
  python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1 $extra 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=j['roofline']['stage_ms']
print('[$extra] value %.4f ref-views/s launches %d' % (j['value'], j['gpu_launches']), {k: round(x, 1) for k, x in st.items()})"
done
