#!/bin/bash
# A/B of a kernel variant selected by an environment variable: identical maps over a whole schedule + stage times.
# usage: bash tools/ab_env.sh VAR [bench args]      (VAR=1 selects the alternative path)
set -e
VAR=$1; shift
python tools/dump_maps.py /tmp/m_a.npz 2>&1 | tail -1
env $VAR=1 python tools/dump_maps.py /tmp/m_b.npz 2>&1 | tail -1
python tools/compare_maps.py /tmp/m_a.npz /tmp/m_b.npz
for v in 0 1; do
  env $VAR=$v python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1 "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=j['roofline']['stage_ms']
print('$VAR=$v value %.4f ref-views/s ' % j['value'], {k: round(x, 1) for k, x in st.items()})"
done
