#!/bin/bash
# A/B of the two strong-propagation kernels: identical maps, stage time.  usage: bash tools/ab_strong.sh [bench args]
set -e
python tools/dump_maps.py /tmp/m_new.npz 2>&1 | tail -1
APDE_STRONG_V1=1 python tools/dump_maps.py /tmp/m_v1.npz 2>&1 | tail -1
python tools/compare_maps.py /tmp/m_new.npz /tmp/m_v1.npz
for v in 0 1; do
  APDE_STRONG_V1=$v python bench.py --no-cpu-baseline --no-fusion --steps 1 --warmup 1 "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=j['roofline']['stage_ms']
print('APDE_STRONG_V1=$v value %.4f ref-views/s  prop_strong %.1f ms  roofline frac %.4f' % (j['value'], st['prop_strong'], j['roofline']['frac']))"
done
