#!/bin/bash
# usage: tools/ab_bench.sh "ENV=1 ..." [bench args]: one short bench line (stage times, texture fractions) with the given environment
envs="$1"; shift
env $envs python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fusion "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('[$envs] ms/step %.1f value %.4f | ' % (d['ms_per_step'], d['value']) + ' '.join('%s %.0f' % (k, v) for k, v in r['stage_ms'].items() if v > 20) + ' | tex ' + ' '.join('%s %.3f' % (k, v) for k, v in r['stage_tex_frac'].items()))"
