#!/usr/bin/env bash
# back-to-back PF launches after one L2 flush: the first pays the cold start, the rest run warm
export FLEXGPU_NO_PAIR=1
for n in 4736 37888; do
  for r in 1 2 3 5; do
    REPEAT=$r PF_MAX_ITER=8 python tools/bench_pf.py $n 20 > /tmp/o.json
    python -c "import json; d=json.load(open('/tmp/o.json')); print('n', $n, 'repeat', $r, round(d['ms']*1e3,1), 'us total')"
  done
done
