#!/bin/bash
for cfg in "444 0" "444 64" "148 0" "296 0" "666 0" "888 0"; do
  set -- $cfg
  VF_PF_DIST=$1 VF_DEBUG_SKIP=$2 python bench.py --skip-extras --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pf $1 skip $2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done
