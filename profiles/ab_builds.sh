#!/bin/bash
# A/B timing of two builds on the same box: gpu_ab.sh libA.so libB.so
for i in 1 2 3; do
  for L in "$@"; do
    VF_LIB_PATH=$PWD/vf-fem_b200/lib/$L python bench.py --skip-extras --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
  done
done
