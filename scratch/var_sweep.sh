#!/bin/bash
# usage: var_sweep.sh "half,tpp,occ half,tpp,occ ..."
for v in $1; do
  IFS=, read t p o <<< "$v"
  ICMSLAM_TILE=$t ICMSLAM_TPP=$p ICMSLAM_OCC=$o timeout 150 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/var_$t-$p-$o.json 2> gpurun_out/var_$t-$p-$o.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/var_$t-$p-$o.json'))
    print('variant $v: ms/step %.4f kernel %.4f frac %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
except Exception as e:
    print('variant $v failed', e)
PY
done
