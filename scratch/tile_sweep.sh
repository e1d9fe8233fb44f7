#!/bin/bash
for t in 64 32 16 8; do
  ICMSLAM_TILE=$t python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/tile_$t.json 2> gpurun_out/tile_$t.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/tile_$t.json'))
    print('tile $t: ms/step %.4f kernel %.4f frac %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
except Exception as e:
    print('tile $t failed', e)
PY
done
ICMSLAM_TILE=16 python -m pytest tests -m gpu -x -q -k "fused or segment" 2>&1 | tail -3
ICMSLAM_TILE=8 python -m pytest tests -m gpu -x -q -k "fused or segment" 2>&1 | tail -3
