#!/bin/bash
# usage: scratch/gpu_check.sh <tag>   (runs on the GPU box through gpurun)
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('ms/step %.4f kernel %.4f frac %.4f e2e %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value']))
PY
