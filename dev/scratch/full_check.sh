#!/bin/bash
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('ms/step %.4f value %.1f kernel %.4f frac %.4f e2e %.1f launches/step %s cpu %s'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d.get('gpu_launches_per_step'),d['cpu_baseline']['value']))
PY
