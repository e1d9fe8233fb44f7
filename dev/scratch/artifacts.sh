#!/bin/bash
# usage: artifacts.sh <tag>  -- plain bench line, ncu launch list, ncu --set full captures of the top kernels
tag=$1
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_fused|k_solve_colour" -c 3 -s 9 -o gpurun_out/${tag}_prof -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu2.log
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_n1.json'))
print('ms/step %.4f value %.1f kernel %.4f frac %.4f e2e %.1f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value']))
PY
