#!/bin/bash
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.4f fused-kernel %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms']))"; }
run ICMSLAM_OCC=512
run ICMSLAM_OCC=768
run ICMSLAM_TILE=32 ICMSLAM_OCC=768
run ICMSLAM_TILE=32 ICMSLAM_OCC=640
