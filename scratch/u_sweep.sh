#!/bin/bash
for u in 4 5 6 8; do
  cp icm_slam_b200/lib/libicmslam_u$u.so icm_slam_b200/lib/libicmslam.so
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('U=$u: ms/step %.4f kernel %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms']))"
done
cp icm_slam_b200/lib/libicmslam_u4.so icm_slam_b200/lib/libicmslam.so
