#!/bin/bash
# bench at N GPUs on the box (run under gpurun --gpus N): dev/scale.sh N [extra bench args]
N=$1; shift
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
    for k in ["value","ms_per_step","e2e","gpu_launches_per_step","result_sha256","exchange","nccl_exchange_arm"]: print(k, d.get(k))
    print("kernel_ms", d["roofline"]["kernel_ms"])
except Exception as e:
    print("no line", e)
PY
grep -v "^frame" gpurun_out/bench_n$N.err | grep -i "error\|Traceback\|raise" | head -5
