#!/bin/bash
# N-GPU check: the 2-GPU NCCL tests and bench.py exactly as the driver launches it
n=${1:-2}; tag=${2:-dist}
mkdir -p gpurun_out
if [ "$n" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${tag}_pytest.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err; echo bench rc=$?
tail -c 600 gpurun_out/${tag}_n$n.err
python - "gpurun_out/${tag}_n$n.json" <<'PY'
import json, sys
try:
    line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
    d = json.loads(line)
    print("n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"] / 1e6, 2), "M/s", "eager", d.get("eager_ms_per_step"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("ERR", e)
PY
