#!/bin/bash
# A/B of the two-threads-per-point SDF kernels (NICER_TC_SPLIT=1, default) against the one-thread-per-point ones (=0)
tag=${1:-split}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_step.py tests/test_gpu_shipped_shapes.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_on.json 2> gpurun_out/${tag}_on.err; echo on rc=$?; tail -2 gpurun_out/${tag}_on.err
NICER_TC_SPLIT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_off.json 2> gpurun_out/${tag}_off.err; echo off rc=$?; tail -2 gpurun_out/${tag}_off.err
python - <<PY
import json
for f in ("${tag}_on","${tag}_off"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["ms_per_step"], d.get("eager_ms_per_step"), d["value"], d.get("core_sdf",{}).get("ms"), d.get("core_full",{}).get("ms"))
    except Exception as e: print(f, "ERR", e)
PY
