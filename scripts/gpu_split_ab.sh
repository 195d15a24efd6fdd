#!/bin/bash
# A/B of the two-threads-per-point SDF kernels: NICER_TC_SPLIT bit mask (1 = A, 2 = B, 4 = T, 8 = R)
tag=${1:-split}
shift
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_step.py tests/test_gpu_shipped_shapes.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${tag}_pytest.log
for m in "$@"; do
  NICER_TC_SPLIT=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_m$m.json 2> gpurun_out/${tag}_m$m.err; echo mask $m rc=$?
done
python - "$tag" "$@" <<'PY'
import json, sys
tag = sys.argv[1]
for m in sys.argv[2:]:
    try:
        d = json.load(open("gpurun_out/%s_m%s.json" % (tag, m)))
        print("mask", m, "ms/step", round(d["ms_per_step"], 3), "eager", round(d.get("eager_ms_per_step", 0), 3), "core_sdf ms", round(d.get("core_sdf", {}).get("ms", 0), 3), "core_full ms", round(d.get("core_full", {}).get("ms", 0), 3))
    except Exception as e:
        print("mask", m, "ERR", e)
PY
