#!/bin/bash
# A/B of the stacked-operand sampler MLP kernel (default) against the tc4 kernel (NICER_TC_SAMPLER=4)
tag=${1:-samp}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_step.py tests/test_gpu_shipped_shapes.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_stack.json 2> gpurun_out/${tag}_stack.err; echo stack rc=$?
NICER_TC_SAMPLER=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_tc4.json 2> gpurun_out/${tag}_tc4.err; echo tc4 rc=$?
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for m in ("stack", "tc4"):
    try:
        d = json.load(open("gpurun_out/%s_%s.json" % (tag, m)))
        rt = d.get("roofline_tensor", {})
        print(m, "ms/step", round(d["ms_per_step"], 3), "eager", round(d.get("eager_ms_per_step", 0), 3), "core_sdf ms", round(d.get("core_sdf", {}).get("ms", 0), 3), "sampler kernel ms", rt.get("ms_per_launch"), rt.get("kernel", "")[:40])
    except Exception as e:
        print(m, "ERR", e)
PY
