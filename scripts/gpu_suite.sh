#!/bin/bash
# One GPU-box pass: the GPU test-suite, the default bench line, the other BASELINE configs.  Outputs under gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; echo n1 rc=$?; tail -3 gpurun_out/${tag}_n1.err
timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err; echo c3 rc=$?; tail -2 gpurun_out/${tag}_c3.err
timeout 300 python bench.py --config C2-tracking --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_trk.json 2> gpurun_out/${tag}_trk.err; echo trk rc=$?; tail -2 gpurun_out/${tag}_trk.err
python - <<PY
import json
for f in ("${tag}_n1","${tag}_c3","${tag}_trk"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["ms_per_step"], d.get("eager_ms_per_step"), d["value"], d["e2e"]["value"], d.get("e2e_frame_upload",{}).get("value"), d.get("core_sdf",{}).get("ray_samples_per_s"), d.get("roofline",{}).get("frac"))
        for k in ("tracking_loop","render_image","sdf_grid_256","cpu_baseline"):
            if k in d: print("   ",k,{a:b for a,b in d[k].items() if a not in ("note","sample")})
    except Exception as e: print(f, "ERR", e)
PY
