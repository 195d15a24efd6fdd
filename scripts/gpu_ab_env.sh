#!/bin/bash
# GPU tests + bench A/B over environment settings: ./scripts/gpu_ab_env.sh tag "NAME=VAL ..." "NAME=VAL ..." ...   ("-" = defaults)
tag=$1; shift
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${tag}_pytest.log
i=0
for envs in "$@"; do
  i=$((i+1))
  if [ "$envs" = "-" ]; then envs=""; fi
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err; echo "[$i] $envs rc=$?"
  python - "gpurun_out/${tag}_$i.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   ms/step", round(d["ms_per_step"], 3), "eager", round(d.get("eager_ms_per_step", 0), 3), "core_sdf ms", round(d.get("core_sdf", {}).get("ms", 0), 3), "core_full ms", round(d.get("core_full", {}).get("ms", 0), 3), "e2e ms", round(d.get("e2e", {}).get("ms_per_step", 0), 3))
except Exception as e:
    print("   ERR", e)
PY
done
