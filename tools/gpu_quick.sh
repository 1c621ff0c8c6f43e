#!/bin/bash
# Quick GPU session: parity tests, then bench lines of the named workloads. Usage: gpu_quick.sh TAG [workloads...]
TAG=${1:-q}; shift
WL=${@:-cfg3 cfg3b}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log; fi
for w in $WL; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$w.json").read().strip().splitlines()[-1])
    print("$w", round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],4), {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$w FAILED", e); print(open("gpurun_out/${TAG}_$w.err").read()[-1500:])
PY
done
