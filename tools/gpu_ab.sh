#!/bin/bash
# A/B of the direct-tile path: GPU tests, then device-timed bench lines with SMR_DIRECT_K11=1 / 0.  Usage: gpu_ab.sh TAG [workloads]
TAG=${1:-ab}; shift
WL=${@:-cfg3 cfg3b cfg4}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log; fi
for w in $WL; do
  for d in 1 0; do
    SMR_DIRECT_K11=$d timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 300 --warmup 20 > gpurun_out/${TAG}_${w}_d$d.json 2> gpurun_out/${TAG}_${w}_d$d.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${w}_d$d.json").read().strip().splitlines()[-1])
    print("$w direct=$d", round(d["value"],1), round(d["ms_per_step"],4), {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$w direct=$d FAILED", e); print(open("gpurun_out/${TAG}_${w}_d$d.err").read()[-800:])
PY
  done
done
