#!/bin/bash
# One GPU session: parity tests, bench lines per workload, ncu launch list of OUR kernels. Usage: gpu_round.sh TAG [workloads...]
TAG=${1:-x}; shift
WL=${@:-cfg3 cfg3b cfg2 cfg4 cfg5 passthrough}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for w in $WL; do
  timeout 300 python bench.py --workload $w > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$w.json").read().strip().splitlines()[-1])
    print("$w", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("$w FAILED", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu rc=$?"
