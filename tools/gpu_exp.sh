#!/bin/bash
# Times every what-if library of tools/exp_libs/ on the named workloads (default cfg3).  Usage: gpu_exp.sh TAG [workloads]
TAG=${1:-exp}; shift
WL=${@:-cfg3}
mkdir -p gpurun_out
for lib in tools/exp_libs/libsmr_*.so; do
  n=$(basename $lib .so); n=${n#libsmr_}
  for w in $WL; do
    SMR_LIB_PATH=$PWD/$lib timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 300 --warmup 20 > gpurun_out/${TAG}_${n}_$w.json 2> gpurun_out/${TAG}_${n}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${n}_$w.json").read().strip().splitlines()[-1])
    print("$n $w", round(d["value"],1), round(d["ms_per_step"],4), {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$n $w FAILED", e); print(open("gpurun_out/${TAG}_${n}_$w.err").read()[-800:])
PY
  done
done
