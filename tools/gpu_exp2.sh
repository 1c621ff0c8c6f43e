#!/bin/bash
# every what-if library x SMR_TILE_SORT on/off on the named workloads, device-timed.  Usage: gpu_exp2.sh TAG [workloads]
TAG=${1:-e2}; shift
WL=${@:-cfg3 cfg3b cfg4}
mkdir -p gpurun_out
for lib in tools/exp_libs/libsmr_*.so; do
  n=$(basename $lib .so); n=${n#libsmr_}
  for w in $WL; do
    for ts in 1 0; do
      SMR_TILE_SORT=$ts SMR_LIB_PATH=$PWD/$lib timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 300 --warmup 20 > gpurun_out/${TAG}_${n}_${w}_$ts.json 2> gpurun_out/${TAG}_${n}_${w}_$ts.err
      python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${n}_${w}_$ts.json").read().strip().splitlines()[-1])
    print("$n $w sort=$ts", round(d["value"],1), round(d["ms_per_step"],4), {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$n $w sort=$ts FAILED", e); print(open("gpurun_out/${TAG}_${n}_${w}_$ts.err").read()[-600:])
PY
    done
  done
done
