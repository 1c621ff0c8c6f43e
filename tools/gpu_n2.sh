#!/bin/bash
# 2-GPU session: the default bench line under torchrun (cfg3 + the cfg4 leg under every exchange mode).  Usage: gpu_n2.sh TAG [N]
TAG=${1:-n2}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-100} --warmup 5 --no-e2e > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}.json").read().strip().splitlines()[-1])
    print("cfg3 N=$N", round(d["value"],1), round(d["ms_per_step"],4))
    print(json.dumps(d["config"]["secondary"], indent=1))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/${TAG}.err").read()[-3000:])
PY
