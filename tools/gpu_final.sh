#!/bin/bash
# Round-end evidence run on 1 GPU: tests, every workload's bench line, launch list, one full ncu capture per hot kernel.
TAG=${1:-final}
mkdir -p gpurun_out
bash tools/gpu_round.sh $TAG cfg3 cfg3b cfg2 cfg4 cfg5 passthrough
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_resample_fused -s 3 -c 1 -o gpurun_out/${TAG}_ncu_fused_cfg3 -f python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e > gpurun_out/${TAG}_ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_composite -s 3 -c 1 -o gpurun_out/${TAG}_ncu_comp_cfg3 -f python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e > gpurun_out/${TAG}_ncu_comp.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
tail -c 600 gpurun_out/${TAG}_reference_arm.json
echo done
