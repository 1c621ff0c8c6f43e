#!/bin/bash
# Round-end evidence run on 1 GPU: tests, every workload's bench line, launch list, one full ncu capture per hot kernel.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for w in cfg3 cfg3b cfg2 cfg4 cfg5 grid25 passthrough; do
  timeout 300 python bench.py --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],4), {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()}, "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3))
except Exception as e: print("$w FAILED", e)
PY
done
# A/B switches of the tile plan on the headline workload (device-timed only)
SMR_DIRECT_K11=0 timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_cfg3_nodirect.json 2>/dev/null
SMR_DIRECT_K11=0 SMR_TILE_SORT=0 timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_cfg3_nodirect_nosort.json 2>/dev/null
python - <<PY
import json
for n in ("nodirect", "nodirect_nosort"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_cfg3_%s.json" % n).read().strip().splitlines()[-1])
        print("cfg3", n, round(d["value"],1), {k: round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print("cfg3", n, "FAILED", e)
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
tail -c 700 gpurun_out/${TAG}_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 24 -c 40 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_resample_tma -s 12 -c 1 -o gpurun_out/${TAG}_ncu_fused_cfg3 -f python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_composite -s 12 -c 1 -o gpurun_out/${TAG}_ncu_comp_cfg3 -f python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_comp.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_resample_tma0 -s 6 -c 1 -o gpurun_out/${TAG}_ncu_anyratio_cfg5 -f python bench.py --workload cfg5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_anyratio.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integer_ratio or tiles_02 or nv12_input or full_range or random_noise or transition_fractional or (box_reduced and 256)" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integer_ratio or (transition_fractional and 0.5) or (box_reduced and 384 and nv12)" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck.log
# the reports are tens of MB each and gpurun_out/ travels back only below 64 MiB: turn them into text here, keep the text
O=gpurun_out/profiles bash tools/collect_profiles.sh $TAG > gpurun_out/${TAG}_collect.log 2>&1
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
echo done
