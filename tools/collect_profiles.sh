#!/bin/bash
# Turns the gpurun_out/ evidence of tools/gpu_final.sh TAG into tracked text artefacts under profiles/.
TAG=${1:-r02}
O=${O:-profiles}   # on the GPU box: O=gpurun_out/profiles (the .ncu-rep files are too big to travel back)
mkdir -p $O
for w in cfg3 cfg3b cfg2 cfg4 cfg5 grid25 passthrough; do [ -s gpurun_out/${TAG}_bench_$w.json ] && tail -1 gpurun_out/${TAG}_bench_$w.json > $O/${TAG}_bench_$w.json; done
for n in nodirect nodirect_nosort; do [ -s gpurun_out/${TAG}_bench_cfg3_$n.json ] && tail -1 gpurun_out/${TAG}_bench_cfg3_$n.json > $O/${TAG}_bench_cfg3_$n.json; done
[ -s gpurun_out/${TAG}_reference_arm.json ] && tail -1 gpurun_out/${TAG}_reference_arm.json > $O/${TAG}_bench_cfg3_reference_arm.json
cp gpurun_out/${TAG}_launches_cfg3.csv $O/${TAG}_launches_cfg3.csv 2>/dev/null
for k in fused comp; do
  rep=gpurun_out/${TAG}_ncu_${k}_cfg3.ncu-rep
  [ -s $rep ] || continue
  { echo "==== ncu --set full --clock-control none, one launch inside bench.py --workload cfg3 (raw page) ===="; ncu -i $rep --page raw 2>/dev/null; } > $O/${TAG}_ncu_raw_${k}_cfg3.txt
  ncu -i $rep --page details --csv 2>/dev/null > $O/${TAG}_ncu_details_${k}_cfg3.csv
done
python tools/ncu_lines.py gpurun_out/${TAG}_ncu_fused_cfg3.ncu-rep k_resample_tma 0.8 > $O/${TAG}_ncu_lines_fused_cfg3.txt 2>/dev/null
python tools/ncu_lines.py gpurun_out/${TAG}_ncu_comp_cfg3.ncu-rep k_composite 0.8 > $O/${TAG}_ncu_lines_comp_cfg3.txt 2>/dev/null
rep=gpurun_out/${TAG}_ncu_anyratio_cfg5.ncu-rep
if [ -s $rep ]; then
  { echo "==== ncu --set full --clock-control none, one launch of k_resample_tma0 inside bench.py --workload cfg5 (raw page) ===="; ncu -i $rep --page raw 2>/dev/null; } > $O/${TAG}_ncu_raw_anyratio_cfg5.txt
  python tools/ncu_lines.py $rep k_resample_tma0 0.8 > $O/${TAG}_ncu_lines_anyratio_cfg5.txt 2>/dev/null
fi
for t in memcheck racecheck; do [ -s gpurun_out/${TAG}_$t.log ] && grep -v "^$" gpurun_out/${TAG}_$t.log | tail -12 > $O/${TAG}_sanitizer_$t.log; done
tail -5 gpurun_out/${TAG}_pytest.log > $O/${TAG}_pytest_gpu_tail.log
python - <<PY
import csv, json, subprocess
tr = {}
for k, cls in (("fused", "resample_fused"), ("comp", "composite")):
    try:
        out = subprocess.run(["ncu", "-i", "gpurun_out/${TAG}_ncu_%s_cfg3.ncu-rep" % k, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, v = rows[0], rows[2]
        g = lambda name: float(v[h.index(name)])
        unit = {r: u for r, u in zip(rows[0], rows[1])}
        def bytes_of(name):
            x = g(name); u = unit[name]
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tr[cls] = int(bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"))
    except Exception as e:
        print("traffic", k, e)
d = json.load(open("profiles/traffic.json"))
d.setdefault("cfg3", {}).update(tr)
json.dump(d, open("$O/traffic.json", "w"), indent=1, sort_keys=True)
print("traffic", tr)
PY
ls -la $O | grep ${TAG} | awk '{print $5, $9}'
