#!/bin/bash
# One box, the whole evidence set of a code state:  tools/record_round.sh <tag>
#   GPU tests, bench (default = two lanes + CPU baseline), bench --lanes 1, the reference arm, the ncu launch list of one
#   step and an `ncu --set full` capture of the same step.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-rec}
o=gpurun_out/$tag
timeout 900 python -m pytest tests -m gpu -x -q > ${o}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 ${o}_pytest.log
timeout 600 python bench.py > ${o}_bench_n1.json 2> ${o}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --lanes 1 --no-cpu-baseline > ${o}_bench_n1_lanes1.json 2> ${o}_bench_n1_lanes1.err; echo "bench lanes1 rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > ${o}_bench_reference_arm.json 2> ${o}_bench_reference_arm.err; echo "reference arm rc=$?"
K='regex:preprocess|flat|stemw|tc_conv|tc2_conv|avgpool'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k "$K" -s 57 -c 19 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --pool 2048 --no-cpu-baseline --lanes 1 > ${o}_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 57 -c 19 -o ${o}_step \
  python bench.py --steps 2 --warmup 3 --pool 2048 --no-cpu-baseline --lanes 1 > ${o}_step.log 2>&1; echo "ncu full rc=$?"
python - <<P
import json
for f in ("bench_n1","bench_n1_lanes1","bench_reference_arm"):
    try:
        d=json.load(open("${o}_%s.json"%f)); print(f, round(d["value"],1), d.get("e2e",{}).get("value"), (d.get("roofline_trunk") or {}).get("avg_ms"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex: print(f,"FAILED",ex)
P
