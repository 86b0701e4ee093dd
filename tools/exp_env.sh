#!/bin/bash
# A/B of environment knobs on one box: GPU tests first, then one bench line per "VAR=val[,VAR=val]" argument
# at one and two lanes.   usage: tools/exp_env.sh <tag> [--notest] <envset> [<envset> ...]   ("-" = no knob)
tag=$1; shift
if [ "$1" = "--notest" ]; then shift; else
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
fi
i=0
for es in "$@"; do i=$((i+1)); for l in ${LANES:-1 2}; do
  envs=$(echo "$es" | tr ',' ' '); [ "$es" = "-" ] && envs=""
  f=gpurun_out/${tag}_e${i}_l${l}
  env $envs timeout 300 python bench.py --no-cpu-baseline --lanes $l > $f.json 2> $f.err
  python - <<P
import json
try:
    d=json.load(open("$f.json"))
    print("[$es] lanes $l:", round(d["value"]), "e2e", round(d["e2e"]["value"]), "trunk_ms", round(d["roofline_trunk"]["avg_ms"],4), [round(x*1000,1) for x in d["layer_ms"] if x], d["clocks"]["sm_mhz"])
except Exception as ex: print("[$es] lanes $l FAILED", ex)
P
done; done
