cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_trunk.py -q -m gpu -x > gpurun_out/exp12_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/exp12_pytest.log
B="python bench.py --steps 60 --warmup 5 --pool 4096 --no-cpu-baseline"
$B --lanes 1 > gpurun_out/exp12_l1.json 2>/dev/null
$B > gpurun_out/exp12_l2.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp12_*.json')):
    d=json.load(open(f))
    print(f, round(d['value']), 'trunk_ms', round(d['roofline_trunk']['avg_ms'],4), 'sum', round(sum(d['layer_ms']),4), 'layers', [round(x,4) for x in d['layer_ms'][:10]], d['clocks']['sm_mhz'])
PY
