cd $GRAFT_REPO_ROOT
FX_FLAT2=1 timeout 600 python -m pytest tests/test_gpu_trunk.py -q -m gpu -x > gpurun_out/exp16_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/exp16_pytest.log
B="python bench.py --steps 60 --warmup 5 --pool 4096 --no-cpu-baseline"
timeout 300 $B --lanes 1 > gpurun_out/exp16_base_l1.json 2>/dev/null
FX_FLAT2=1 timeout 300 $B --lanes 1 > gpurun_out/exp16_flat2_l1.json 2>gpurun_out/exp16_flat2_l1.err
timeout 300 $B > gpurun_out/exp16_base_l2.json 2>/dev/null
FX_FLAT2=1 timeout 300 $B > gpurun_out/exp16_flat2_l2.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp16_*.json')):
    try:
        d=json.load(open(f))
        print(f, round(d['value']), 'trunk_ms', round(d['roofline_trunk']['avg_ms'],4), 'layers', [round(x,4) for x in d['layer_ms'][:6]], d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/exp16_flat2_l1.err
