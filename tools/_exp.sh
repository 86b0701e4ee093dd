cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_trunk.py -q -m gpu -x > gpurun_out/exp17_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/exp17_pytest.log
B="python bench.py --steps 60 --warmup 5 --pool 4096 --no-cpu-baseline"
FX_FLAT128X2=0 timeout 300 $B --lanes 1 > gpurun_out/exp17_base_l1.json 2>/dev/null
timeout 300 $B --lanes 1 > gpurun_out/exp17_x2_l1.json 2>gpurun_out/exp17_x2_l1.err
FX_FLAT128X2=0 timeout 300 $B > gpurun_out/exp17_base_l2.json 2>/dev/null
timeout 300 $B > gpurun_out/exp17_x2_l2.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp17_*.json')):
    try:
        d=json.load(open(f))
        print(f, round(d['value']), 'trunk_ms', round(d['roofline_trunk']['avg_ms'],4), 'layers', [round(x,4) for x in d['layer_ms'][5:10]], d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/exp17_x2_l1.err
