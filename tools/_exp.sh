cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_trunk.py -x -q -m gpu > gpurun_out/exp15_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/exp15_pytest.log
for ppb in 16 8; do
FX_DEBUG_S2D_PPB=$ppb ncu --metrics gpu__time_duration.sum --clock-control none -k regex:preprocess -c 6 --csv --log-file gpurun_out/exp15_pre$ppb.csv python bench.py --steps 3 --warmup 3 --pool 2048 --no-cpu-baseline > /dev/null 2>&1
echo "ppb $ppb:"; grep -o '"[0-9]*"$' gpurun_out/exp15_pre$ppb.csv | tr '\n' ' '; echo
done
python bench.py --steps 60 --warmup 5 --pool 4096 --no-cpu-baseline > gpurun_out/exp15_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/exp15_bench.json'))
print(round(d['value']), 'pre', d['roofline_preprocess']['avg_ms'], d['roofline_preprocess']['frac'], d['clocks']['sm_mhz'])
PY
