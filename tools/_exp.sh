cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_preprocess.py -q -m gpu > gpurun_out/exp6_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/exp6_pytest.log
python bench.py --steps 100 --warmup 5 --pool 8192 --no-cpu-baseline > gpurun_out/exp6_bench.json 2> gpurun_out/exp6_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/exp6_bench.json'))
print(round(d['value']), round(d['e2e']['value']), d['e2e']['pass_seconds'], 'ms/step', round(d['ms_per_step'],4), 'trunk_ms', round(d['roofline_trunk']['avg_ms'],4), 'pre_ms', round(d['roofline_preprocess']['avg_ms'],4), 'host_enq', round(d['host_enqueue_ms_per_step'],3), d['clocks'])
PY
