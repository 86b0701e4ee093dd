cd $GRAFT_REPO_ROOT
timeout 1500 python tools/gpu_stage.py > gpurun_out/r01f_stage_stdout.log 2>&1; echo "stage rc=$?"; tail -3 gpurun_out/r01f_stage_stdout.log
python bench.py > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err; echo "bench rc=$?"
python bench.py --lanes 1 --no-cpu-baseline > gpurun_out/r01f_bench_n1_lanes1.json 2> gpurun_out/r01f_bench_n1_lanes1.err
B="python bench.py --steps 2 --warmup 3 --pool 2048 --no-cpu-baseline --lanes 1"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"preprocess|flat|tc_conv|tc2_conv|avgpool" -s 57 -c 19 --csv --log-file gpurun_out/r01f_launches.csv $B > gpurun_out/r01f_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:preprocess_s2d -s 4 -c 1 -o gpurun_out/prof_pre_s2d_v2 -f $B > gpurun_out/r01f_ncu_pre.log 2>&1
python tools/sweep.py > gpurun_out/r01f_sweep.md 2> gpurun_out/r01f_sweep.err; echo "sweep rc=$?"
python tools/launch_table.py gpurun_out/r01f_launches.csv
python - <<'PY'
import json
for f in ['r01f_bench_n1','r01f_bench_n1_lanes1']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, round(d['value']), round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), 'trunk', d['roofline_trunk']['frac'], 'pre', d['roofline_preprocess']['frac'], d['roofline_preprocess']['avg_ms'], 'cpu', (d['cpu_baseline'] or {}).get('value'), d['clocks'])
PY
