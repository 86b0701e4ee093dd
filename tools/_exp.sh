cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r01f_bench_n2.json 2> gpurun_out/r01f_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r01f_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01f_bench_n2.json'))
print(round(d['value']), round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), d['config']['batch_per_gpu'], d['clocks'], d['e2e']['pass_seconds'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multigpu_dropin.py > gpurun_out/r01f_dropin_n2.log 2>&1; echo "dropin rc=$?"; tail -5 gpurun_out/r01f_dropin_n2.log
