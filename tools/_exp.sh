cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_postprocess.py -q -m gpu > gpurun_out/exp8_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/exp8_pytest.log
