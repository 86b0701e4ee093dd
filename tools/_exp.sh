cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dropin.py -q -m gpu -x > gpurun_out/exp9_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/exp9_pytest.log
timeout 600 python tools/real_files_bench.py 4096 > gpurun_out/r01f_real_files.md 2> gpurun_out/r01f_real_files.err; echo "rc=$?"; cat gpurun_out/r01f_real_files.md; tail -3 gpurun_out/r01f_real_files.err
