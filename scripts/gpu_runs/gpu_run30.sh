cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pfh.py tests/test_plugin_host.py -x -q -m gpu > gpurun_out/r30_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r30_pytest.log; tail -15 gpurun_out/r30_pytest.log | cut -c1-400
