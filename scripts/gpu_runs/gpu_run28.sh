cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plugin_host.py -x -q -m gpu > gpurun_out/r28_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r28_pytest.log; tail -12 gpurun_out/r28_pytest.log | cut -c1-400
