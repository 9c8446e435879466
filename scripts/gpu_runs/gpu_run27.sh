cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r27_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r27_pytest.log; tail -15 gpurun_out/r27_pytest.log | cut -c1-400
