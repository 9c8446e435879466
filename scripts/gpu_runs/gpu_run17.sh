cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r17_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r17_pytest.log; tail -8 gpurun_out/r17_pytest.log
