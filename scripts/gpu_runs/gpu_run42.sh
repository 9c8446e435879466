cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r42_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r42_pytest.log; tail -6 gpurun_out/r42_pytest.log | cut -c1-400
