cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_grsd_cloud.py -x -q -m gpu > gpurun_out/r25_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r25_pytest.log; tail -25 gpurun_out/r25_pytest.log | cut -c1-500
