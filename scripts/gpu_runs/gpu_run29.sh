cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_comm.py tests/test_grsd_cloud.py -x -q -m gpu > gpurun_out/r29_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r29_pytest.log; tail -6 gpurun_out/r29_pytest.log | cut -c1-300
