cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_comm.py tests/test_gpu_parity.py tests/test_plugin_host.py -x -q -m gpu > gpurun_out/r20_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r20_pytest.log; tail -12 gpurun_out/r20_pytest.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r20_maxnn_launches.csv python scripts/maxnn_step_probe.py > gpurun_out/r20_ncu2.log 2>&1
tail -n 2 gpurun_out/r20_ncu2.log
