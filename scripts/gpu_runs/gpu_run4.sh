set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/shard_probe.py 1 > gpurun_out/r4_probe1.txt 2>&1
timeout 900 python -m pytest tests/test_comm.py tests/test_gpu_parity.py tests/test_plugin_host.py -x -q -m gpu > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r4_pytest.log
timeout 300 python scripts/shard_probe.py 8 > gpurun_out/r4_probe8.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_shard_launches.csv python scripts/shard_build_probe.py 3 8 > gpurun_out/r4_ncu.log 2>&1
tail -5 gpurun_out/r4_pytest.log; cat gpurun_out/r4_probe1.txt gpurun_out/r4_probe8.txt
