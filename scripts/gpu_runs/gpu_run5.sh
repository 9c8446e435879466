set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/san_probe.py 40000 > gpurun_out/r5_plain.log 2>&1
timeout 900 python -m pytest tests/test_comm.py tests/test_gpu_parity.py tests/test_plugin_host.py -x -q -m gpu > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r5_pytest.log
timeout 300 python scripts/shard_probe.py 1 > gpurun_out/r5_probe1.txt 2>&1
timeout 300 python scripts/shard_probe.py 8 > gpurun_out/r5_probe8.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rsd_fast -c 1 -o gpurun_out/r5_rsd_fast python scripts/shard_build_probe.py 0 1 > gpurun_out/r5_ncu.log 2>&1
cat gpurun_out/r5_plain.log; tail -4 gpurun_out/r5_pytest.log; cat gpurun_out/r5_probe1.txt gpurun_out/r5_probe8.txt
