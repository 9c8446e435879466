set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest.log
timeout 300 python scripts/shard_probe.py 1 > gpurun_out/r2_probe1_fast.txt 2>&1
CAB_RSD_LEGACY=1 timeout 300 python scripts/shard_probe.py 1 > gpurun_out/r2_probe1_legacy.txt 2>&1
timeout 300 python scripts/shard_probe.py 8 > gpurun_out/r2_probe8.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_shard_launches.csv python scripts/shard_build_probe.py 3 8 > gpurun_out/r2_ncu.log 2>&1
tail -5 gpurun_out/r2_pytest.log; cat gpurun_out/r2_probe1_fast.txt gpurun_out/r2_probe1_legacy.txt gpurun_out/r2_probe8.txt
