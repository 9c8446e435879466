set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r12_pytest.log
timeout 300 python scripts/shard_probe.py 8 > gpurun_out/r12_probe8.txt 2>&1
tail -30 gpurun_out/r12_pytest.log | cut -c1-600; cat gpurun_out/r12_probe8.txt
