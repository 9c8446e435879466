set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r1_gpus.txt
timeout 1200 python -m pytest tests/test_comm.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r1_pytest.log
timeout 300 python scripts/shard_probe.py 1 > gpurun_out/r1_probe1.txt 2>&1
timeout 300 python scripts/shard_probe.py 8 > gpurun_out/r1_probe8.txt 2>&1
timeout 300 python scripts/shard_probe.py 2 > gpurun_out/r1_probe2.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc $?" >> gpurun_out/r1_bench.err
tail -5 gpurun_out/r1_pytest.log; cat gpurun_out/r1_probe8.txt; tail -3 gpurun_out/r1_bench.err
