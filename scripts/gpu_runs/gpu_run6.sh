set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r6_topo.txt 2>&1
timeout 600 python -m pytest tests/test_comm.py -x -q -m gpu > gpurun_out/r6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r6_pytest.log
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r6_bench_n2.json 2> gpurun_out/r6_bench_n2.err; echo "bench rc $?" >> gpurun_out/r6_bench_n2.err
timeout 300 python bench.py --gpus 2 --steps 5 --warmup 3 --workload grsd > gpurun_out/r6_grsd_n2.json 2> gpurun_out/r6_grsd_n2.err; echo "grsd rc $?" >> gpurun_out/r6_grsd_n2.err
tail -5 gpurun_out/r6_pytest.log; tail -5 gpurun_out/r6_bench_n2.err; cat gpurun_out/r6_bench_n2.json | cut -c1-1500; tail -3 gpurun_out/r6_grsd_n2.err
