cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(for i in 1 2; do echo "== NEW fb repl"; CAB_COMM_DEBUG=1 timeout 60 python scripts/comm_debug.py 3 0 0; echo "== NEW fb ranges"; timeout 60 python scripts/comm_debug.py 4 0 1; done) > gpurun_out/r16.log 2>&1
grep "it 0\|seq 1 \|rror" gpurun_out/r16.log
timeout 900 python -m pytest tests/test_comm.py -x -q -m gpu > gpurun_out/r16_pytest.log 2>&1; tail -5 gpurun_out/r16_pytest.log
