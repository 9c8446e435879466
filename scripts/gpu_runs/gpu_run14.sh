cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(echo "== 3 0 repl"; timeout 120 python scripts/comm_debug.py 3 0 0; echo "== 3 0 ranges"; timeout 120 python scripts/comm_debug.py 3 0 1; echo "== 2 60 repl nofb"; CAB_NO_FEEDBACK=1 timeout 120 python scripts/comm_debug.py 2 60 0) > gpurun_out/r14.log 2>&1
cat gpurun_out/r14.log
