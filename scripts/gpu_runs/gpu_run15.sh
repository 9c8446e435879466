cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(cd _old; echo "== OLD 3 0 repl"; timeout 120 python scripts/comm_debug.py 3 0 0; echo "== OLD 3 0 repl again"; timeout 120 python scripts/comm_debug.py 3 0 0 ) > gpurun_out/r15.log 2>&1
(echo "== NEW all serial nofb"; CAB_NO_FEEDBACK=1 CAB_SERIAL_SORT=1 CAB_SERIAL_PHASES=1 timeout 120 python scripts/comm_debug.py 3 0 0) >> gpurun_out/r15.log 2>&1
cat gpurun_out/r15.log
