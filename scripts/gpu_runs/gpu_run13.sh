cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="tests/test_comm.py::test_local_group_every_rank_holds_all_results"
for v in NONE CAB_SERIAL_PHASES CAB_SERIAL_SORT CAB_NO_FEEDBACK; do
  echo "=== $v" >> gpurun_out/r13.log
  env $v=1 timeout 300 python -m pytest "$T" -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r13.log
done
cat gpurun_out/r13.log
