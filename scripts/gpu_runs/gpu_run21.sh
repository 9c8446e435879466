cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "X=1" "CAB_TRUNC_NOSHRINK=1" "CAB_FAST_PER_SM=3" "CAB_FAST_PER_SM=2"; do echo "== $v"; env $v timeout 300 python scripts/maxnn_probe.py 20000000 2>&1 | tail -3; done > gpurun_out/r21.log 2>&1
cat gpurun_out/r21.log
