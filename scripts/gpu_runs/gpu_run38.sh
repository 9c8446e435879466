cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for U in 8 4 2; do for MN in 0 150; do
  echo "unroll $U max_nn $MN: $(CAB_FAST_UNROLL=$U timeout 300 python scripts/maxnn_step_probe.py 20000000 $MN 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read().strip()); print({k:round(d[k],3) for k in ('normals_ms','rsd_ms','step_ms') if k in d})")"
done; done 2>&1 | tee gpurun_out/r38_unroll.txt
