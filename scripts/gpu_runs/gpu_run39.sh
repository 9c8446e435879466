cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "$1 max_nn $2: $(env $1 timeout 300 python scripts/maxnn_step_probe.py 20000000 $2 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read().strip()); print({k:round(d[k],3) for k in ('normals_ms','rsd_ms','step_ms') if k in d})")"; }
( run CAB_X=0 0; run CAB_FAST_UNROLL=2 0; run CAB_NORMALS_SCALAR=1 0; run CAB_X=0 150; run CAB_FAST_UNROLL=2 150 ) 2>&1 | tee gpurun_out/r39_ab.txt
