cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r19_shard_launches.csv python scripts/shard_build_probe.py 3 8 > gpurun_out/r19_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r19_maxnn_launches.csv python scripts/maxnn_step_probe.py > gpurun_out/r19_ncu2.log 2>&1
tail -2 gpurun_out/r19_ncu1.log gpurun_out/r19_ncu2.log
