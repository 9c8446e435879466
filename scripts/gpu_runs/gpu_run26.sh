cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/density_sweep.py --out gpurun_out/r26_density_sweep.json > gpurun_out/r26_sweep.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum
for k in 16 32 64 128 256 512; do
  timeout 300 ncu --kernel-name regex:'normals_kernel|rsd_fast_kernel' --metrics $M --clock-control none --csv --log-file gpurun_out/r26_density_ncu_$k.csv python scripts/density_ncu.py $k > gpurun_out/r26_density_ncu_$k.log 2>&1
  tail -n 1 gpurun_out/r26_density_ncu_$k.log
done
# C4 traffic of the two pass kernels at the bench configuration (20 M points), one ncu pass
timeout 600 ncu --kernel-name regex:'normals_kernel|rsd_fast_kernel' --launch-skip 2 --launch-count 2 --metrics $M --clock-control none --csv --log-file gpurun_out/r26_c4_ncu.csv python scripts/density_ncu.py 0 20000000 > gpurun_out/r26_c4_ncu.log 2>&1
tail -n 2 gpurun_out/r26_c4_ncu.log
