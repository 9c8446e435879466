cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r43_bench.json 2> gpurun_out/r43_bench.err ) 2>&1 | grep real; echo "bench rc $?"
tail -2 gpurun_out/r43_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r43_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','phases_ms','gpu_launches','exact_ms_per_step','grsd_clouds_per_s','max_nn_150'): print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
print(d['roofline']['frac'], d['roofline']['traffic'])
print(d['parity'].get('fast-fp32'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r43_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r43_ncu_launches.log 2>&1; echo "ncu launches rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rsd_fast|normals_kernel' -s 2 -c 2 -o gpurun_out/r43_passes python scripts/maxnn_step_probe.py 20000000 0 > gpurun_out/r43_ncu_full.log 2>&1; echo "ncu full rc $?"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum
for k in 16 32 64 128 256 512; do
  timeout 300 ncu --kernel-name regex:'normals_kernel|rsd_fast_kernel' --metrics $M --clock-control none --csv --log-file gpurun_out/r43_density_ncu_$k.csv python scripts/density_ncu.py $k > gpurun_out/r43_density_ncu_$k.log 2>&1
  tail -n 1 gpurun_out/r43_density_ncu_$k.log
done
ls -la gpurun_out/r43_* | head -20
