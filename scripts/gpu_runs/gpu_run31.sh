cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r31_bench.json 2> gpurun_out/r31_bench.err; echo "bench rc $?" >> gpurun_out/r31_bench.err
tail -4 gpurun_out/r31_bench.err; 
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r31_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','phases_ms','gpu_launches','exact_ms_per_step','grsd_clouds_per_s','max_nn_150','grsd_one_cloud'): print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
print(d['roofline']['traffic'], d['roofline']['frac'])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d['parity'].items()})
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
