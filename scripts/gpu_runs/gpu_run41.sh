cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r41_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r41_pytest.log; tail -6 gpurun_out/r41_pytest.log | cut -c1-400
timeout 600 python scripts/density_sweep.py --out gpurun_out/r41_density_sweep.json > /dev/null 2> gpurun_out/r41_sweep.log; echo "sweep rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r41_density_sweep.json'))
for r in d['rows']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('k_target','mean_neighbours','candidates_per_query','build_ms','normals_ms','rsd_ms','step_ms','step_frac_of_hbm_peak')})
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phases_ms'])"
