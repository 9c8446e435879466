cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r47_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r47_pytest.log; tail -4 gpurun_out/r47_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r47_bench.json 2> gpurun_out/r47_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r47_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','steps','warmup','phases_ms','gpu_launches','clocks'): print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])
print(d['per_rank_phase_ms'])
PY
