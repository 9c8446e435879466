cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r50_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r50_pytest.log; tail -4 gpurun_out/r50_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r50_bench.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r50_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])
PY
