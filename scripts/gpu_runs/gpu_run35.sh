cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r35_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r35_pytest.log; tail -8 gpurun_out/r35_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r35_bench.json 2> gpurun_out/r35_bench.err; echo "rc $?"
tail -3 gpurun_out/r35_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r35_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','phases_ms','results_concatenated'): print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
PY
