cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_comm.py tests/test_grsd_cloud.py -x -q -m gpu > gpurun_out/r45_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r45_pytest.log; tail -3 gpurun_out/r45_pytest.log | cut -c1-300
timeout 400 python bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r45_bench_n2.json 2> gpurun_out/r45_bench_n2.err; echo "rc $?"
tail -2 gpurun_out/r45_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r45_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['clocks'])
for r in d['per_rank_phase_ms']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
PY
