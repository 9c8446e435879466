cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r46_bench_n8.json 2> gpurun_out/r46_bench_n8.err; echo "rc $?"
tail -2 gpurun_out/r46_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r46_bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['clocks'])
for r in d['per_rank_phase_ms']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
PY
