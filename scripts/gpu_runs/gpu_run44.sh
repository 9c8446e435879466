set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r44_bench_n8.json 2> gpurun_out/r44_bench_n8.err; echo "rc $?" >> gpurun_out/r44_bench_n8.err
timeout 400 python bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/r44_bench_n4.json 2> gpurun_out/r44_bench_n4.err; echo "rc $?" >> gpurun_out/r44_bench_n4.err
timeout 400 python bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r44_bench_n2.json 2> gpurun_out/r44_bench_n2.err; echo "rc $?" >> gpurun_out/r44_bench_n2.err
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r44_bench_n1.json 2> gpurun_out/r44_bench_n1.err; echo "rc $?" >> gpurun_out/r44_bench_n1.err
timeout 300 python bench.py --gpus 8 --workload grsd --steps 5 --warmup 2 > gpurun_out/r44_grsd_n8.json 2> gpurun_out/r44_grsd_n8.err; echo "rc $?" >> gpurun_out/r44_grsd_n8.err
tail -3 gpurun_out/r44_bench_n8.err; tail -2 gpurun_out/r44_grsd_n8.err; for f in gpurun_out/r44_bench_n8.json gpurun_out/r44_bench_n4.json gpurun_out/r44_bench_n2.json gpurun_out/r44_bench_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', d['value'])
    for r in d['per_rank_phase_ms']: print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
    if d.get('e2e'): print('   e2e', d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'], d['e2e'].get('shared_host_array'))
    print('   concat', d.get('results_concatenated'))
    print('   grsd_one_cloud', d.get('grsd_one_cloud'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r44_grsd_n8.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','unit','ms_per_step','n_gpus') if k in d})
except Exception as e: print('grsd ERR', e)
PY
