set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r11_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err; echo "bench rc $?" >> gpurun_out/r11_bench.err
tail -8 gpurun_out/r11_pytest.log; tail -3 gpurun_out/r11_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r11_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','phases_ms','gpu_launches','exact_ms_per_step','grsd_clouds_per_s','max_nn_150'): print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
PY
