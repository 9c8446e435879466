cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_comm.py tests/test_gpu_parity.py tests/test_plugin_host.py -x -q -m gpu > gpurun_out/r22_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r22_pytest.log; tail -12 gpurun_out/r22_pytest.log | cut -c1-400
timeout 300 python scripts/maxnn_probe.py 20000000 2>&1 | tail -3
python - <<'PY'
import sys, time, pathlib
sys.path.insert(0, '.')
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth
pts = synth.room(20_000_000)
ctx = cab.Context(0)
ctx.upload(pts)
for mn in (0, 150, 75):
    for _ in range(3):
        t0=time.perf_counter(); ctx.step_normals_rsd(0.02, 0.02, max_nn_rsd=mn); dt=time.perf_counter()-t0
    p=ctx.profile()
    print(f"step max_nn_rsd {mn}: wall {dt*1e3:.2f} ms build {p['build_ms']:.2f} normals {p['normals_ms']:.2f} rsd {p['rsd_ms']:.2f} step {p['step_ms']:.2f}")
PY
