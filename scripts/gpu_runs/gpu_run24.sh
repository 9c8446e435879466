cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/maxnn_probe.py 20000000 2>&1 | tail -3
timeout 900 python -m pytest tests/test_comm.py::test_step_with_truncated_rsd tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
