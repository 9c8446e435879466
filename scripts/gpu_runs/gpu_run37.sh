cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r37_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r37_pytest.log; tail -4 gpurun_out/r37_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r37_bench.json 2> gpurun_out/r37_bench.err; echo "bench rc $?"
tail -2 gpurun_out/r37_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r37_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','phases_ms','gpu_launches','exact_ms_per_step','grsd_clouds_per_s','max_nn_150'): print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
timeout 600 ncu --kernel-name regex:'rsd_fast_kernel|normals_kernel' --launch-skip 2 --launch-count 2 --clock-control none --metrics $M --csv --log-file gpurun_out/r37_plain_c4.csv python scripts/maxnn_step_probe.py 20000000 0 > gpurun_out/r37_ncu0.log 2>&1; echo "ncu plain rc $?"
timeout 600 ncu --kernel-name regex:'rsd_fast_kernel|normals_kernel' --launch-skip 2 --launch-count 2 --clock-control none --metrics $M --csv --log-file gpurun_out/r37_trunc_c4.csv python scripts/maxnn_step_probe.py 20000000 150 > gpurun_out/r37_ncu1.log 2>&1; echo "ncu trunc rc $?"
