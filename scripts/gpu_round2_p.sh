M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__inst_executed_op_branch.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,sm__inst_executed_pipe_xu.sum,smsp__thread_inst_executed_per_inst_executed.ratio"
for cfg in 3 12; do
  B200_EXACT_CFG=$cfg ncu --clock-control none --metrics $M -k regex:body_exact --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r02_exact_cfg$cfg.csv python scripts/exact_kernel_run.py > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_exact_cfg$cfg.csv')) if len(r)>10]
h=rows[0]; i_m=h.index('Metric Name'); i_v=h.index('Metric Value'); i_k=h.index('Kernel Name')
print('cfg $cfg', rows[1][i_k][:90])
for r in rows[1:]: print('   ', r[i_m], r[i_v])
PY
done
