set -x
timeout 900 python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_eager.log 2>&1 && ncu --set full --clock-control none -k regex:"conv_halo_tc_kernel|conv_wgrad_halo|conv_wgrad_reduce|gn_pool_kernel|gn_bwd_reduce_kernel|gn_scale_kernel|gn_bwd_apply|gn_se_cluster|finalconv_gs|finalconv_bwd_tc|upsample2x|conv_tap_gemm|conv_s2_dgrad|adam_kernel|ffl_rows|ffl_cols" --launch-skip 150 --launch-count 64 -o /tmp/prof_r2_step python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_eager.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/prof_r2_step.ncu-rep > gpurun_out/r2_ncu_step_kernels.txt 2>&1
ls -la /tmp/prof_r2_step.ncu-rep
ncu -i /tmp/prof_r2_step.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/raw.csv')))
H=rows[0]
keep=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__cycles_active.avg','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
idx=[H.index(k) for k in keep if k in H]
with open('gpurun_out/r2_ncu_step_raw.csv','w') as f:
    w=csv.writer(f)
    for r in rows:
        w.writerow([ (r[i][:60] if i==H.index('Kernel Name') else r[i]) for i in idx])
PY
head -5 gpurun_out/r2_ncu_step_kernels.txt | cut -c1-200
