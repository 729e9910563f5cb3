# Round-end measurement pass on one GPU: smoke, bench lines of every workload, reference arm, launch list, ncu capture.
set -x
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/gpu.txt
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 > $O/bench_default_n1.json 2> $O/bench_default_n1.err; echo "bench rc=$?"
for w in overfit_capacity overfit_capacity_wide128 inference; do timeout 600 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_${w}_n1.json 2> $O/bench_${w}_n1.err; echo "$w rc=$?"; done
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv --log-file $O/launches_step.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu-list rc=$?"
python tools/launch_summary.py $O/launches_step.csv 2 > $O/launch_summary.txt; head -40 $O/launch_summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_tc|conv_wgrad_halo|gn_se_cluster|finalconv_gs|gn_pool|gn_scale|gn_bwd_reduce|gn_bwd_apply|adam_kernel|upsample2x|conv_tap_gemm|finalconv_bwd_tc|s2_dgrad|ffl_|gn_bwd_params_batched|fc_wide|fc_reduce" -c 90 -o /tmp/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu-full rc=$?"
python tools/ncu_summary.py /tmp/prof_final.ncu-rep > $O/ncu_step_kernels.txt; wc -l $O/ncu_step_kernels.txt
ncu -i /tmp/prof_final.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_columns.py > $O/ncu_step_raw.csv; wc -c $O/ncu_step_raw.csv
ls -la $O
