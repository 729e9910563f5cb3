set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 beta-vae_b200/csrc/dev/umma_rate 512 > gpurun_out/umma_rate.txt 2>&1; echo "umma rc=$?"
tail -50 gpurun_out/umma_rate.txt
timeout 1500 python -m pytest tests/test_gpu_timed_workload.py tests/test_gpu_integration.py tests/test_gpu_train.py -x -q -s > gpurun_out/pytest_new.log 2>&1; echo "pytest-new rc=$?"
tail -30 gpurun_out/pytest_new.log
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
cat gpurun_out/bench_r2a.json | head -c 3000
for w in overfit_capacity overfit_capacity_wide128 inference; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_r2a_$w.json 2> gpurun_out/bench_r2a_$w.err; echo "$w rc=$?"; head -c 1500 gpurun_out/bench_r2a_$w.json; done
