set -x
timeout 600 python tools/halo_ablate.py > gpurun_out/halo_ablate.txt 2>&1; echo "ablate rc=$?"; cat gpurun_out/halo_ablate.txt | grep -v "^+" | tail -40
timeout 1500 python -m pytest tests/test_gpu_timed_workload.py tests/test_gpu_integration.py tests/test_gpu_train.py -x -q -s > gpurun_out/pytest_new.log 2>&1; echo "pytest-new rc=$?"
tail -30 gpurun_out/pytest_new.log
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
