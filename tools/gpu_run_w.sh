set -x
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo "pytest-all rc=$?"; tail -3 gpurun_out/pytest_all.log
