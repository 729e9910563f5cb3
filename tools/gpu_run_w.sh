set -x
timeout 900 python -m pytest tests/test_gpu_augment.py tests/test_gpu_train.py -x -q > gpurun_out/pytest_aug.log 2>&1; echo "pytest-aug rc=$?"; tail -30 gpurun_out/pytest_aug.log
