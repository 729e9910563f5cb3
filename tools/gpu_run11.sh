set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "focal" > gpurun_out/pytest_ffl.log 2>&1; echo "pytest-ffl rc=$?"; tail -5 gpurun_out/pytest_ffl.log
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_timed_workload.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2i.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
