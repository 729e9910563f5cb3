set -x
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_timed_workload.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2h.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv --log-file gpurun_out/launches_r2_step.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2_step.csv 2 > gpurun_out/r2_launch_summary.txt; head -75 gpurun_out/r2_launch_summary.txt
