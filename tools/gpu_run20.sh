set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/pytest_ops.log 2>&1; echo "pytest-ops rc=$?"; tail -3 gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_timed_workload.py tests/test_gpu_inference.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2k.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for m in d['roofline_more']: print(m['kernel'][:60], m['us_per_launch'], m['frac'])"
timeout 600 python tools/bench_ops.py --only gn,final > gpurun_out/bench_ops_gn.txt 2>&1; grep "128x128\|final" gpurun_out/bench_ops_gn.txt
