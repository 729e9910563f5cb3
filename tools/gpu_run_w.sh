set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "conv3x3" > gpurun_out/pytest_conv.log 2>&1; echo "pytest-conv rc=$?"; tail -3 gpurun_out/pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_timed_workload.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2w.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'])"
grep -E "conv_wgrad" gpurun_out/bench_r2w.err | head -12
