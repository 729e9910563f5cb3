set -x
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo "pytest-all rc=$?"; tail -5 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2z.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
