set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn_relu" > gpurun_out/pytest_gn.log 2>&1; echo "pytest-gn rc=$?"; tail -5 gpurun_out/pytest_gn.log
timeout 600 python tools/bench_ops.py --only gn --reps 5 > gpurun_out/plain_gn.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gn.csv python tools/bench_ops.py --only gn --reps 5 > gpurun_out/ncu_gn.log 2>&1
grep "gn_" gpurun_out/plain_gn.log
timeout 1500 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_ops.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for m in d['roofline_more']: print(m['kernel'][:60], m['us_per_launch'], m['frac'])
PY
