set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "focal or stats_epilogue" > gpurun_out/pytest_ffl.log 2>&1; echo "pytest-ffl rc=$?"; tail -15 gpurun_out/pytest_ffl.log
timeout 1500 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_ops.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
PY
