set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "conv" > gpurun_out/pytest_ops.log 2>&1; echo "pytest-ops rc=$?"; tail -5 gpurun_out/pytest_ops.log
timeout 600 python tools/halo_ablate.py > gpurun_out/halo_ablate.txt 2>&1; echo "ablate rc=$?"; grep "dbg= 0\|dbg= 4\|dbg= 2\|dbg=14" gpurun_out/halo_ablate.txt
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -9 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_timed_workload.py -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest-b rc=$?"; tail -5 gpurun_out/pytest_b.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['roofline']['us_per_launch'], d['roofline']['frac'])
for m in d['roofline_more']: print(m['kernel'][:60], m['us_per_launch'], m['frac'])
PY
