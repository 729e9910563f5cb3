set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn_relu or upsample" > gpurun_out/pytest_gn.log 2>&1; echo "pytest-gn rc=$?"; tail -25 gpurun_out/pytest_gn.log
timeout 600 python tools/bench_ops.py --only gn > gpurun_out/bench_ops_gn_stream.txt 2>&1; grep "gn_" gpurun_out/bench_ops_gn_stream.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_ops.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-gpu rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for m in d['roofline_more']: print(m['kernel'][:60], m['us_per_launch'], m['frac'])
PY
BVAE_GN_CLUSTER=1 timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cluster:', d['value'], d['ms_per_step'])"
