set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn" > gpurun_out/pytest_gn.log 2>&1; echo "pytest-gn rc=$?"; tail -2 gpurun_out/pytest_gn.log
for v in 1 0; do BVAE_GN_WAVE=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('wave=$v', d['ms_per_step'], [ (r['kernel'][:28], round(r['us_per_launch'],1)) for r in d['roofline_more']])"; done
