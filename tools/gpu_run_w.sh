set -x
O=gpurun_out/final2
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_style.json 2> $O/bench_driver_style.err; echo "bench-driver rc=$?"
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > $O/bench_default_n1.json 2> $O/bench_default_n1.err; echo "bench rc=$?"
for w in overfit_capacity overfit_capacity_wide128; do timeout 600 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_${w}_n1.json 2> $O/bench_${w}_n1.err; echo "$w rc=$?"; done
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
for f in bench_driver_style bench_default_n1 bench_overfit_capacity_n1 bench_overfit_capacity_wide128_n1 bench_reference_arm; do python - <<PY
import json
d=json.loads([l for l in open('$O/$f.json') if l.startswith('{')][-1])
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('gpu_launches'), (d.get('roofline') or {}).get('frac'), [(r['kernel'][:24], round(r['us_per_launch'],1), round(r['frac'],3)) for r in d.get('roofline_more', [])])
PY
done
