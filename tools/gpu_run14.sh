set -x
for mode in single buckets; do
BVAE_DP_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "n2 $mode rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_n2_$mode.json'))
print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'))"
done
