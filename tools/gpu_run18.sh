set -x
run() { n=$1; tag=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/bench_n${n}_$tag.json 2> gpurun_out/bench_n${n}_$tag.err; echo "$tag rc=$?"; grep "^{" gpurun_out/bench_n${n}_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d.get('dp_check',{}); print('$tag', round(d['value']), d['ms_per_step'], c.get('exchange'), c.get('exchange_us_isolated'), c.get('exchange_sums_exactly'), c.get('identical'), c.get('exchange_timeouts'))"; grep -i "error\|Traceback" gpurun_out/bench_n${n}_$tag.err | head -5; }
run 8 nvls64 BVAE_NVLS_BLOCKS=64
run 8 nccl BVAE_DP_MODE=single
