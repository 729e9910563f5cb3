set -x
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_n8_$tag.json 2> gpurun_out/bench_n8_$tag.err; echo "$tag rc=$?"; grep "^{" gpurun_out/bench_n8_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$tag', round(d['value']), d['ms_per_step'], d.get('dp_check',{}).get('identical'))"; }
run skip BVAE_DP_SKIP_ALL=1
run bf16 BVAE_DP_REDUCE_DTYPE=bf16
run ring NCCL_ALGO=Ring
run nvls16 NCCL_NVLS_NCHANNELS=16 NCCL_MAX_CTAS=16
