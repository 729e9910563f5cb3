set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/bench_n8_single.json 2> gpurun_out/bench_n8_single.err; echo "n8 single rc=$?"
grep "^{" gpurun_out/bench_n8_single.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d.get('dp_check',{}).get('identical'))"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/nccl_n8.log 2>&1; echo "nccl rc=$?"
grep -E "NVLS|Algo|Channel|nChannels|Connected|comm .* rank 0 |AllReduce" gpurun_out/nccl_n8.log | grep -v "Channel [0-9]*/[0-9]* :" | head -40 > gpurun_out/nccl_n8_excerpt.txt; wc -l gpurun_out/nccl_n8.log gpurun_out/nccl_n8_excerpt.txt; head -30 gpurun_out/nccl_n8_excerpt.txt
rm -f gpurun_out/nccl_n8.log
