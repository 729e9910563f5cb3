# usage: gpu_final_multi.sh N  -- bench lines at N GPUs (default workload; inference replicas)
N=$1
O=gpurun_out/final
mkdir -p $O
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > $O/bench_default_n$N.out 2> $O/bench_default_n$N.err; echo "default rc=$?"
grep "^{" $O/bench_default_n$N.out > $O/bench_default_n$N.json; head -c 600 $O/bench_default_n$N.json; echo
if [ "$N" = "8" ] && [ -z "$SKIP_INFERENCE" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload inference --steps 30 --warmup 3 > $O/bench_inference_n$N.out 2> $O/bench_inference_n$N.err; echo "inference rc=$?"
grep "^{" $O/bench_inference_n$N.out > $O/bench_inference_n$N.json; head -c 600 $O/bench_inference_n$N.json; echo
fi
