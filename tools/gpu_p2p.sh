# usage: gpu_p2p.sh N  -- exchange algorithms side by side at N GPUs (short runs; dp_check reports the isolated exchange time)
N=$1; CFGS=${2:-"mc:64 p2p:64 p2p:128"}
O=gpurun_out/final
mkdir -p $O
set -x
for cfg in $CFGS; do
  A=${cfg%%:*}; B=${cfg##*:}
  BVAE_NVLS_ALGO=$A BVAE_NVLS_BLOCKS=$B timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 2> $O/p2p_n${N}_${A}${B}.err | grep "^{" > $O/r2_scale_n${N}_${A}${B}.json
  python -c "
import json
d=json.load(open('$O/r2_scale_n${N}_${A}${B}.json')); c=d['dp_check']
print('$A $B', d['ms_per_step'], d['value'], c['exchange'], c['exchange_sums_exactly'], c['identical'], c['exchange_timeouts'], c['exchange_us_isolated'])" || tail -5 $O/p2p_n${N}_${A}${B}.err
done
