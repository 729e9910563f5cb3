set -x
timeout 600 python tools/smoke_debug.py > gpurun_out/smoke_debug.log 2>&1; grep -v "^$" gpurun_out/smoke_debug.log | tail -60
for sa in 4 5; do BVAE_HALO_SA=$sa timeout 600 python tools/halo_ablate.py 2>&1 | grep "dbg= 0" ; done
BVAE_GN_NO_CLUSTER=1 timeout 600 python tools/bench_ops.py --only gn > gpurun_out/bench_ops_gn_nocluster.txt 2>&1; grep gn_ gpurun_out/bench_ops_gn_nocluster.txt
timeout 600 python tools/bench_ops.py --only gn,up > gpurun_out/bench_ops_gn.txt 2>&1; grep "gn_\|upsample" gpurun_out/bench_ops_gn.txt
