set -x
bash tools/gpu_run11.sh
timeout 300 python tools/prof_wgrad.py > gpurun_out/plain_wgrad.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad" -c 12 -o gpurun_out/prof_r2_wgrad python tools/prof_wgrad.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu rc=$?"
