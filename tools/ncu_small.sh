mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"wgrad_kernel" -c 1 -o gpurun_out/ncu_wgrad_64 python tools/prof_conv.py 8 128 128 64 64 3 1 wgrad > gpurun_out/ncu_small.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"igemm2_kernel" -c 1 -o gpurun_out/ncu_fwd_64 python tools/prof_conv.py 8 128 128 64 64 3 1 fwd >> gpurun_out/ncu_small.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"igemm2_kernel" -c 1 -o gpurun_out/ncu_fwd_256_32 python tools/prof_conv.py 8 32 32 256 256 3 1 fwd >> gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out/*.ncu-rep
