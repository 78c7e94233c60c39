timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv5x5_wgrad_tc_kernel|mix_bwd_kernel|conv5x5_wgrad_kernel" -s 4 -c 8 -o gpurun_out/prof_r01_bwd -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu6.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
