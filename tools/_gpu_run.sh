set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t_kernels2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_kernels2.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/plain3.json 2> gpurun_out/plain3.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv5x5_wgrad_tc_kernel -s 8 -c 4 -o gpurun_out/prof_wgtc_r01 -f python bench.py --steps 1 --warmup 1 --precision tf32x3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "ncu rc=$?"
