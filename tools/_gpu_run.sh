timeout 900 python -m pytest tests/test_gpu_mpc_large.py -x -q -m gpu -s > gpurun_out/t_mpc.log 2>&1; echo "pytest rc=$?"
grep -E "vs reference|4096 cand|passed|failed|Error|assert" gpurun_out/t_mpc.log | head
timeout 300 python tools/bench_mpc.py f16x3 2>/dev/null | tail -1
timeout 300 python tools/bench_mpc.py fp32 2>/dev/null | tail -1
