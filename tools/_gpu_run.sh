timeout 900 python -m pytest tests/test_gpu_edge.py -x -q -m gpu -s > gpurun_out/t_edge.log 2>&1; echo "pytest rc=$?"
grep -E "^edge|passed|failed|Error|error|assert" gpurun_out/t_edge.log | head -30
