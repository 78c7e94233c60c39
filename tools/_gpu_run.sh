timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/t_model2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t_model2.log
