timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all3.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/t_all3.log
