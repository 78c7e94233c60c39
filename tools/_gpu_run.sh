timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_scaled.py -x -q -m gpu > gpurun_out/t_km.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/t_km.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_f16.json 2> gpurun_out/b_f16.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/b_f16.json']:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(j['config']['precision'], j['value'], j['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items()}, j['roofline_mixture']['ms_per_call'], j['roofline_mixture']['frac'])
PY
