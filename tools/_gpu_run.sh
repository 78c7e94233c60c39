timeout 900 python -m pytest tests/test_gpu_precision.py tests/test_gpu_scaled.py -x -q -m gpu > gpurun_out/t_prec3.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_prec3.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/b_x3.json 2> gpurun_out/b_x3.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/b_tf32.json 2> gpurun_out/b_tf32.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/b_x3.json','gpurun_out/b_tf32.json']:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(j['config']['precision'], j['value'], j['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items()})
PY
