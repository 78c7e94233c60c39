timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/t_all.log
for p in f16x3 tf32x3 tf32; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision $p > gpurun_out/b_$p.json 2> gpurun_out/b_$p.err; echo "bench $p rc=$?"
python - <<PY
import json
j=json.loads(open('gpurun_out/b_$p.json').read().strip().splitlines()[-1])
print(j['config']['precision'], round(j['value'],1), round(j['ms_per_step'],2), round(j['e2e']['value'],1), {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items()})
PY
done
