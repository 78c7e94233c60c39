timeout 600 python bench.py > gpurun_out/bench_r01_v5.json 2> gpurun_out/bench_r01_v5.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_v5_reference.json 2>/dev/null; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step6.csv python tools/one_step.py > gpurun_out/ncu_step5.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv5x5_wgrad_ring_kernel -s 2 -c 2 -o gpurun_out/prof_r01_v5_ring -f python tools/one_step.py > gpurun_out/ncu_v5_ring.log 2>&1; echo "ncu ring rc=$?"
python -c "
import json
j=json.loads(open('gpurun_out/bench_r01_v5.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['frac_of_mode_ceiling'], j['clocks'], j['cpu_baseline']['value'])"
