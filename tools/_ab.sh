#!/bin/bash
# back-to-back A/B of two library builds on the default bench: _ab.sh prev|work ... (alternating)
for i in 1 2 3; do for l in prev work; do
  if [ $l = prev ]; then export OP3_B200_LIB=$PWD/op3_b200/_build/var/lib_prev.so; else unset OP3_B200_LIB; fi
  timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/ab_${l}_$i.json 2> gpurun_out/ab_${l}_$i.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${l}_$i.json").read().strip().splitlines()[-1])
k=d["kernel_families"]
print("$i $l", round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], "conv %.2f wgrad %.2f"%(k["conv5x5 forward/dgrad"]["ms_per_step"], k["conv5x5 wgrad"]["ms_per_step"]))
PY
done; done
