#!/bin/bash
# usage: _gpu_multi.sh N "cfg1 cfg2 ..." [steps]
N=$1; CFGS=$2; STEPS=${3:-10}
mkdir -p gpurun_out
for c in $CFGS; do
  st=$STEPS; if [ "$c" = "c4" ] || [ "$c" = "c5-stack" ]; then st=3; fi
  if [ "$N" = "1" ]; then
    python bench.py --config $c --gpus 1 --steps $st --warmup 3 > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --config $c --gpus $N --steps $st --warmup 3 > gpurun_out/r02_bench_${c}_${N}gpu.json 2> gpurun_out/r02_bench_${c}_${N}gpu.err
  fi
  echo "$c x$N exit $?"; tail -c 400 gpurun_out/r02_bench_${c}_${N}gpu.json | head -c 400; echo
done
