#!/bin/bash
# ncu --set full captures of the round-2 kernels inside one trainer step of the bench workload (tools/one_step.py)
mkdir -p gpurun_out
python tools/one_step.py f16x3 > gpurun_out/one_step_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/one_step_plain.log; exit 1; }
for k in conv5x5_tcq_kernel conv5x5_wgrad_ring_kernel mix_fused_kernel mix_bwd_kernel dyn_fwd_fused_kernel dyn_bwd_fused_kernel ref_head_fwd_kernel ref_head_bwd_kernel gemm_wgrad_batch_kernel; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -f -o gpurun_out/r02_$k python tools/one_step.py f16x3 > gpurun_out/ncu_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/ncu_$k.log)"
done
