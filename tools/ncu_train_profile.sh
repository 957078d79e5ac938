#!/bin/bash
# One `ncu --set full` capture per hot kernel of the training step (run on the GPU box through gpurun):
#   bash tools/ncu_train_profile.sh          -> gpurun_out/prof_train_<kernel>.ncu-rep
set -u
mkdir -p gpurun_out
for k in bn_bwd_apply_kernel bn_bwd_reduce_kernel dwconv_kernel wgrad_gemm_kernel dw3_wgrad_kernel bn_stats_kernel; do
  skip=12; [ $k = dw3_wgrad_kernel ] && skip=5
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_train_$k \
      python tools/ncu_train_target.py > gpurun_out/prof_train_$k.log 2>&1 || echo "ncu $k failed"
done
ls -la gpurun_out/*.ncu-rep
