#!/bin/bash
# Runs on the GPU box (through gpurun): for each op, one plain run (must exit 0), then one ncu --set full
# capture of launches 2..3 of the same command.  Reports land in gpurun_out/prof/.
set -u
mkdir -p gpurun_out/prof
OPS=${OPS:-"sgemm_tf32 sgemm_tf32x3 add reduce_sum reduce_sum_axis0 softmax im2col2d conv_gemm copy transpose gemv gemv_trans"}
for op in $OPS; do
  if python tools/profile_ops.py $op 3 > gpurun_out/prof/$op.plain.log 2>&1; then
    K=""; S=1; C=2
    case $op in
      sgemm_tf32) K="-k regex:gemm_tf32"; C=1;;
      sgemm_tf32x3) K="-k regex:gemm_tf32"; C=1;;
      conv_gemm) K="-k regex:skinny"; C=1;;
    esac
    timeout 600 ncu --set full --clock-control none --import-source on $K -s $S -c $C \
      -o gpurun_out/prof/$op -f python tools/profile_ops.py $op 3 > gpurun_out/prof/$op.ncu.log 2>&1
    echo "$op ncu rc=$?"
  else
    echo "$op plain run FAILED"
  fi
done
