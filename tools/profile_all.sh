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
      conv2d_fused) K="-k regex:conv_rows"; C=1;;
      maximum) K="-k regex:ew_vec"; C=1;;
      exp) K="-k regex:unary_vec"; C=1;;
      astype) K="-k regex:astype"; C=1;;
      transpose_nd) K="-k regex:transpose_tile"; C=1;;
      gather) K="-k regex:gather_rows"; C=1;;
      scatter_add) K="-k regex:scatter_add_sorted"; C=1;;
      sum) K="-k regex:scalar_reduce"; C=1;;
      add) K="-k regex:ew_vec"; C=1;;
      copy) K="-k regex:blas1_vec"; C=1;;
      transpose) K="-k regex:omatcopy"; C=1;;
      gemv) K="-k regex:gemv_n"; C=1;;
      gemv_trans) K="-k regex:gemv_t"; C=1;;
      reduce_sum) K="-k regex:reduce_"; C=1;;
      reduce_sum_axis0) K="-k regex:reduce_cols"; C=1;;
      softmax) K="-k regex:softmax"; C=1;;
      im2col2d) K="-k regex:im2col2d"; C=1;;
      im2col3d) K="-k regex:im2col3d"; C=1;;
      cumsum_rows) K="-k regex:scan_rows_vec"; C=1;;
      cumsum_cols) K="-k regex:scan_cols"; C=2;;
      topk) K="-k regex:topk"; C=1;;
    esac
    timeout 600 ncu --set full --clock-control none --import-source on $K -s $S -c $C \
      -o gpurun_out/prof/$op -f python tools/profile_ops.py $op 3 > gpurun_out/prof/$op.ncu.log 2>&1
    echo "$op ncu rc=$?"
    # the compact summary is what gets committed; reports are ~4 MB each and gpurun returns at most 64 MiB
    python tools/ncu_summary.py gpurun_out/prof/$op.ncu-rep gpurun_out/prof/ncu_$op.txt > /dev/null 2>&1
    case " ${KEEP_REP:-sgemm_tf32 add} " in
      *" $op "*) ;;
      *) rm -f gpurun_out/prof/$op.ncu-rep;;
    esac
  else
    echo "$op plain run FAILED"
  fi
done
