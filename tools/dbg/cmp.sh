for r in 1 2 3 4 5 6 8; do echo "rows $r" >> gpurun_out/c2.log; RB200_IC3_ROWS=$r timeout 100 python tools/hbm_quick.py im2col3d 2>&1 | head -1 >> gpurun_out/c2.log; done
