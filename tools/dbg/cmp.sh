for t in 32 64 128 256; do echo "tx $t" >> gpurun_out/c2.log; RB200_RD_TX=$t timeout 100 python tools/hbm_quick.py reduce_sum_axis0 2>&1 | head -1 >> gpurun_out/c2.log; done
