timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv" > gpurun_out/c1.log 2>&1; echo rc=$? >> gpurun_out/c1.log
for d in 0 128 64 0; do echo "dbg $d" >> gpurun_out/c2.log; RB200_CR_DBG=$d timeout 100 python tools/hbm_quick.py conv2d_fused 2>&1 | head -1 >> gpurun_out/c2.log; done
