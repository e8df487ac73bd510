set -u
mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$?" > gpurun_out/final/rc.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$?" >> gpurun_out/final/rc.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final/rc.log
OPS="${FINAL_OPS:-sgemm_tf32 sgemm_tf32x3 add reduce_sum reduce_sum_axis0 softmax im2col2d im2col3d conv_gemm conv2d_fused copy transpose gemv gemv_trans maximum exp astype transpose_nd gather scatter_add sum cumsum_rows}" bash tools/profile_all.sh > gpurun_out/final/profile_all.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_plain.json 2> gpurun_out/final/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_ncu.log 2>&1
echo "launches rc=$?" >> gpurun_out/final/rc.log
