#!/bin/bash
# Round-2 ncu evidence, run on the GPU box through gpurun.  Every capture follows a plain run of the same command that
# exited 0; summaries (tools/ncu_summary.py) are what gets committed under profiles/.
set -u
P=gpurun_out/prof2
mkdir -p $P
cap() {  # name, kernel regex, skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  if "$@" > $P/$name.plain.log 2>&1; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $P/$name "$@" > $P/$name.ncu.log 2>&1
    echo "$name ncu rc=$?"
    python tools/ncu_summary.py $P/$name.ncu-rep $P/ncu_$name.txt > /dev/null 2>&1
  else
    echo "$name plain run FAILED"; tail -3 $P/$name.plain.log
  fi
}
cap sgemm_tf32_16384 gemm_tf32_2cta 3 python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline
cap sgemm_tf32_8192 gemm_tf32_2cta 3 python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline --size 8192
cap sgemm_tf32_1024 gemm_tf32_kernel 3 python tools/gemm_sweep.py tf32 1024
cap dgemm_dmma_4096 gemm_dmma 1 python tools/simt_probe.py 4096
cap conv_tma_64ch conv_tma 2 python tools/conv_quick.py - 0
rm -f $P/sgemm_tf32_8192.ncu-rep $P/sgemm_tf32_1024.ncu-rep
# launch list of the default bench command (after it exited 0 without ncu)
if python bench.py --steps 5 --warmup 3 > $P/bench_plain.json 2> $P/bench_plain.err; then
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/launches_bench_r02.csv \
     python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $P/bench_ncu.json 2> $P/bench_ncu.err
  echo "launch list rc=$?"
fi
ls -la $P | head -40
