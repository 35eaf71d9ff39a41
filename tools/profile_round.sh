#!/bin/bash
# GPU box only: ncu captures of the round's top kernels into gpurun_out/ (tag = $1)
TAG=${1:-r01d}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_full.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_loglike -s 3 -c 1 -f -o gpurun_out/${TAG}_k1 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_loglike -s 1 -c 1 -f -o gpurun_out/${TAG}_k2 \
    python tools/k2_probe.py 65536 > gpurun_out/${TAG}_k2.log 2>&1
ls -la gpurun_out/
