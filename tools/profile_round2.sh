#!/bin/bash
# GPU box only: round-2 ncu captures into gpurun_out/ (tag = $1, what = $2: all | new | k1 | launches)
TAG=${1:-r02a}
WHAT=${2:-all}
mkdir -p gpurun_out
export_rep() {   # name
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_ncu_full_raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_ncu_details.txt 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1_ncu_source.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
if [ "$WHAT" = all ] || [ "$WHAT" = launches ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_bench_full.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lrt-sims 1024 > gpurun_out/${TAG}_launches_bench.log 2>&1
fi
if [ "$WHAT" = all ] || [ "$WHAT" = new ]; then
# K11 at B = 148 (probe: 11 launches per kernel per batch size; sizes 1, 32, 148 -> the scan launches of B = 148 start at 2*22+11)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k11_loglike_grad_scan -s 24 -c 1 -f -o gpurun_out/${TAG}_k11 \
    python tools/k11_probe.py 1000 > gpurun_out/${TAG}_k11.log 2>&1
export_rep ${TAG}_k11
# the persistent L-BFGS kernel on the alternative fits of a 1250-simulation shard (second launch of the run = alternative model, rep 0)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lbfgs_persist -s 1 -c 1 -f -o gpurun_out/${TAG}_persist \
    python tools/lrt_probe.py 1250 fast 0 1 > gpurun_out/${TAG}_persist.log 2>&1
export_rep ${TAG}_persist
fi
if [ "$WHAT" = all ] || [ "$WHAT" = k1 ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_loglike -s 3 -c 1 -f -o gpurun_out/${TAG}_k1 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_k1.log 2>&1
export_rep ${TAG}_k1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2_loglike -s 1 -c 1 -f -o gpurun_out/${TAG}_k2 \
    python tools/k2_probe.py 10000 > gpurun_out/${TAG}_k2.log 2>&1
export_rep ${TAG}_k2
fi
du -sh gpurun_out; ls -la gpurun_out/ | grep ${TAG}
