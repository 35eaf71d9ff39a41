#!/bin/bash
# GPU box only: the round's standing check -- gpu tests, default bench, reference arm, ncu captures (tag = $1).
# .ncu-rep files are exported to csv/txt on the box and dropped (gpurun_out/ is capped at 64 MiB).
TAG=${1:-r01e}
WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "$WHAT" = all ] || [ "$WHAT" = test ]; then
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.log; tail -5 gpurun_out/${TAG}_bench.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -3 gpurun_out/${TAG}_bench_ref.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = ncu ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_full.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lrt-sims 1024 > gpurun_out/${TAG}_launches_bench.log 2>&1
export_rep() {   # name
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_ncu_full_raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_ncu_details.txt 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1_ncu_source.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_loglike -s 3 -c 1 -f -o gpurun_out/${TAG}_k1 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_k1.log 2>&1
export_rep ${TAG}_k1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2_loglike -s 1 -c 1 -f -o gpurun_out/${TAG}_k2 \
    python tools/k2_probe.py 151552 > gpurun_out/${TAG}_k2.log 2>&1
export_rep ${TAG}_k2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k8_ensemble -s 4 -c 1 -f -o gpurun_out/${TAG}_k8 \
    python tools/k8_probe.py > gpurun_out/${TAG}_k8.log 2>&1
export_rep ${TAG}_k8
fi
if [ "$WHAT" = all ] || [ "$WHAT" = ncu ] || [ "$WHAT" = ncu_scan ]; then
export_rep() {   # name
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_ncu_full_raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_ncu_details.txt 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1_ncu_source.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
# K9 at B = 32 (the half ensemble of config C2): launches of the probe are 21 per batch size, B = 1, 6, 16, 32
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k9_loglike_scan -s 63 -c 1 -f -o gpurun_out/${TAG}_k9 \
    python tools/k9_probe.py 1000 > gpurun_out/${TAG}_k9.log 2>&1
export_rep ${TAG}_k9
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k10_ensemble_scan -s 1 -c 1 -f -o gpurun_out/${TAG}_k10 \
    python tools/k10_probe.py > gpurun_out/${TAG}_k10.log 2>&1
export_rep ${TAG}_k10
fi
du -sh gpurun_out; ls -la gpurun_out/
