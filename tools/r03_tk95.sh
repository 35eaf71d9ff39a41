#!/bin/bash
# GPU box only: fused TK95 path (K12 / K13) -- parity tests, throughput against the cuFFT chain, ncu captures.  tag = $1
TAG=${1:-r03a}
mkdir -p gpurun_out
if [ "$3" != notests ]; then
timeout 900 python -m pytest tests/test_tk95_fused_gpu.py tests/test_simulator_gpu.py tests/test_simulator_stats_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
fi
for src in celerite table; do
timeout 300 python tools/sim_dev_probe.py 4096 fast $src 2>&1 | tail -2
MTG_SIM_FUSED=0 timeout 300 python tools/sim_dev_probe.py 4096 fast $src 2>&1 | tail -1
done
echo "--- host API"; timeout 300 python tools/sim_probe.py 10000 fast 2>&1 | tail -2
if [ "$2" != noncu ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_sim.csv \
    python tools/sim_dev_probe.py 1024 fast celerite > gpurun_out/${TAG}_launches_sim.log 2>&1
fi
if [ "$2" = full ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tk95_columns -s 1 -c 1 -f -o gpurun_out/${TAG}_k12 \
    python tools/sim_dev_probe.py 256 fast celerite > gpurun_out/${TAG}_k12.log 2>&1
ncu -i gpurun_out/${TAG}_k12.ncu-rep --page raw --csv > gpurun_out/${TAG}_k12_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k12.ncu-rep --page details > gpurun_out/${TAG}_k12_ncu_details.txt 2>/dev/null
ncu -i gpurun_out/${TAG}_k12.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_k12_ncu_source.csv.gz
rm -f gpurun_out/${TAG}_k12.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tk95_rows -s 1 -c 1 -f -o gpurun_out/${TAG}_k13 \
    python tools/sim_dev_probe.py 256 fast celerite > gpurun_out/${TAG}_k13.log 2>&1
ncu -i gpurun_out/${TAG}_k13.ncu-rep --page details > gpurun_out/${TAG}_k13_ncu_details.txt 2>/dev/null
ncu -i gpurun_out/${TAG}_k13.ncu-rep --page raw --csv > gpurun_out/${TAG}_k13_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k13.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_k13_ncu_source.csv.gz
rm -f gpurun_out/${TAG}_k13.ncu-rep
fi
ls -la gpurun_out | grep ${TAG}
